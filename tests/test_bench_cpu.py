"""CPU: bench.py's reference arm (`--impl reference`: the C restatement of the reference path on the host threads) runs without
a GPU and prints the contract's JSON line -- for the default workload and for the rolling one."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"] + list(args),
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("extra", [["--steps", "2", "--warmup", "1"],
                                   ["--workload", "c5", "--batch", "48", "--steps", "1", "--warmup", "1"]])
def test_reference_arm_json_line(extra):
    j = _run(*extra)
    assert j["impl"] == "reference" and j["unit"] == "env-steps/s" and j["higher_is_better"] is True
    assert j["metric"].startswith("env-steps/sec") and j["value"] > 0 and j["n_gpus"] == 1
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "sample" in cb
    e = j["e2e"]
    assert e["value"] == j["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in j["config"] and 0.0 < j["reward_mean"] < 1.0


def test_reference_arm_nonzero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
