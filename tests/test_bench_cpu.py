"""CPU: bench.py's reference arm (`--impl reference`) runs without a GPU and prints the contract's JSON line: the unmodified
Python reference in worker processes (kind "reference") when the reference tree is available, the C restatement (kind
"port") with --port, for the rolling workload, and as the fallback.  Both arms print the same `config` object."""
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


@pytest.mark.parametrize("extra,kind", [(["--steps", "2", "--warmup", "1", "--ref-envs-per-core", "8"], "reference"),
                                        (["--steps", "2", "--warmup", "1", "--port"], "port"),
                                        (["--workload", "c5", "--batch", "48", "--steps", "1", "--warmup", "1"], "port")])
def test_reference_arm_json_line(extra, kind):
    from oracle import refshim
    if kind == "reference" and not refshim.available():
        pytest.skip("reference tree not present")
    j = _run(*extra)
    assert j["impl"] == "reference" and j["unit"] == "env-steps/s" and j["higher_is_better"] is True
    assert j["metric"].startswith("env-steps/sec") and j["value"] > 0 and j["n_gpus"] == 1
    cb = j["cpu_baseline"]
    assert cb["kind"] == kind and cb["cores"] >= 1 and cb["value"] == j["value"] and "sample" in cb
    if kind == "reference":
        assert cb["value_one_core"] > 0 and cb["port"]["value"] > cb["value"]
    e = j["e2e"]
    assert e["value"] == j["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in j["config"] and 0.0 < j["reward_mean"] < 1.0


def test_reference_arm_nonzero_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_both_arms_print_the_same_config():
    """The driver compares the two arms' `config` objects (same_config): it is a pure function of workload, batch, world."""
    import argparse
    sys.path.insert(0, ROOT)
    import bench
    args = argparse.Namespace(no_graph=False, no_reduce=False, nccl_reduce=False)
    j = _run("--steps", "1", "--warmup", "1", "--port")
    assert j["config"] == bench.bench_config("c2", 4096, 1, args)
    assert bench.default_batch("c4", 8) == 1024 and bench.default_batch("c5", 8) == 8192 and bench.default_batch("c2", 8) == 4096
