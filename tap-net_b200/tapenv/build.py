"""Build the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU).

    python tap-net_b200/tapenv/build.py        # or  tapenv.build.build()
"""
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)                      # tap-net_b200/
CSRC = os.path.join(ROOT, "csrc")
LIB_DIR = os.path.join(ROOT, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libtapenv.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(ROOT), "include", "tapenv.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """Compile csrc/*.cu into lib/libtapenv.so.  Returns the library path.
    `defines` / `out` build tuning variants (e.g. TAPENV_WARPS_PER_CTA=8) next to the default library."""
    path = LIB_PATH if out is None else os.path.join(LIB_DIR, out)
    if out is None and not force and not _stale():
        return path
    os.makedirs(LIB_DIR, exist_ok=True)
    cmd = [_nvcc()] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", path] + sources()
    subprocess.check_call(cmd)
    return path


if __name__ == "__main__":
    print(build(force=True, verbose=True))
