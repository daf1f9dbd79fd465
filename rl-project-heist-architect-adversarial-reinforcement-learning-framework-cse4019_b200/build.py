"""In-tree nvcc build of the CUDA library (sm_100a only)."""
import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libheist_b200.so")
DBG_LIB_PATH = os.path.join(LIB_DIR, "libheist_b200_dbg.so")  # -DHEIST_DEBUG_BOUNDS: range-checked cell-map accesses

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # parity: CPython never contracts a*b+c; FMAs are written explicitly where wanted
    "-shared", "-Xcompiler", "-fPIC",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build libheist_b200.so")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))) + \
        [os.path.join(os.path.dirname(_HERE), "include", "heist_b200.h")]


def is_stale(path=None):
    path = path or LIB_PATH
    if not os.path.exists(path):
        return True
    t = os.path.getmtime(path)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    """Compile csrc/heist_b200.cu -> lib/libheist_b200.so (cross-compiles without a GPU)."""
    os.makedirs(LIB_DIR, exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None)   # the image's $CC wrapper is not a usable nvcc host compiler
    env.pop("CXX", None)
    for path, extra in ((LIB_PATH, []), (DBG_LIB_PATH, ["-DHEIST_DEBUG_BOUNDS"])):
        if not force and not is_stale(path):
            continue
        cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + \
            ["-o", path, os.path.join(CSRC, "heist_b200.cu")]
        res = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
        if verbose:
            print(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
