"""Build recipe for the CUDA library (nvcc, sm_100a only; cross-compiles without a GPU)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SOURCES = [os.path.join(HERE, "csrc", "wavernn_b200.cu")]
DEPENDS = SOURCES + [os.path.join(HERE, "csrc", "wavernn_kernel.cuh"), os.path.join(HERE, "csrc", "wavernn_dense.cuh"), os.path.join(HERE, "csrc", "wavernn_wide.cuh"), os.path.join(ROOT, "include", "wavernn_b200.h")]
LIB = os.path.join(HERE, "libwavernn_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in DEPENDS)


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libwavernn_b200.so next to this file (in-tree, git-ignored)."""
    if not force and not needs_build():
        return LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
    if verbose:
        print(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
