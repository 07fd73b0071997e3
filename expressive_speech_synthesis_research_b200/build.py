"""Build recipe for the CUDA library (nvcc, sm_100a only; cross-compiles without a GPU)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SOURCES = [os.path.join(HERE, "csrc", "wavernn_b200.cu")]
DEPENDS = SOURCES + [os.path.join(HERE, "csrc", "wavernn_kernel.cuh"), os.path.join(HERE, "csrc", "wavernn_dense.cuh"), os.path.join(HERE, "csrc", "wavernn_wide.cuh"), os.path.join(HERE, "csrc", "wavernn_cond.cuh"), os.path.join(ROOT, "include", "wavernn_b200.h")]
LIB = os.path.join(HERE, "libwavernn_b200.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)")


STAMP = LIB + ".stamp"


def source_digest():
    """sha256 of everything the library is built from (sources, headers, flags): what `needs_build` compares, so that file times
    rewritten by a checkout or a copy to another box never trigger (or hide) a rebuild."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for path in DEPENDS:
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_digest()


def build(force=False, verbose=False):
    """Compile csrc/*.cu into libwavernn_b200.so next to this file (in-tree, git-ignored).  Safe when several processes import the
    package at once (one rank per GPU): one of them builds under a file lock into a temporary name and renames it into place, the
    others wait for the lock and find the library current."""
    import fcntl
    if not force and not needs_build():
        return LIB
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():
                return LIB
            digest = source_digest()
            tmp = "%s.tmp.%d" % (LIB, os.getpid())
            cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp] + SOURCES
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.remove(tmp)
                raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
            os.replace(tmp, LIB)
            with open(STAMP + ".tmp", "w") as f:
                f.write(digest + "\n")
            os.replace(STAMP + ".tmp", STAMP)
            if verbose:
                print(res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
