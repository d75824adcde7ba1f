"""Builds the CUDA library in-tree: nyxus_b200/libnyxus_b200.so (sm_100a only)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libnyxus_b200.so")
SRC = [os.path.join(HERE, "csrc", "api.cu")]
DEPS = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))] + [os.path.join(ROOT, "include", "nyxus_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
           "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-o", LIB] + SRC
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libnyxus_b200.so")
    if verbose:
        print(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print("built", LIB)
