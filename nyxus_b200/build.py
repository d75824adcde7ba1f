"""Builds the native code in-tree (sm_100a only):
  nyxus_b200/libnyxus_b200.so            the CUDA library behind the C ABI (include/nyxus_b200.h), nvcc
  nyxus_b200/backend.<ext-suffix>.so     the pybind11 module `backend` over that C ABI (csrc/backend_py.cpp), g++
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libnyxus_b200.so")
BACKEND = os.path.join(HERE, "backend" + sysconfig.get_config_var("EXT_SUFFIX"))
SRC = [os.path.join(HERE, "csrc", "api.cu")]
BACKEND_SRC = os.path.join(HERE, "csrc", "backend_py.cpp")
HEADER = os.path.join(ROOT, "include", "nyxus_b200.h")
DEPS = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc")) if not f.endswith(".cpp")] + [HEADER]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    return _stale(LIB, DEPS)


def build_backend(force=False, verbose=False):
    """the pybind11 module; links libnyxus_b200.so from its own directory ($ORIGIN)"""
    if not force and not _stale(BACKEND, [BACKEND_SRC, HEADER, LIB]):
        return BACKEND
    import pybind11
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-fvisibility=hidden", "-Wall",
           "-I", sysconfig.get_paths()["include"], "-I", pybind11.get_include(), "-I", os.path.join(ROOT, "include"),
           BACKEND_SRC, "-o", BACKEND, "-L", HERE, "-lnyxus_b200", "-Wl,-rpath,$ORIGIN", "-lz"]
    if verbose:
        print(" ".join(cmd))
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("g++ failed building the pybind11 backend module")
    if verbose:
        print(r.stdout + r.stderr)
    return BACKEND


def build(force=False, verbose=False):
    if force or needs_build():
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
    build_backend(force=force, verbose=verbose)
    return LIB


if __name__ == "__main__":
    build(force="--backend-only" not in sys.argv, verbose="-v" in sys.argv)
    if "--backend-only" in sys.argv:
        build_backend(force=True, verbose="-v" in sys.argv)
    print("built", LIB, BACKEND)
