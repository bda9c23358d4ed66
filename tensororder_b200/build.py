"""Builds the CUDA extension in-tree: tensororder_b200/csrc/libtob200.so (sm_100a only).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the
repository snapshot.  `python -m tensororder_b200.build` or `__graft_entry__.build()`."""
import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")
LIB = os.path.join(CSRC, "libtob200.so")
SOURCES = ["tob_compile.cpp", "tob_kernels.cu", "tob_exec.cu"]
HEADERS = ["tob_internal.h", "tob_kernels.cuh", "tob_dispatch_table.h", os.path.join("..", "..", "include", "tob200.h")]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC", "-shared",
]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libtob200.so")
    if verbose:
        sys.stderr.write(res.stdout + res.stderr)
    return LIB


# ---- the compiled flatten loop (Cython, like the reference's own src/setup.py modules) ----
PKG = os.path.dirname(os.path.abspath(__file__))
FLATTEN_PYX = os.path.join(PKG, "_flatten_fast.pyx")


def flatten_ext_path():
    import sysconfig

    return os.path.join(PKG, "_flatten_fast" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_flatten(force=False):
    """cython -> C++ -> g++ -shared, in-tree.  The product path works without it (flatten.py has the same loop
    in Python: host logic, not arithmetic); it takes ~40 % off the host side of a contract_sliced call."""
    import sysconfig

    out = flatten_ext_path()
    if not force and os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(FLATTEN_PYX):
        return out
    import numpy

    cpp = os.path.join(PKG, "_flatten_fast.cpp")
    res = subprocess.run([sys.executable, "-m", "cython", "-3", "--cplus", FLATTEN_PYX, "-o", cpp],
                         capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("cython failed on _flatten_fast.pyx")
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-shared", "-fPIC", "-std=c++17", "-w",
           "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
           "-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(), cpp, "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("g++ failed building _flatten_fast")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_flatten(force="--force" in sys.argv))
