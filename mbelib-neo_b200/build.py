"""Builds libmbe_b200.so (CUDA kernels + C-ABI) in-tree for sm_100a with nvcc.

    python mbelib-neo_b200/build.py [--force]

nvcc cross-compiles without a GPU; the resulting .so is git-ignored but travels to the GPU box."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmbe_b200.so")
SOURCES = ["mbe_b200.cu", "mbe_pool.cpp"]
DEPS = ["mbe_b200.cu", "mbe_pool.cpp", "mbe_common.cuh", "mbe_frontend.cuh", "mbe_parms.cuh", "mbe_synth.cuh", "mbe_split.cuh", "mbe_libm.cuh",
        "mbe_tables.inc", "mbe_exp2_tab.inc", os.path.join("..", "..", "include", "mbe_b200.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # the reference's float arithmetic is unfused; FMAs are spelled out where glibc has them
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2",
    "-shared",
]


SHIM_OUT = os.path.join(HERE, "libmbe-neo-b200shim.so")
SHIM_DEPS = ["mbe_single_shim.c", os.path.join("..", "..", "include", "mbe_b200.h"),
             os.path.join("..", "..", "include", "mbe_b200_compat.h")]


def build_shim(force=False):
    """The single-stream drop-in shim: plain C on top of the C-ABI, linked against libmbe_b200.so."""
    if not force and os.path.exists(SHIM_OUT) and os.path.exists(OUT):
        t = os.path.getmtime(SHIM_OUT)
        if t >= os.path.getmtime(OUT) and all(os.path.getmtime(os.path.join(CSRC, d)) <= t for d in SHIM_DEPS):
            return SHIM_OUT
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-std=gnu99", "-Wall", "-fPIC", "-fvisibility=hidden", "-shared", "-o", SHIM_OUT,
           os.path.join(CSRC, "mbe_single_shim.c"), "-L" + HERE, "-lmbe_b200", "-Wl,-rpath,$ORIGIN", "-lpthread"]
    subprocess.check_call(cmd)
    return SHIM_OUT


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    if not force and not needs_build():
        build_shim()
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("MBE_NVCC_EXTRA", "").split()
    out = os.environ.get("MBE_LIB_OUT", OUT)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    if out == OUT:
        build_shim(force=True)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
