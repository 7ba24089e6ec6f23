"""Builds fft_lines_b200/libfftlines_cuda.so in-tree with nvcc for sm_100a.

The library is self-contained (static cudart); it loads on a machine without
a GPU (every compute call then fails loudly), which is what the CPU-side
symbol tests rely on.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libfftlines_cuda.so")
SOURCES = ["api.cu"]
HEADERS = ["kernels.cuh", "fft_device.cuh", "stft_rdif.cuh", "stft_rdif_ws2.cuh", "stream.cuh", "../../include/fftlines.h", "../../include/fft_calculate.h",
           "../../include/beatDetector.h", "../../include/fftl_complex.h"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    files = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(f) > t for f in files)


def build(force=False, verbose=False, extra=(), out=None):
    """out: alternative output path (developer A/B builds with extra -D flags)."""
    if out is None and not force and not needs_build():
        return LIB
    cmd = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--use_fast_math",
           "-std=c++17", "-shared", "-Xcompiler", "-fPIC,-O2,-fvisibility=default", "-cudart", "static",
           "-o", out or LIB] + list(extra) + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return out or LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
