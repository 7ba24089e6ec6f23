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
# translation units: (source, extra defines, object name).  The fast path's template instantiations are the
# bulk of the compile time; one unit per hop variant lets them build in parallel.
# The fast-path units use Blackwell's packed FP32 instructions (FFTL_F32X2, fft_device.cuh); the compiler folds the
# swizzles and negations of the packed forms into operand modifiers only without flush-to-zero, hence -ftz=false there.
F32X2 = ["-DFFTL_F32X2", "-ftz=false"]
UNITS = [("api.cu", [], "api"),
         ("rdif_launch.cu", ["-DFFTL_RDIF_HOPQ=0"] + F32X2, "rdif_h0"), ("rdif_launch.cu", ["-DFFTL_RDIF_HOPQ=1"] + F32X2, "rdif_h1"),
         ("rdif_launch.cu", ["-DFFTL_RDIF_HOPQ=2"] + F32X2, "rdif_h2"), ("rdif_launch.cu", ["-DFFTL_RDIF_HOPQ=4"] + F32X2, "rdif_h4")]
SOURCES = sorted({u[0] for u in UNITS})
HEADERS = ["kernels.cuh", "device_common.cuh", "fft_device.cuh", "stft_rdif.cuh", "stft_rdif_tm.cuh", "rdif_launch.h", "stream.cuh", "stream_tick.cuh", "stream_carry.cuh", "../../include/fftlines.h",
           "../../include/fft_calculate.h", "../../include/beatDetector.h", "../../include/fftl_complex.h"]


def kernel_source_hash():
    """Short hash of the sources the dominant kernel is built from: profiles/traffic.json is stamped with it so that
    bench.py can tell a capture of other code from a current one."""
    import hashlib
    h = hashlib.sha256()
    for f in ("stft_rdif.cuh", "stft_rdif_tm.cuh", "fft_device.cuh", "device_common.cuh", "rdif_launch.cu", "rdif_launch.h"):
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


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


def build(force=False, verbose=False, extra=(), out=None, dev=False):
    """out: alternative output path (developer A/B builds with extra -D flags); dev: -DFFTL_DEV (environment knobs of
    the host path, never in the shipped library)."""
    if dev:
        extra = tuple(extra) + ("-DFFTL_DEV",)
    if out is None and not force and not needs_build():
        return LIB
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    target = out or LIB
    common = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--use_fast_math",
              "-std=c++17", "-Xcompiler", "-fPIC,-O2,-fvisibility=default"]
    if verbose:
        common += ["-Xptxas", "-v"]
    with tempfile.TemporaryDirectory(prefix="fftl_build_") as tmp:
        def compile_unit(unit):
            src, defs, name = unit
            obj = os.path.join(tmp, name + ".o")
            if "-DFFTL_NO_F32X2" in extra:  # developer A/B builds: the scalar forms
                defs = [d for d in defs if d not in F32X2]
            cmd = common + list(extra) + defs + ["-c", os.path.join(CSRC, src), "-o", obj]
            r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            return obj, cmd, r
        with ThreadPoolExecutor(max_workers=min(len(UNITS), os.cpu_count() or 1)) as pool:
            results = list(pool.map(compile_unit, UNITS))
        for obj, cmd, r in results:
            if verbose or r.returncode != 0:
                print(" ".join(cmd))
            if r.stdout.strip():
                print(r.stdout, end="")
            if r.returncode != 0:
                raise subprocess.CalledProcessError(r.returncode, cmd)
        link = [nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o", target] + [o for o, _, _ in results]
        if verbose:
            print(" ".join(link))
        subprocess.check_call(link)
    return target


if __name__ == "__main__":
    if "--dev" in sys.argv:
        print(build(force=True, verbose="-v" in sys.argv, dev=True, out=os.path.join(HERE, "libfftlines_cuda_dev.so")))
    else:
        build(force="--force" in sys.argv, verbose="-v" in sys.argv)
        print(LIB)
