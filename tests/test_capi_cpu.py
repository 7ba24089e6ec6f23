"""CPU-side tests of the product's host logic and of the C-ABI surface: the
library loads without a GPU, exports every symbol include/*.h declares, its
host tables agree with the oracle's, and it fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INCLUDE = os.path.join(ROOT, "include")


def _declared_functions(header):
    """Function names declared in a header (prototype = identifier followed by '(' at depth 0)."""
    src = open(os.path.join(INCLUDE, header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r"^\s*#.*$", "", src, flags=re.M)
    src = src.replace('extern "C" {', "")
    names = []
    for stmt in src.split(";"):
        m = re.search(r"\b([A-Za-z_]\w*)\s*\((?!\s*\*)", stmt)
        if m and "typedef" not in stmt and "{" not in stmt.split("(")[0]:
            names.append(m.group(1))
    return [n for n in names if n not in ("defined",)]


def test_library_exports_every_declared_symbol(lib):
    from fft_lines_b200 import _capi
    L = _capi.lib()
    nm = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = {line.split()[-1] for line in nm.splitlines() if line.strip()}
    for header, names in _capi.EXPORTED_FUNCTIONS.items():
        declared = _declared_functions(header)
        assert sorted(declared) == sorted(names), (header, sorted(set(declared) ^ set(names)))
        for n in names:
            assert n in exported, "%s (declared in %s) is not exported" % (n, header)
            getattr(L, n)
    for header, names in _capi.EXPORTED_DATA.items():
        for n in names:
            assert n in exported
    # the reference's four globals are plain data symbols of the right size
    assert _capi.data_symbol("beatMovementPerSec", C.c_float).value == 0.0
    _capi.data_symbol("timeDelta", C.c_float).value = 0.25
    assert lib.beat_detector.get_timeDelta() == 0.25
    lib.beat_detector.set_timeDelta(0.0)


def test_struct_layouts_match_header(lib):
    from fft_lines_b200 import _capi
    assert C.sizeof(_capi.FftlConfig) == 84 and C.sizeof(_capi.FftlBeat) == 32
    src = open(os.path.join(INCLUDE, "fftlines.h")).read()
    body = src[src.index("typedef struct fftl_config {"):src.index("} fftl_config;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"(?:int32_t|float)\s+(\w+)", body)
    assert fields == [f[0] for f in _capi.FftlConfig._fields_]


def test_configs_and_validation(lib):
    from fft_lines_b200 import _capi
    L = _capi.lib()
    c = lib.reference_config(2048, 512, 100)
    assert (c.window, c.binning, c.sg_half, c.bass_bin, list(c.beat_bars)) == (0, 0, 0, 4, [3, 4, 5, 6])
    assert abs(c.zoom - 1e-4) < 1e-10 and c.strict_int == 1
    assert L.fftl_config_validate(C.byref(c)) == 0
    x = lib.extended_config()
    assert (x.n, x.hop, x.bars, x.window, x.binning, x.sg_half, x.sg_degree) == (4096, 512, 200, 1, 1, 3, 2)
    for field, bad in (("n", 3000), ("n", 128), ("n", 32768), ("hop", 0), ("bars", 0), ("bars", 1026), ("window", 2),
                       ("binning", 3), ("sg_half", 9), ("bass_bin", 1025)):
        b = lib.reference_config(2048, 512, 100)
        setattr(b, field, bad)
        L.fftl_clear_error()
        assert L.fftl_config_validate(C.byref(b)) == _capi.FFTL_ERR_ARG, (field, bad)
        code, msg = _capi.last_error()
        assert code == _capi.FFTL_ERR_ARG and msg
    L.fftl_clear_error()
    assert _capi.last_error() == (0, "")


def test_host_tables_equal_oracle(lib, oracle):
    for n, bars, fs in ((4096, 200, 48000.0), (1024, 64, 48000.0), (16384, 512, 44100.0), (2048, 1000, 96000.0)):
        c = lib.extended_config(n, 512, bars, sample_rate=fs, fmax=min(20000.0, fs / 2))
        oc = oracle.extended_config(n, 512, bars, sample_rate=fs, fmax=min(20000.0, fs / 2))
        lo, hi = lib.build_bins(c)
        lo2, hi2 = oracle.build_bins(oc)
        assert (lo == lo2).all() and (hi == hi2).all()
        assert (hi > lo).all() and hi.max() <= n // 2 + 1 and (np.diff(lo) >= 0).all()
    for w, d in ((1, 1), (3, 2), (5, 3), (8, 4)):
        assert np.abs(lib.build_sg(w, d) - oracle.build_sg(w, d)).max() < 2e-7
    c = lib.reference_config(2048, 512, 500)
    lo, hi = lib.build_bins(c)
    assert (lo == np.arange(500)).all() and (hi == lo + 1).all()  # bar i = bin i (fft_lines.c:199-203)


def test_no_gpu_fails_loudly_no_cpu_fallback(lib):
    """On a machine without a CUDA device every compute entry point errors; nothing is computed on the CPU."""
    from fft_lines_b200 import _capi
    if lib.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_capi.FftLinesError) as ei:
        lib.BatchedSpectrum(lib.reference_config())
    assert ei.value.code == _capi.FFTL_ERR_CUDA
    inp = np.ones(2048, np.complex64)
    out = np.full(2048, 7 + 7j, np.complex64)
    L = _capi.lib()
    L.fftl_clear_error()
    lib.fft_calculate.initializefft(inp, out)
    assert _capi.last_error()[0] == _capi.FFTL_ERR_CUDA
    L.fftl_clear_error()
    lib.fft_calculate.executefft(inp, out)
    assert _capi.last_error()[0] in (_capi.FFTL_ERR_CUDA, _capi.FFTL_ERR_STATE)
    assert (out == 7 + 7j).all()           # outputs untouched
    L.fftl_clear_error()
    lib.beat_detector.AddBeatSample(5)
    assert _capi.last_error()[0] == _capi.FFTL_ERR_CUDA and lib.beat_detector.GetBeatValue() == 0
    lib.fft_calculate.cleanupfft()         # idempotent, no crash
    lib.fft_calculate.cleanupfft()
    L.fftl_clear_error()


def test_set_n(lib):
    fc = lib.fft_calculate
    assert fc.get_n() == 2048               # reference N (global.h:3)
    fc.set_n(4096)
    assert fc.get_n() == 4096
    from fft_lines_b200 import _capi
    with pytest.raises(_capi.FftLinesError):
        fc.set_n(3000)
    fc.set_n(2048)
    _capi.lib().fftl_clear_error()


def test_product_does_not_reference_the_oracle():
    """The shipped path must not import, include, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "fft_lines_b200")
    bad = [r"^\s*(from|import)\s+oracle\b", r"#\s*include\s+\"[^\"]*oracle", r"\borc_\w+", r"libfftl_oracle",
           r"cpu_port", r"fftwf_shim", r"_ref/"]
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                for pat in bad:
                    assert not re.search(pat, src, flags=re.M), (f, pat)
    out = subprocess.run(["ldd", os.path.join(pkg, "libfftlines_cuda.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out and "fftw" not in out
    nm = subprocess.run(["nm", "-D", os.path.join(pkg, "libfftlines_cuda.so")], capture_output=True, text=True).stdout
    assert "orc_" not in nm and "port_run" not in nm
