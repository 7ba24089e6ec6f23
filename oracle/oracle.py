"""ctypes binding of the CPU oracle (oracle/libfftl_oracle.so) and, when it was
built, of the unmodified-reference harness (oracle/_ref/libfftlines_ref.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg.  The product package
(fft_lines_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libfftl_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libfftlines_ref.so")
REFERENCE_DIR = "/root/reference/fft_lines"

BEAT_SAMPLES = 160
TEMPO_RING = 256


class Config(C.Structure):
    """Mirror of fftl_config (include/fftlines.h)."""
    _fields_ = [
        ("n", C.c_int32), ("hop", C.c_int32), ("bars", C.c_int32), ("window", C.c_int32),
        ("binning", C.c_int32), ("sg_half", C.c_int32), ("sg_degree", C.c_int32),
        ("circle", C.c_int32), ("zoom", C.c_float), ("sample_rate", C.c_float),
        ("fmin", C.c_float), ("fmax", C.c_float), ("beat_bars", C.c_int32 * 4),
        ("bass_bin", C.c_int32), ("strict_int", C.c_int32), ("reserved", C.c_int32 * 3),
    ]


class Beat(C.Structure):
    """Mirror of fftl_beat."""
    _fields_ = [
        ("w", C.c_int32), ("thr", C.c_int32), ("beat_value", C.c_int32), ("flag", C.c_int32),
        ("bass", C.c_int32), ("peak_index", C.c_int32), ("movement_per_sec", C.c_float),
        ("reserved", C.c_int32),
    ]


BEAT_DTYPE = np.dtype([("w", "<i4"), ("thr", "<i4"), ("beat_value", "<i4"), ("flag", "<i4"),
                       ("bass", "<i4"), ("peak_index", "<i4"), ("movement_per_sec", "<f4"),
                       ("reserved", "<i4")])


class BeatState(C.Structure):
    _fields_ = [
        ("ring", C.c_int32 * BEAT_SAMPLES), ("pos", C.c_int32), ("sum", C.c_int32),
        ("thr", C.c_int32), ("beat_value", C.c_int32), ("tempo_ring", C.c_int32 * TEMPO_RING),
        ("tempo_pos", C.c_int32), ("hb", C.c_int32 * (BEAT_SAMPLES + 1)),
        ("peak_index", C.c_int32), ("movement_per_sec", C.c_float),
    ]


class RefFrameOut(C.Structure):
    _fields_ = [("w", C.c_int32), ("thr", C.c_int32), ("beat_value", C.c_int32),
                ("bass", C.c_int32), ("movement_per_sec", C.c_float)]


REF_FRAME_DTYPE = np.dtype([("w", "<i4"), ("thr", "<i4"), ("beat_value", "<i4"), ("bass", "<i4"),
                            ("movement_per_sec", "<f4")])


def reference_config(n=2048, hop=512, bars=100, zoom=1e-4, sample_rate=48000.0, circle=0):
    """What the reference computes (SURVEY: parity mode)."""
    c = Config()
    c.n, c.hop, c.bars = n, hop, bars
    c.window, c.binning, c.sg_half, c.sg_degree, c.circle = 0, 0, 0, 0, circle
    c.zoom, c.sample_rate, c.fmin, c.fmax = zoom, sample_rate, 20.0, 20000.0
    c.beat_bars[:] = [3, 4, 5, 6]
    c.bass_bin, c.strict_int = 4, 1
    return c


def extended_config(n=4096, hop=512, bars=200, zoom=1e-4, sample_rate=48000.0, fmin=20.0,
                    fmax=20000.0, sg_half=3, sg_degree=2, circle=0):
    """north_star's chain: Hann, LOG bins, Savitzky-Golay (oracle-defined)."""
    c = reference_config(n, hop, bars, zoom, sample_rate, circle)
    c.window, c.binning, c.sg_half, c.sg_degree = 1, 1, sg_half, sg_degree
    c.fmin, c.fmax = fmin, fmax
    return c


def build(force=False):
    """Compile the oracle (always) and oracle/_ref (when the reference tree is present)."""
    if force or not os.path.exists(ORACLE_SO) or _stale(ORACLE_SO, ["fftl_oracle.c", "cpu_port.c", "fftwf_shim.c", "fftl_oracle.h"]):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libfftl_oracle.so"])
    if os.path.isdir(REFERENCE_DIR) and (force or not os.path.exists(REF_SO) or _stale(REF_SO, ["ref_harness.c", "fftwf_shim.c"])):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "ref"])


def _stale(so, srcs):
    t = os.path.getmtime(so)
    return any(os.path.getmtime(os.path.join(_HERE, s)) > t for s in srcs)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        i32p, f32p, f64p, i16p = (C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_double),
                                  C.POINTER(C.c_int16))
        L.orc_trunc_i32.restype = C.c_int32
        L.orc_trunc_i32.argtypes = [C.c_double, C.c_int]
        L.orc_fft.argtypes = [C.c_int, f64p, f64p, f64p, f64p]
        L.orc_dft_naive.argtypes = [C.c_int, f64p, f64p, f64p, f64p]
        L.orc_config_validate.argtypes = [C.POINTER(Config)]
        L.orc_build_bins.argtypes = [C.POINTER(Config), i32p, i32p]
        L.orc_build_sg.argtypes = [C.c_int, C.c_int, f64p]
        L.orc_window.argtypes = [C.POINTER(Config), f64p]
        L.orc_num_frames.restype = C.c_int64
        L.orc_num_frames.argtypes = [C.POINTER(Config), C.c_int64]
        L.orc_frame.argtypes = [C.POINTER(Config), i16p, C.c_int64, f64p, f64p, i32p]
        L.orc_beat_reset.argtypes = [C.POINTER(BeatState)]
        L.orc_add_beat_sample.argtypes = [C.POINTER(BeatState), C.c_int32]
        L.orc_bass_beat.argtypes = [C.POINTER(BeatState), C.c_int32, C.c_int, C.c_float, C.c_int]
        L.orc_peak_pick.restype = C.c_int32
        L.orc_peak_pick.argtypes = [i32p]
        L.orc_tempo_hb.argtypes = [i32p, C.c_int, i32p]
        L.orc_tempo_mag_replay.argtypes = [i32p, C.c_int64, C.c_int64, f64p, f64p]
        L.orc_stft.argtypes = [C.POINTER(Config), i16p, C.c_int32, C.c_int64, C.c_int64, C.c_int64,
                               C.c_int64, C.c_int64, f64p, f32p, i32p, C.c_void_p]
        L.orc_beat_replay.argtypes = [C.POINTER(Config), i32p, i32p, C.c_int64, C.c_int64, C.c_void_p, i32p]
        L.orc_synth_pcm.argtypes = [i16p, C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_uint32]
        L.orc_settings_parse.argtypes = [C.c_char_p, C.c_long, i32p, f32p]
        L.orc_range255.restype = C.c_uint32
        L.orc_range255.argtypes = [C.c_int, C.c_int]
        L.orc_led_byte.argtypes = [C.c_int32, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8)]
        L.orc_marker_sweep.argtypes = [f32p, C.c_int64, C.c_float, C.c_int32, C.c_int32, i32p, i32p]
        L.port_max_threads.restype = C.c_int
        L.port_run.restype = C.c_double
        L.port_run.argtypes = [C.POINTER(Config), i16p, C.c_int32, C.c_int64, C.c_int64, C.c_int64,
                               C.c_int64, C.c_int64, C.c_int, f32p, i32p, C.c_void_p]
        _lib = L
    return _lib


def ref_available():
    return os.path.exists(REF_SO)


def ref():
    """The unmodified reference files behind ref_harness.c (None if not built)."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            return None
        R = C.CDLL(REF_SO)
        i32p, f32p, i16p = C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_int16)
        R.ref_compiled_n.restype = C.c_int
        R.ref_open.argtypes = [C.c_int]
        R.ref_execute.argtypes = [f32p, f32p]
        R.ref_execute256.argtypes = [f32p, f32p]
        R.ref_add_beat_sample.argtypes = [C.c_int]
        R.ref_get_beat_value.restype = C.c_int
        R.ref_get_threshold.restype = C.c_int
        R.ref_get_movement.restype = C.c_float
        R.ref_tick.argtypes = [i16p, i16p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_float, i32p, i32p,
                               C.POINTER(RefFrameOut)]
        R.ref_run.restype = C.c_double
        R.ref_run.argtypes = [i16p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_float, C.c_int,
                              C.c_int, C.c_float, i32p, C.c_void_p]
        _ref = R
    return _ref


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


# ------------------------------------------------------------------ helpers
def num_frames(cfg, samples):
    return int(lib().orc_num_frames(C.byref(cfg), samples))


def fft(x):
    x = np.asarray(x, dtype=np.complex128)
    n = x.shape[-1]
    ir, ii = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
    orr, oi = np.empty(n), np.empty(n)
    lib().orc_fft(n, _p(ir, C.c_double), _p(ii, C.c_double), _p(orr, C.c_double), _p(oi, C.c_double))
    return orr + 1j * oi


def dft_naive(x):
    x = np.asarray(x, dtype=np.complex128)
    n = x.shape[-1]
    ir, ii = np.ascontiguousarray(x.real), np.ascontiguousarray(x.imag)
    orr, oi = np.empty(n), np.empty(n)
    lib().orc_dft_naive(n, _p(ir, C.c_double), _p(ii, C.c_double), _p(orr, C.c_double), _p(oi, C.c_double))
    return orr + 1j * oi


def build_bins(cfg):
    lo = np.empty(cfg.bars, np.int32)
    hi = np.empty(cfg.bars, np.int32)
    lib().orc_build_bins(C.byref(cfg), _p(lo, C.c_int32), _p(hi, C.c_int32))
    return lo, hi


def build_sg(w, d):
    m = 2 * w + 1
    coef = np.empty((m, m), np.float64)
    rc = lib().orc_build_sg(w, d, _p(coef, C.c_double))
    if rc:
        raise ValueError("bad SG parameters")
    return coef


def window(cfg):
    w = np.empty(cfg.n, np.float64)
    lib().orc_window(C.byref(cfg), _p(w, C.c_double))
    return w


def frame(cfg, pcm):
    """One frame -> (mag[n/2+1], height_f64[bars], bars_i32[bars])."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    assert pcm.shape[-1] >= cfg.n
    mag = np.empty(cfg.n // 2 + 1)
    h = np.empty(cfg.bars)
    hi = np.empty(cfg.bars, np.int32)
    lib().orc_frame(C.byref(cfg), _p(pcm, C.c_int16), 1, _p(mag, C.c_double), _p(h, C.c_double), _p(hi, C.c_int32))
    return mag, h, hi


def stft(cfg, pcm, frame_begin=0, frame_end=None, want_beat=True):
    """pcm: int16 [channels, samples] planar. Returns dict of arrays."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    if pcm.ndim == 1:
        pcm = pcm[None, :]
    ch, S = pcm.shape
    F = num_frames(cfg, S)
    if frame_end is None:
        frame_end = F
    nf = frame_end - frame_begin
    f64 = np.empty((ch, nf, cfg.bars), np.float64)
    f32 = np.empty((ch, nf, cfg.bars), np.float32)
    i32 = np.empty((ch, nf, cfg.bars), np.int32)
    beat = np.zeros((ch, nf), BEAT_DTYPE) if want_beat else None
    rc = lib().orc_stft(C.byref(cfg), _p(pcm, C.c_int16), ch, S, S, 1, frame_begin, frame_end,
                        _p(f64, C.c_double), _p(f32, C.c_float), _p(i32, C.c_int32),
                        beat.ctypes.data if want_beat else None)
    if rc:
        raise ValueError("orc_stft failed rc=%d" % rc)
    return {"bars_f64": f64, "bars": f32, "bars_i32": i32, "beat": beat}


def beat_replay(cfg, w, bass, first_frame=0, want_hb=False):
    w = np.ascontiguousarray(w, np.int32)
    bass = np.ascontiguousarray(bass, np.int32)
    n = w.shape[0]
    out = np.zeros(n, BEAT_DTYPE)
    hb = np.empty((n, BEAT_SAMPLES + 1), np.int32) if want_hb else None
    lib().orc_beat_replay(C.byref(cfg), _p(w, C.c_int32), _p(bass, C.c_int32), n, first_frame,
                          out.ctypes.data, _p(hb, C.c_int32))
    return (out, hb) if want_hb else out


def peak_pick(hb):
    hb = np.ascontiguousarray(hb, np.int32)
    assert hb.shape[-1] == BEAT_SAMPLES + 1
    return int(lib().orc_peak_pick(_p(hb, C.c_int32)))


def tempo_hb(ring, strict=1):
    ring = np.ascontiguousarray(ring, np.int32)
    hb = np.empty(BEAT_SAMPLES + 1, np.int32)
    lib().orc_tempo_hb(_p(ring, C.c_int32), strict, _p(hb, C.c_int32))
    return hb


def tempo_mag_replay(bass, first_frame=0):
    """fp64 magnitudes |X256[0..160]| of the tempo ring after every frame (before the reference's (int) cast,
    beatDetector.c:73) and the ring's Euclidean norm: (mag [frames,161], norm [frames])."""
    bass = np.ascontiguousarray(bass, np.int32)
    n = bass.shape[0]
    mag = np.empty((n, BEAT_SAMPLES + 1), np.float64)
    norm = np.empty(n, np.float64)
    lib().orc_tempo_mag_replay(_p(bass, C.c_int32), n, first_frame, _p(mag, C.c_double), _p(norm, C.c_double))
    return mag, norm


def synth_pcm(channels, samples, first_sample=0, sample_rate=48000.0, seed=0x5EED0000):
    pcm = np.empty((channels, samples), np.int16)
    lib().orc_synth_pcm(_p(pcm, C.c_int16), channels, samples, samples, first_sample, sample_rate, seed)
    return pcm


def port_run(cfg, pcm, frame_begin=0, frame_end=None, threads=1, want_bars=True, want_i32=True,
             want_beat=True):
    """fp32 multi-threaded CPU port.  Returns (seconds, dict)."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    if pcm.ndim == 1:
        pcm = pcm[None, :]
    ch, S = pcm.shape
    F = num_frames(cfg, S)
    if frame_end is None:
        frame_end = F
    nf = frame_end - frame_begin
    f32 = np.empty((ch, nf, cfg.bars), np.float32) if want_bars else None
    i32 = np.empty((ch, nf, cfg.bars), np.int32) if want_i32 else None
    beat = np.zeros((ch, nf), BEAT_DTYPE) if want_beat else None
    sec = lib().port_run(C.byref(cfg), _p(pcm, C.c_int16), ch, S, S, 1, frame_begin, frame_end, threads,
                         _p(f32, C.c_float), _p(i32, C.c_int32), beat.ctypes.data if want_beat else None)
    if sec < 0:
        raise ValueError("port_run failed")
    return sec, {"bars": f32, "bars_i32": i32, "beat": beat}


def ref_run(pcm, hop, bars, zoom=1e-4, circle=0, beat=True, sample_rate=48000.0, n=None):
    """Run the UNMODIFIED reference files over strided frames. pcm [1|2, S] planar int16."""
    R = ref()
    if R is None:
        raise RuntimeError("oracle/_ref not built (needs /root/reference)")
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    if pcm.ndim == 1:
        pcm = pcm[None, :]
    ch, S = pcm.shape
    assert ch in (1, 2)
    nn = n or R.ref_compiled_n()
    if R.ref_open(nn) != 0:
        raise RuntimeError("ref_open failed")
    try:
        R.ref_reset_beat()
        F = 0 if S < nn else (S - nn) // hop + 1
        heights = np.zeros((ch, F, bars), np.int32)
        fo = np.zeros(F, REF_FRAME_DTYPE)
        sec = R.ref_run(_p(pcm, C.c_int16), ch, S, S, hop, bars, zoom, circle, 1 if beat else 0,
                        sample_rate, _p(heights, C.c_int32), fo.ctypes.data)
    finally:
        R.ref_close()
    return sec, heights, fo


def settings_parse(text):
    """readSettings restated: returns ([11 ints, zoom slot unused], zoom)."""
    raw = text.encode() if isinstance(text, str) else bytes(text)
    v = np.zeros(11, np.int32)
    z = C.c_float()
    lib().orc_settings_parse(raw, len(raw), _p(v, C.c_int32), C.byref(z))
    return v, z.value


def range255(i, count):
    return int(lib().orc_range255(i, count))


def led_bytes(heights):
    heights = np.ascontiguousarray(heights, np.int32)
    b = np.zeros(heights.shape, np.uint8)
    v = np.zeros(heights.shape, np.uint8)
    bb, vv = C.c_uint8(), C.c_uint8()
    for i, h in enumerate(heights.ravel()):
        lib().orc_led_byte(int(h), C.byref(bb), C.byref(vv))
        b.ravel()[i], v.ravel()[i] = bb.value, vv.value
    return b, v


def marker_sweep(movement, time_delta, pos=0, direction=0):
    movement = np.ascontiguousarray(movement, np.float32)
    p = np.zeros(movement.shape[0], np.int32)
    d = np.zeros(movement.shape[0], np.int32)
    lib().orc_marker_sweep(_p(movement, C.c_float), movement.shape[0], time_delta, pos, direction, _p(p, C.c_int32), _p(d, C.c_int32))
    return p, d
