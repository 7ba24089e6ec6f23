"""CPU tests of the checker itself: the oracle against the definition of the
DFT, closed-form known answers, numpy/scipy, the UNMODIFIED reference files
(oracle/_ref, when built) and the committed golden vectors they produced."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
I32_MIN = np.iinfo(np.int32).min


def _golden(name):
    return np.load(os.path.join(GOLDEN, name))


# ------------------------------------------------------------------ DFT
@pytest.mark.parametrize("n", [256, 1024, 2048, 4096, 16384])
def test_fft_matches_numpy(oracle, n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    ref = np.fft.fft(x)
    assert np.abs(oracle.fft(x) - ref).max() <= 1e-12 * np.abs(ref).max()


def test_fft_matches_definition(oracle):
    rng = np.random.default_rng(7)
    x = rng.integers(-32768, 32768, 512).astype(np.float64)
    a, b = oracle.fft(x), oracle.dft_naive(x)
    assert np.abs(a - b).max() <= 1e-9 * np.abs(b).max()


def test_known_answers_single_frame(oracle):
    n = 2048
    cfg = oracle.reference_config(n, 512, 100)
    t = np.arange(n)
    # on-bin sine: |X[5]| = A*N/2 -> 1024 at zoom 1e-4, nothing elsewhere
    pcm = np.round(10000 * np.sin(2 * np.pi * 5 * t / n)).astype(np.int16)
    mag, h, hi = oracle.frame(cfg, pcm)
    assert abs(h[5] - 1024.0) < 0.01 and hi[5] == 1024 and np.delete(hi, 5).max() == 0
    # impulse: flat magnitude A
    pcm = np.zeros(n, np.int16); pcm[3] = 12345
    mag, h, hi = oracle.frame(cfg, pcm)
    assert np.allclose(mag, 12345.0, rtol=0, atol=1e-8) and (hi == 1).all()
    # DC: everything in bin 0
    pcm = np.full(n, -32768, np.int16)
    mag, h, hi = oracle.frame(cfg, pcm)
    assert mag[0] == 32768.0 * n and np.abs(mag[1:]).max() < 1e-6 and hi[0] == 6710
    # Parseval on white noise
    rng = np.random.default_rng(1)
    pcm = rng.integers(-32768, 32768, n).astype(np.int16)
    mag, _, _ = oracle.frame(cfg, pcm)
    full = np.concatenate([mag, mag[-2:0:-1]])
    assert np.isclose((full ** 2).sum() / n, (pcm.astype(np.float64) ** 2).sum(), rtol=1e-12)


def test_trunc_matches_x86_semantics(oracle):
    L = oracle.lib()
    assert L.orc_trunc_i32(3.99, 1) == 3 and L.orc_trunc_i32(-3.99, 1) == -3
    assert L.orc_trunc_i32(float("nan"), 1) == I32_MIN and L.orc_trunc_i32(float("inf"), 1) == I32_MIN
    assert L.orc_trunc_i32(2147483648.0, 1) == I32_MIN and L.orc_trunc_i32(2147483647.9, 1) == 2147483647
    assert L.orc_trunc_i32(3e9, 0) == 2147483647 and L.orc_trunc_i32(float("nan"), 0) == 0


# ------------------------------------------------------------------ extended mode vs numpy / scipy
def test_extended_mode_matches_numpy_scipy(oracle):
    from scipy.signal import savgol_filter
    cfg = oracle.extended_config(4096, 512, 200)
    pcm = oracle.synth_pcm(1, 4096 + 512 * 4)
    out = oracle.stft(cfg, pcm, want_beat=False)["bars_f64"][0]
    lo, hi = oracle.build_bins(cfg)
    # independent restatement with numpy
    e = 20.0 * (20000.0 / 20.0) ** (np.arange(201) / 200.0)
    lo2 = np.clip(np.floor(e[:-1] * 4096 / 48000.0).astype(int), 0, 2048)
    hi2 = np.minimum(np.maximum(np.floor(e[1:] * 4096 / 48000.0).astype(int), lo2 + 1), 2049)
    assert (lo == lo2).all() and (hi == hi2).all()
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(4096) / 4096)
    assert np.allclose(oracle.window(cfg), win, atol=1e-15)
    for f in range(out.shape[0]):
        x = pcm[0, f * 512:f * 512 + 4096].astype(np.float64) * win
        mag = np.abs(np.fft.rfft(x))
        raw = np.array([mag[a:b].mean() for a, b in zip(lo2, hi2)])
        ref = savgol_filter(raw, 7, 2, mode="interp") * float(np.float32(1e-4))
        assert np.abs(out[f] - ref).max() <= 1e-9 * np.abs(ref).max()


@pytest.mark.parametrize("w,d", [(1, 1), (2, 2), (3, 2), (4, 3), (8, 4)])
def test_sg_projection_matches_scipy(oracle, w, d):
    from scipy.signal import savgol_coeffs
    P = oracle.build_sg(w, d)
    m = 2 * w + 1
    for pos in range(m):
        ref = savgol_coeffs(m, d, pos=pos, use="dot")
        assert np.abs(P[pos] - ref).max() < 1e-10


# ------------------------------------------------------------------ beat chain
def test_add_beat_sample_step_response(oracle):
    C = __import__("ctypes")
    st = oracle.BeatState()
    L = oracle.lib()
    L.orc_beat_reset(C.byref(st))
    # first call with w=255 -> threshold 255/160 = 1 -> beatValue 32640
    L.orc_add_beat_sample(C.byref(st), 255)
    assert st.thr == 1 and st.beat_value == 32640
    for _ in range(159):
        L.orc_add_beat_sample(C.byref(st), 255)
    assert st.thr == 255 and st.beat_value == 128
    L.orc_add_beat_sample(C.byref(st), 0)
    assert st.thr == (255 * 159) // 160 and st.beat_value == 0
    # threshold 0: x/0 and 0/0 give INT32_MIN like cvttss2si
    L.orc_beat_reset(C.byref(st))
    L.orc_add_beat_sample(C.byref(st), 0)
    assert st.thr == 0 and st.beat_value == I32_MIN
    L.orc_add_beat_sample(C.byref(st), 100)
    assert st.thr == 0 and st.beat_value == I32_MIN


def test_golden_add_beat_sample(oracle):
    g = _golden("ref_add_beat_sample.npz")
    rep = oracle.beat_replay(oracle.reference_config(), g["w"], np.zeros_like(g["w"]))
    assert (rep["thr"] == g["thr"]).all() and (rep["beat_value"] == g["beat_value"]).all()


def test_peak_pick_cases(oracle):
    hb = np.zeros(161, np.int32)
    assert oracle.peak_pick(hb) == 0            # flat: every index is a "peak", first 30 only, none > 30
    hb[:] = np.arange(161)[::-1]
    assert oracle.peak_pick(hb) == 0            # monotone down: only index 0
    hb[:] = 0; hb[50] = 7; hb[90] = 9; hb[120] = 9
    # zeros 0..29 use up the 30 peak slots before index 50 is reached
    assert oracle.peak_pick(hb) == 0
    hb[:] = np.arange(161); hb[50] = 500; hb[90] = 900; hb[120] = 900
    assert oracle.peak_pick(hb) == 90           # highest peak beyond index 30, first of equals


def test_am_bass_tone_gives_modulation_peak(oracle):
    """2 Hz AM on a 94 Hz tone -> modulation peak at 2 Hz * 256 * hop / fs."""
    fs, n, hop = 48000, 2048, 512
    t = np.arange(n + hop * 600) / fs
    x = 6000 * (0.5 + 0.5 * np.sin(2 * np.pi * 2.0 * t)) * np.sin(2 * np.pi * 93.75 * t)
    pcm = np.round(x).astype(np.int16)[None, :]
    r = oracle.stft(oracle.reference_config(n, hop, 100), pcm)
    k = r["beat"]["peak_index"][0, 300:]
    expect = 2.0 * 256 * hop / fs
    assert abs(np.median(k) - expect) <= 1.0 or np.median(k) == 0


# ------------------------------------------------------------------ golden vectors from the unmodified reference
@pytest.mark.parametrize("name", ["ref_c1_stereo.npz", "ref_mono_circle.npz"])
def test_oracle_matches_reference_golden(oracle, name):
    g = _golden(name)
    ch, S = int(g["channels"]), int(g["samples"])
    first = int(g["first_sample"]) if "first_sample" in g else 0
    pcm = oracle.synth_pcm(ch, S, first_sample=first)
    assert int(pcm.astype(np.int64).sum()) == int(g["pcm_sum"]) and (pcm[:, :32] == g["pcm_head"]).all()
    cfg = oracle.reference_config(2048, int(g["hop"]), int(g["bars"]), zoom=float(g["zoom"]), circle=int(g["circle"]))
    r = oracle.stft(cfg, pcm)
    d = r["bars_i32"].astype(np.int64) - g["heights"]
    assert np.abs(d).max() <= 1
    # the reference's fp32 FFT and the oracle's fp64 one may straddle an integer: only near boundaries
    frac = np.abs(r["bars_f64"] - np.rint(r["bars_f64"]))
    fmax = np.abs(r["bars_f64"]).max(axis=-1, keepdims=True)
    assert (frac[d != 0] <= np.broadcast_to(1e-4 * fmax, frac.shape)[d != 0]).all()
    assert (d != 0).mean() < 1e-3
    # integer chain replayed on the reference's own integers: bit exact
    last = ch - 1  # BassBeatDetector sees the last executed spectrum (right channel when stereo)
    rep = oracle.beat_replay(cfg, g["w"], g["bass"])
    assert (rep["thr"] == g["thr"]).all() and (rep["beat_value"] == g["beat_value"]).all()
    # tempo peak index: the reference's fp32 transform against the oracle's fp64 one.  Every frame where they differ
    # must be an integer-boundary flip of the scan (helpers._peak_possible); the reference also reads
    # heightBuffer[160] out of bounds (beatDetector.c:87), so that one value is unknown.
    from helpers import _peak_possible, TEMPO_REL
    n_dt = np.float32(np.float32(2048) * (np.float32(int(g["hop"])) / np.float32(48000.0)))
    ref_peak = np.rint(g["movement"] / n_dt).astype(np.int64)
    assert (g["movement"] == n_dt * ref_peak.astype(np.float32)).all()
    differ = np.nonzero(rep["movement_per_sec"] != g["movement"])[0]
    mag, norm = oracle.tempo_mag_replay(g["bass"])
    for f in differ:
        assert _peak_possible(mag[f], 1.0 + TEMPO_REL * norm[f], int(ref_peak[f]), oob=True), "frame %d unexplained" % f
    assert np.abs(r["beat"]["bass"][last].astype(np.int64) - g["bass"]).max() <= 1


def test_port_is_bit_identical_to_reference_golden(oracle):
    """The fp32 CPU port (the cpu_baseline) against the unmodified reference's outputs."""
    g = _golden("ref_c1_stereo.npz")
    pcm = oracle.synth_pcm(2, int(g["samples"]))
    cfg = oracle.reference_config(2048, int(g["hop"]), int(g["bars"]))
    _, p = oracle.port_run(cfg, pcm, threads=2)
    assert (p["bars_i32"] == g["heights"]).all()
    assert (p["beat"]["w"][0] == g["w"]).all() and (p["beat"]["thr"][0] == g["thr"]).all()
    assert (p["beat"]["beat_value"][0] == g["beat_value"]).all()
    assert (p["beat"]["bass"][1] == g["bass"]).all()


def test_live_reference_agrees_with_golden(oracle):
    """When oracle/_ref is present (build container), re-run the unmodified reference."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref not built here")
    g = _golden("ref_c1_stereo.npz")
    pcm = oracle.synth_pcm(2, int(g["samples"]))
    _, heights, fo = oracle.ref_run(pcm, int(g["hop"]), int(g["bars"]))
    assert (heights == g["heights"]).all() and (fo["w"] == g["w"]).all() and (fo["beat_value"] == g["beat_value"]).all()
    # reference FFT provider (shim) against the oracle's double DFT
    R = oracle.ref()
    R.ref_open(0)
    try:
        rng = np.random.default_rng(3)
        x = (rng.integers(-32768, 32768, 2048) + 0j).astype(np.complex64)
        out = np.empty(2048, np.complex64)
        import ctypes as C
        R.ref_execute(x.view(np.float32).ctypes.data_as(C.POINTER(C.c_float)), out.view(np.float32).ctypes.data_as(C.POINTER(C.c_float)))
    finally:
        R.ref_close()
    ref = oracle.fft(x.astype(np.complex128))
    assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()


def test_port_multithread_ranges_equal_single_thread(oracle):
    cfg = oracle.extended_config(1024, 256, 64)
    pcm = oracle.synth_pcm(2, 1024 + 256 * 1500)
    _, a = oracle.port_run(cfg, pcm, threads=1)
    _, b = oracle.port_run(cfg, pcm, threads=4)
    assert (a["bars"] == b["bars"]).all()
    for k in ("w", "thr", "beat_value", "flag", "bass", "peak_index"):
        assert (a["beat"][k] == b["beat"][k]).all(), k


def test_stft_shard_equals_full(oracle):
    cfg = oracle.reference_config(2048, 512, 100)
    pcm = oracle.synth_pcm(1, 2048 + 512 * 700)
    full = oracle.stft(cfg, pcm)
    part = oracle.stft(cfg, pcm, 300, 650)
    assert (part["bars_i32"] == full["bars_i32"][:, 300:650]).all()
    for k in ("w", "thr", "beat_value", "flag", "bass", "peak_index", "movement_per_sec"):
        assert (part["beat"][k] == full["beat"][k][:, 300:650]).all(), k


def test_validation_and_empty(oracle):
    import ctypes as C
    c = oracle.reference_config(2048, 512, 100)
    assert oracle.lib().orc_config_validate(C.byref(c)) == 0
    for field, bad in (("n", 3000), ("n", 128), ("hop", 0), ("bars", 0), ("bars", 2000), ("window", 5), ("bass_bin", 5000)):
        c2 = oracle.reference_config(2048, 512, 100)
        setattr(c2, field, bad)
        assert oracle.lib().orc_config_validate(C.byref(c2)) != 0, (field, bad)
    assert oracle.num_frames(c, 2047) == 0 and oracle.num_frames(c, 2048) == 1 and oracle.num_frames(c, 2048 + 511) == 1
    out = oracle.stft(c, np.zeros((2, 100), np.int16))
    assert out["bars"].shape == (2, 0, 100)


@pytest.mark.parametrize("provider", ["scipy", "torch"])
def test_library_fft_chain_matches_oracle(oracle, provider):
    """oracle/numpy_chain.py (the `pocketfft` / `torch.fft` CPU baseline lines of bench.py) computes the same chain:
    float bars within 1e-4 of the frame maximum of the fp64 oracle, integer beat stages exact when replayed."""
    from oracle import numpy_chain as NC
    from helpers import assert_bars_close, _peak_possible, TEMPO_REL
    cfg = oracle.extended_config(4096, 512, 200)
    pcm = oracle.synth_pcm(2, 4096 + 699 * 512)
    got = NC.run(oracle, cfg, pcm, provider=provider, block=256)
    ref = oracle.stft(cfg, pcm)
    assert_bars_close(got["bars"], ref["bars_f64"], provider)
    assert np.abs(got["w"] - ref["beat"]["w"]).max() <= 1 and np.abs(got["bass"] - ref["beat"]["bass"]).max() <= 1
    for ch in range(2):
        rep = oracle.beat_replay(cfg, got["w"][ch].astype(np.int32), got["bass"][ch].astype(np.int32))
        for k in ("thr", "beat_value", "flag"):
            assert (rep[k] == got[k][ch]).all(), k
        differ = np.nonzero(rep["peak_index"] != got["peak_index"][ch])[0]
        mag, norm = oracle.tempo_mag_replay(got["bass"][ch].astype(np.int32))
        for f in differ:
            assert _peak_possible(mag[f], 1.0 + TEMPO_REL * norm[f], int(got["peak_index"][ch][f])), f
