"""Property / known-answer tests of the GPU path that SURVEY 8c lists (the reference ships none):
random int16 frames against the O(N^2) definition of the DFT, Parseval on white noise, the amplitude-modulated
bass tone's modulation peak, exact-integer magnitudes, and the capture-format (float32) input conversion."""
import ctypes as C

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from helpers import PAIR_LEAK, assert_bars_close, assert_int_heights, assert_peaks_explained

pytestmark = pytest.mark.gpu


def _ocfg(O, cfg):
    oc = O.Config()
    C.memmove(C.byref(oc), C.byref(cfg), C.sizeof(oc))
    return oc


@settings(max_examples=12, deadline=None, derandomize=True)  # the same examples in every run: a CI suite must not be a lottery
@given(n=st.sampled_from([256, 512, 1024, 2048, 4096]), seed=st.integers(0, 2 ** 31 - 1),
       shape=st.sampled_from(["uniform", "sparse", "loud", "tiny", "square"]))
def test_random_frames_against_the_definition(oracle, lib, n, seed, shape):
    """|X[k]| * zoom of random int16 frames vs orc_dft_naive (the plain double sum): 1e-4 of the frame maximum for the
    float heights, the integer rule for (int) heights; both the generic kernel (256, 512) and the fast path."""
    rng = np.random.default_rng(seed)
    frames, hop = 6, n // 4
    S = n + (frames - 1) * hop
    if shape == "uniform":
        x = rng.integers(-32768, 32768, S)
    elif shape == "sparse":
        x = np.where(rng.random(S) < 0.01, rng.integers(-32768, 32768, S), 0)
    elif shape == "loud":
        x = rng.choice([-32768, 32767], S)
    elif shape == "tiny":
        x = rng.integers(-2, 3, S)
    else:
        x = np.where((np.arange(S) // int(rng.integers(1, 40))) % 2 == 0, 12345, -12345)
    pcm = x.astype(np.int16)[None, :]
    B = n // 2 + 1
    cfg = lib.reference_config(n, hop, B, zoom=1e-3)
    with lib.BatchedSpectrum(cfg) as sp:
        got = sp.run(pcm, want_beat=False)
    ref = np.empty((1, frames, B))
    for f in range(frames):
        X = oracle.dft_naive(pcm[0, f * hop:f * hop + n].astype(np.float64))
        ref[0, f] = np.abs(X[:B]) * float(np.float32(1e-3))
    # "sparse" can produce a frame of pure digital silence next to a loud one: see assert_bars_close(pair_leak)
    assert_bars_close(got["bars"], ref, "n=%d %s" % (n, shape), pair_leak=PAIR_LEAK if shape == "sparse" else 0.0)
    assert_int_heights(got["bars_i32"], ref.astype(np.int64).astype(np.int32), ref, "n=%d %s" % (n, shape))


@pytest.mark.parametrize("n", [256, 512, 2048, 4096])
def test_silent_frame_next_to_a_loud_one(oracle, lib, n):
    """Digital silence in frame 2k, full-scale noise in frame 2k+1 (found by the random test above, seeds 1625994511 /
    1694285286 at n = 512 / 256): the silent frame's (int) heights are exactly 0 and its float heights at most
    PAIR_LEAK of the loud frame's maximum -- the rounding residue of the packed pair transform, not 0.0 as fftwf gives."""
    rng = np.random.default_rng(n)
    hop = n
    pcm = np.zeros((1, 4 * n), np.int16)
    pcm[0, n:2 * n] = rng.integers(-32768, 32768, n)      # frame 1 loud, frames 0, 2, 3 silent
    B = n // 2 + 1 if n <= 2048 else 2048
    cfg = lib.reference_config(n, hop, B, zoom=1e-3)
    with lib.BatchedSpectrum(cfg) as sp:
        got = sp.run(pcm, want_beat=False)
    loud = np.abs(got["bars"][0, 1]).max()
    assert loud > 100.0
    assert np.abs(got["bars"][0, 0]).max() <= PAIR_LEAK * loud
    assert (got["bars"][0, 2:] == 0.0).all()               # a pair of silent frames is exactly zero
    assert (got["bars_i32"][0, 0] == 0).all() and (got["bars_i32"][0, 2:] == 0).all()


@pytest.mark.parametrize("n", [512, 2048, 4096, 16384])
def test_white_noise_parseval(oracle, lib, n):
    """sum_k |X[k]|^2 = N * sum_n x[n]^2 (full band from the half spectrum the bars hold)."""
    if n // 2 + 1 > 4096:
        pytest.skip("bars are capped at FFTL_MAX_BARS: the full band of n = 16384 does not fit one bar per bin")
    rng = np.random.default_rng(n)
    pcm = rng.normal(0, 6000, (2, n + 7 * 512)).clip(-32768, 32767).astype(np.int16)
    cfg = lib.reference_config(n, 512, n // 2 + 1, zoom=1.0)
    with lib.BatchedSpectrum(cfg) as sp:
        bars = sp.run(pcm, want_beat=False, want_i32=False)["bars"].astype(np.float64)
    wgt = np.full(n // 2 + 1, 2.0)
    wgt[0] = wgt[-1] = 1.0
    for ch in range(2):
        for f in range(8):
            x = pcm[ch, f * 512:f * 512 + n].astype(np.float64)
            lhs = (wgt * bars[ch, f] ** 2).sum()
            assert abs(lhs - n * (x ** 2).sum()) <= 2e-6 * lhs, (ch, f)


@pytest.mark.parametrize("n,hop", [(2048, 512), (4096, 512)])
def test_am_bass_tone_gives_modulation_peak_on_the_gpu(oracle, lib, n, hop):
    """2 Hz amplitude modulation of a tone on the bass bin -> BassBeatDetector's peak index sits at
    2 Hz * 256 * hop / fs (beatDetector.c:51-106), once the 256-frame ring has filled."""
    fs = 48000
    tone = 4 * fs / n                                   # exactly on bin 4 = bass_bin
    t = np.arange(n + hop * 900) / fs
    x = 6000 * (0.5 + 0.5 * np.sin(2 * np.pi * 2.0 * t)) * np.sin(2 * np.pi * tone * t)
    pcm = np.round(x).astype(np.int16)[None, :]
    cfg = lib.reference_config(n, hop, 100)
    with lib.BatchedSpectrum(cfg) as sp:
        got = sp.run(pcm)
    k = got["beat"]["peak_index"][0, 300:]
    expect = 2.0 * 256 * hop / fs
    # the scan only looks above bin 30 for the highest peak: the 2 Hz line (bin ~5) is found as peaks[0] when nothing
    # above bin 30 qualifies, exactly as in the oracle's run of the same signal
    oc = _ocfg(oracle, cfg)
    ref = oracle.stft(oc, pcm)["beat"]["peak_index"][0, 300:]
    assert_peaks_explained(oracle, oc, got["beat"][0], "AM bass")
    assert (k == ref).mean() == 1.0 or abs(np.median(k) - np.median(ref)) <= 1
    assert abs(np.median(ref) - expect) <= 1.0 or np.median(ref) == 0
    # the ring itself carries the modulation: its transform peaks at the 2 Hz bin
    mag, _ = oracle.tempo_mag_replay(got["beat"]["bass"][0])
    assert abs(int(np.argmax(mag[600, 1:129])) + 1 - expect) <= 1.0


def test_exact_integer_magnitudes(oracle, lib):
    """A constant signal of value c: X[0] = N*c exactly, every other bin 0.  With bass_bin = 0 the tempo ring
    becomes constant, so heightBuffer[0] = 256*N*c exactly.  Integer boundaries are where --use_fast_math's
    approximate sqrt could truncate one low: the beat stages use the correctly rounded forms, the bar heights
    follow the +-1 rule at an exact boundary."""
    n, hop, c = 2048, 512, 5
    pcm = np.full((1, n + 400 * hop), c, np.int16)
    cfg = lib.reference_config(n, hop, 100, zoom=1.0)
    cfg.bass_bin = 0
    cfg.strict_int = 0
    with lib.BatchedSpectrum(cfg) as sp:
        got = sp.run(pcm)
    assert (np.abs(got["bars"][0, :, 0] - n * c) <= 1e-2).all() and (np.abs(got["bars"][0, :, 1:]) <= 1e-3).all()
    assert (np.abs(got["beat"]["bass"][0] - n * c) <= 1).all()
    oc = _ocfg(oracle, cfg)
    assert_peaks_explained(oracle, oc, got["beat"][0], "constant ring")
    # drop-in path: |(3, 4)| = 5 exactly, ring of 256 fives -> beatoutput[0] = 1280 exactly
    from fft_lines_b200 import beat_detector as bd, fft_calculate as fc
    bi, bo = np.zeros(256, np.complex64), np.zeros(256, np.complex64)
    fc.initializefft256(bi, bo)
    bd.reset()
    out = np.zeros(2048, np.complex64)
    out[4] = 3 + 4j
    bd.set_timeDelta(0.01)
    for _ in range(256):
        bd.BassBeatDetector(None, out, bi, bo)
    assert (bi.real == 5).all() and bo[0] == 1280 and np.abs(bo[1:]).max() == 0
    fc.cleanupfft256()


def test_capture_format_push_matches_int16_push(oracle, lib):
    """fftl_stream_push_device_f32: (int16_t)(x * 65536) as the reference's build computes it (wasapi_audio.cpp:398),
    including the C-undefined region |x| >= 0.5 (low 16 bits of the truncated int32), NaN and infinities (-> 0);
    interleaved stereo capture layout (sample stride 2)."""
    import torch
    n, B, streams, count = 2048, 100, 6, 480
    cfg = lib.reference_config(n, count, B)
    rng = np.random.default_rng(3)
    ticks = 8
    x = rng.normal(0, 0.12, (ticks, streams, count, 2)).astype(np.float32)   # [.., 0] = left, [.., 1] = right
    x[1, 0, :8, 0] = [0.5, -0.5, 0.75, -0.75, 1.0, -1.0, 0.49999997, 123.456]
    x[2, 1, :4, 0] = [np.nan, np.inf, -np.inf, 4e4]
    prod = x.astype(np.float32) * np.float32(65536)
    with np.errstate(invalid="ignore"):
        t = np.where(np.isfinite(prod) & (np.abs(prod) < 2.0 ** 31), np.trunc(prod), -2.0 ** 31).astype(np.int64)
    want16 = (t & 0xffff).astype(np.uint16).view(np.int16)
    assert want16[1, 0, 0, 0] == -32768 and want16[1, 0, 4, 0] == 0 and want16[2, 1, 0, 0] == 0
    with lib.StreamingSpectrum(cfg, streams) as a, lib.StreamingSpectrum(cfg, streams) as b:
        bars_a = torch.empty((streams, B), dtype=torch.float32, device="cuda")
        bars_b = torch.empty_like(bars_a)
        beat_a = torch.empty((streams, 8), dtype=torch.int32, device="cuda")
        beat_b = torch.empty_like(beat_a)
        for k in range(ticks):
            cap = torch.from_numpy(x[k]).cuda()                      # [streams, count, 2] interleaved
            a.push_device_f32(cap[:, :, 0], bars_a, None, beat_a)   # left channel: sample stride 2
            b.push_device(torch.from_numpy(np.ascontiguousarray(want16[k, :, :, 0])).cuda(), bars_b, None, beat_b)
            torch.cuda.synchronize()
            assert torch.equal(bars_a, bars_b) and torch.equal(beat_a, beat_b), k
        wf = a.waveform_device(points=n, scale=1.0, offset=0.0).cpu().numpy()
        assert (wf[:, -count:] == want16[ticks - 1, :, :, 0]).all()


def test_1024_stream_tick_against_the_oracle(oracle, lib):
    """BASELINE configs[2] at its real shape: 1024 mono streams, 1024-point window, 64 bars, one tick per batch.
    16 streams spread over the batch (both members of packed pairs, first and last) are checked against the oracle
    over 40 ticks: float bars, int heights, and the beat chain replayed on the GPU's own integers."""
    import torch
    n, B, streams, count, ticks = 1024, 64, 1024, 512, 40
    cfg = lib.extended_config(n, count, B)
    oc = _ocfg(oracle, cfg)
    spots = [0, 1, 2, 3, 62, 63, 64, 65, 510, 511, 512, 777, 1000, 1021, 1022, 1023]
    src = oracle.synth_pcm(streams, count * ticks)
    d_src = torch.from_numpy(src).cuda()
    bars = torch.empty((streams, B), dtype=torch.float32, device="cuda")
    i32 = torch.empty((streams, B), dtype=torch.int32, device="cuda")
    beat = torch.empty((streams, 8), dtype=torch.int32, device="cuda")
    win = np.zeros((len(spots), n), np.int16)
    series = []
    with lib.StreamingSpectrum(cfg, streams) as ss:
        for k in range(ticks):
            ss.push_device(d_src[:, k * count:(k + 1) * count], bars, i32, beat)
            torch.cuda.synchronize()
            hb, hi, bt = bars.cpu().numpy(), i32.cpu().numpy(), lib.BatchedSpectrum.beat_to_numpy(beat)
            ref_h = np.zeros((len(spots), B))
            ref_i = np.zeros((len(spots), B), np.int32)
            for q, s in enumerate(spots):
                win[q] = np.concatenate([win[q, count:], src[s, k * count:(k + 1) * count]])
                _, ref_h[q], ref_i[q] = oracle.frame(oc, win[q])
            assert_bars_close(hb[spots], ref_h, "tick %d" % k)
            assert_int_heights(hi[spots], ref_i, ref_h, "tick %d" % k)
            series.append(bt[spots].copy())
        assert ss.fused_tick_count == ticks and ss.launch_count == ticks   # the one-kernel tick (stream_tick.cuh) ran
    for q, s in enumerate(spots):
        one = np.array([series[k][q] for k in range(ticks)], dtype=series[0].dtype)
        assert_peaks_explained(oracle, oc, one, "stream %d" % s)
