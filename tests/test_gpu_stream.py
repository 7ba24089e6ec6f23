"""GPU tests of the many-stream real-time front end (fftl_stream_*): the sliding window of
GetAudioBuffer/MoveArray (reference wasapi_audio.cpp:390-484) + one WM_TIMER tick per push."""
import ctypes as C

import numpy as np
import pytest

from helpers import assert_bars_close, assert_int_heights, assert_peaks_explained

pytestmark = pytest.mark.gpu


def _ocfg(O, cfg):
    oc = O.Config()
    C.memmove(C.byref(oc), C.byref(cfg), C.sizeof(oc))
    return oc


def move_array_reference(dest, fresh):
    """MoveArray as written in the reference (wasapi_audio.cpp:390-400), int16 input instead of float*65536:
    the old samples are copied from k + addCount - 1, i.e. shifted by one less than addCount."""
    n, add = dest.shape[0], fresh.shape[0]
    out = dest.copy()
    for k in range(n - add):
        out[k] = out[k + add - 1]
    out[n - add:] = fresh
    return out


def move_array_exact(dest, fresh):
    n, add = dest.shape[0], fresh.shape[0]
    if add >= n:
        return fresh[add - n:].copy()
    return np.concatenate([dest[add:], fresh])


@pytest.mark.parametrize("mode,n,tick,bars", [("exact", 1024, "one_kernel", 64), ("exact", 2048, "one_kernel", 64), ("exact", 1024, "three_launches", 64),
                                              ("exact", 4096, "three_launches", 64), ("reference", 1024, "three_launches", 100),
                                              ("exact", 1024, "one_kernel", 200),      # fewer lanes than 2 per bar: one lane per bar
                                              ("exact_parity", 1024, "one_kernel", 400),  # more bars than threads, no window, no SG, bar i = bin i
                                              ("exact_parity", 2048, "one_kernel", 100), ("exact_circle", 1024, "one_kernel", 32)])
def test_streaming_ticks_match_oracle(oracle, lib, mode, n, tick, bars):
    """Every tick of every stream against the oracle, through both forms of a push: the one-kernel tick (stream_tick.cuh:
    n = 1024 / 2048 with the exact window move) and the three launches (window move, batched bars kernel, beat stage)."""
    B, streams = bars, 7
    cfg = lib.extended_config(n, 1024, B) if mode in ("exact", "exact_circle") else lib.reference_config(n, 1024, B)
    if mode == "exact_circle":
        cfg.circle = 1                                # heights + 10 (fft_lines.c:204-207)
    if mode != "reference":
        mode = "exact"
    B = cfg.bars
    oc = _ocfg(oracle, cfg)
    rng = np.random.default_rng(11)
    counts = [n, 512, 480, 1, 777, n, 300] + [int(c) for c in rng.integers(200, n + 1, 40)]
    if mode == "exact":
        counts[3] = n + 476                      # more new samples than the window holds
    total = sum(counts)
    src = oracle.synth_pcm(streams, total)
    win = np.zeros((streams, n), np.int16)
    mover = move_array_exact if mode == "exact" else move_array_reference
    got_w = np.zeros((streams, len(counts)), np.int32)
    got_bass = np.zeros_like(got_w)
    got_beat = []
    pos = 0
    with lib.StreamingSpectrum(cfg, streams, lib.MOVE_EXACT if mode == "exact" else lib.MOVE_REFERENCE) as ss:
        if tick == "three_launches":
            ss.set_tick_kernel(lib.TICK_THREE_LAUNCHES)
        for k, c in enumerate(counts):
            fresh = src[:, pos:pos + c]
            pos += c
            got = ss.push(fresh)
            ref_h = np.zeros((streams, B))
            ref_i = np.zeros((streams, B), np.int32)
            for s in range(streams):
                win[s] = mover(win[s], fresh[s])
                _, ref_h[s], ref_i[s] = oracle.frame(oc, win[s])
            assert_bars_close(got["bars"], ref_h, "tick %d" % k)
            assert_int_heights(got["bars_i32"], ref_i, ref_h, "tick %d" % k)
            got_w[:, k] = got["beat"]["w"]
            got_bass[:, k] = got["beat"]["bass"]
            got_beat.append(got["beat"].copy())
        assert ss.launch_count == (1 if tick == "one_kernel" else 3) * len(counts)
        assert ss.fused_tick_count == (len(counts) if tick == "one_kernel" else 0)
    # integer stages replayed by the oracle on the GPU's own (w, bass) series: bit exact
    # (peak index: equal, or explained as an integer-boundary flip; movement = N * timeDelta * index with this tick's count)
    n_dt = np.array([np.float32(n) * (np.float32(c) / np.float32(cfg.sample_rate)) for c in counts], np.float32)
    for s in range(streams):
        series = np.array([got_beat[k][s] for k in range(len(counts))], dtype=got_beat[0].dtype)
        assert_peaks_explained(oracle, oc, series, "stream %d" % s, n_dt=n_dt)


def test_streaming_equals_batched_stft_for_constant_hop(oracle, lib):
    """With a constant count the stream is an STFT with hop = count over [zeros(n), samples]."""
    n, hop, B, streams, ticks = 2048, 512, 100, 4, 60
    cfg = lib.reference_config(n, hop, B)
    src = oracle.synth_pcm(streams, hop * ticks)
    padded = np.concatenate([np.zeros((streams, n), np.int16), src], axis=1)
    with lib.BatchedSpectrum(cfg) as sp:
        full = sp.run(padded)               # frame f+1 = window after push f
    with lib.StreamingSpectrum(cfg, streams) as ss:
        for k in range(ticks):
            got = ss.push(src[:, k * hop:(k + 1) * hop])
            ref = full["bars"][:, k + 1]
            fmax = np.abs(ref).max(axis=-1, keepdims=True)
            assert (np.abs(got["bars"] - ref) <= 2e-6 * fmax + 1e-12).all()   # different frame pairing: rounding only
        ss.reset()
        again = ss.push(src[:, :hop])
        assert (np.abs(again["bars"] - full["bars"][:, 1]) <= 2e-6 * np.abs(full["bars"][:, 1]).max()).all()
        assert (again["beat"]["thr"] == again["beat"]["w"] // 160).all()       # rings were zeroed


def test_streaming_cuda_graph_replay(oracle, lib):
    import torch
    n, B, streams, count = 1024, 64, 1024, 256
    cfg = lib.extended_config(n, count, B)
    src = torch.from_numpy(oracle.synth_pcm(streams, count * 12)).cuda()
    fresh = torch.empty((streams, count), dtype=torch.int16, device="cuda")
    bars = torch.empty((streams, B), dtype=torch.float32, device="cuda")
    beat = torch.empty((streams, 8), dtype=torch.int32, device="cuda")
    ref_bars, ref_beat = [], []
    with lib.StreamingSpectrum(cfg, streams) as ss:
        for k in range(12):
            fresh.copy_(src[:, k * count:(k + 1) * count])
            ss.push_device(fresh, bars, None, beat)
            torch.cuda.synchronize()
            ref_bars.append(bars.clone()); ref_beat.append(beat.clone())
    with lib.StreamingSpectrum(cfg, streams) as ss:
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            fresh.copy_(src[:, :count])
            ss.push_device(fresh, bars, None, beat, stream=st)          # tick 0 outside the graph (warm-up)
            st.synchronize()
            assert torch.equal(bars, ref_bars[0])
            g = torch.cuda.CUDAGraph()
            fresh.copy_(src[:, count:2 * count])
            with torch.cuda.graph(g, stream=st):
                ss.push_device(fresh, bars, None, beat, stream=st)
            for k in range(1, 12):                                      # capture did not execute: replay ticks 1..11
                fresh.copy_(src[:, k * count:(k + 1) * count])
                g.replay()
                st.synchronize()
                assert torch.equal(bars, ref_bars[k]), k
                assert torch.equal(beat, ref_beat[k]), k


def test_history_ring_and_waveform(oracle, lib):
    import torch
    n, B, streams, hop, depth = 2048, 100, 5, 512, 6
    cfg = lib.reference_config(n, hop, B)
    src = oracle.synth_pcm(streams, hop * 10)
    frames = []
    with lib.StreamingSpectrum(cfg, streams) as ss:
        ss.set_history(depth)
        for k in range(10):
            frames.append(ss.push(src[:, k * hop:(k + 1) * hop])["bars"].copy())
            hist = ss.history_device().cpu().numpy()
            for a in range(depth):                      # k = depth-1 is the newest frame
                want = frames[k - a] if k - a >= 0 else np.zeros((streams, B), np.float32)
                assert (hist[:, depth - 1 - a] == want).all(), (k, a)
        # waveform view of the current window: (int)(sample * 0.02f + offset), fft_lines.c:245
        win = np.concatenate([np.zeros((streams, n), np.int16), src], axis=1)[:, -n:]
        off = np.float32((1080 - 30) // 2 + 50)
        wf = ss.waveform_device(points=2048, scale=0.02, offset=float(off)).cpu().numpy()
        want = (win.astype(np.float32) * np.float32(0.02) + off).astype(np.int32)
        assert (wf == want).all()


def test_streaming_tempo_transform_carried_across_ring_wraps(oracle, lib):
    """The streaming beat stage carries the 256-point transform of the tempo ring from tick to tick
    (stream_beat_carry_kernel): X[k] += (new - old) * W256^(pos k).  700 ticks wrap the 256-slot ring almost three times,
    so old values are non-zero and subtracted; a reset in the middle must restart from the zero ring.  The integer
    stages are replayed by the oracle on the GPU's own (w, bass) series (bit exact), every tempo index is equal to the
    fp64 oracle's or an explained integer-boundary flip (helpers.assert_peaks_explained)."""
    n, B, streams, count, ticks = 1024, 32, 5, 128, 700
    cfg = lib.extended_config(n, count, B)
    oc = oracle.Config()
    import ctypes as C
    C.memmove(C.byref(oc), C.byref(cfg), C.sizeof(oc))
    rng = np.random.default_rng(5)
    # loud, strongly varying material so that the bass bin (and with it the tempo ring) moves a lot
    env = (1.0 + np.sin(np.arange(ticks * count) * 2 * np.pi / (count * 37.0))) * 0.5
    src = (rng.normal(0, 9000, (streams, ticks * count)) * env + 6000 * np.sin(np.arange(ticks * count) * 2 * np.pi * 90.0 / 48000.0) * env)
    src = src.clip(-32768, 32767).astype(np.int16)
    recs = []
    with lib.StreamingSpectrum(cfg, streams) as ss:
        for k in range(ticks):
            recs.append(ss.push(src[:, k * count:(k + 1) * count], want_i32=False)["beat"].copy())
        n_dt = np.full(ticks, np.float32(n) * (np.float32(count) / np.float32(cfg.sample_rate)), np.float32)
        for s in range(streams):
            series = np.array([recs[k][s] for k in range(ticks)], dtype=recs[0].dtype)
            assert len(set(series["bass"][300:].tolist())) > 50       # the ring really changes
            assert_peaks_explained(oracle, oc, series, "stream %d" % s, n_dt=n_dt)
        # after a reset the carried transform starts again from the zero ring: the same input gives the same records
        ss.reset()
        again = [ss.push(src[:, k * count:(k + 1) * count], want_i32=False)["beat"].copy() for k in range(300)]
        for k in range(300):
            for name in recs[0].dtype.names:
                assert (again[k][name] == recs[k][name]).all(), (k, name)
