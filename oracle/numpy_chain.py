"""The whole chain of BASELINE configs[1] on library FFTs (scipy.fft = pocketfft, torch.fft = MKL/pocketfft),
vectorised over frames and run on all host cores.

TEST INFRASTRUCTURE / CPU BASELINE ONLY (bench.py's cpu_baseline leg and tests/): a second, honest CPU number
beside the scalar port, because the port's FFT is this repo's own radix-4 code and a tuned host FFT is several
times faster.  Every stage follows the oracle (oracle/fftl_oracle.c), which follows the reference:
frames (fft_lines.c:190-194) -> window -> real FFT -> |X| (fft_lines.c:202) -> bin means -> Savitzky-Golay ->
zoom -> (int) heights -> band energy (fft_lines.c:281) -> AddBeatSample (beatDetector.c:30-44) ->
BassBeatDetector (beatDetector.c:51-106: ring transform per frame, integer heights, peak scan).
"""
import numpy as np

BEAT_SAMPLES, TEMPO_RING = 160, 256


def _tables(O, cfg):
    lo, hi = O.build_bins(cfg)
    win = O.window(cfg).astype(np.float32)
    sg = O.build_sg(cfg.sg_half, cfg.sg_degree).astype(np.float32) if cfg.sg_half > 0 else None
    return lo, hi, win, sg


def _peak_scan(hb):
    """beatDetector.c:78-103 for all frames at once.  hb: [F, 161] int64."""
    dnew = hb[:, 1:] - hb[:, :-1]                         # [F, 160]
    dold = np.concatenate([np.zeros((hb.shape[0], 1), hb.dtype), dnew[:, :-1]], axis=1)
    pk = (dold >= 0) & (dnew <= 0)
    rank = np.cumsum(pk, axis=1) - 1
    idx = np.arange(160)[None, :]
    cand = pk & (rank < 30) & (idx > 30) & (hb[:, :160] > 0)
    masked = np.where(cand, hb[:, :160], -1)
    best = masked.argmax(axis=1)                          # first of equals, like the strict '>' of the scan
    first = np.where(pk.any(axis=1), pk.argmax(axis=1), 0)
    return np.where(masked.max(axis=1) > 0, best, first)


def run(O, cfg, pcm, provider="scipy", workers=None, block=512):
    """pcm int16 [channels, samples].  Returns dict(bars float32 [C,F,B], w, thr, beat_value, flag, bass, peak_index).
    provider: "scipy" (scipy.fft = pocketfft) or "torch" (torch.fft.rfft on the CPU).  Blocks of `block` frames are
    processed by `workers` threads (default: all cores), each running single-threaded library calls (numpy, scipy
    and torch release the GIL inside them), so every stage -- not only the FFT -- uses all the cores."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    n, hop, B = cfg.n, cfg.hop, cfg.bars
    lo, hi, win, sg = _tables(O, cfg)
    ch, S = pcm.shape
    F = (S - n) // hop + 1
    if not workers or workers < 1:
        workers = len(os.sched_getaffinity(0))
    bars = np.empty((ch, F, B), np.float32)
    bass = np.empty((ch, F), np.int64)
    w = np.empty((ch, F), np.int64)
    peak = np.empty((ch, F), np.int64)
    inv = (1.0 / (hi - lo)).astype(np.float32)
    m, wd = 2 * cfg.sg_half + 1, cfg.sg_half
    bb = list(cfg.beat_bars)
    zoom = np.float32(cfg.zoom)
    if provider == "torch":
        import torch
        torch.set_num_threads(1)
        twin = torch.from_numpy(win)

        def rfft_mag(x, windowed):
            t = torch.from_numpy(np.ascontiguousarray(x))
            t = t.float() * twin if windowed else t
            return torch.fft.rfft(t, dim=1).abs().numpy()
    else:
        import scipy.fft as sfft

        def rfft_mag(x, windowed):
            return np.abs(sfft.rfft(x.astype(np.float32) * win if windowed else x, axis=1))

    def bars_block(job):
        c, f0 = job
        frames = np.lib.stride_tricks.as_strided(pcm[c], shape=(F, n), strides=(hop * 2, 2), writeable=False)
        mag = rfft_mag(frames[f0:f0 + block], True)
        # bin means through a prefix sum in float64 (one vectorised pass; the oracle sums directly)
        cs = np.concatenate([np.zeros((mag.shape[0], 1)), np.cumsum(mag, axis=1, dtype=np.float64)], axis=1)
        raw = ((cs[:, hi] - cs[:, lo]) * inv).astype(np.float32)
        if sg is not None:
            # interior bars: centre row as a FIR; the w bars at either edge: the edge rows of the fit (mode="interp")
            y = np.empty_like(raw)
            acc = np.zeros((raw.shape[0], B - 2 * wd), np.float32)
            for k in range(m):
                acc += sg[wd, k] * raw[:, k:k + B - 2 * wd]
            y[:, wd:B - wd] = acc
            y[:, :wd] = raw[:, :m] @ sg[:wd].T
            y[:, B - wd:] = raw[:, B - m:] @ sg[wd + 1:].T
            raw = y
        h = raw * zoom
        bars[c, f0:f0 + block] = h + (10.0 if cfg.circle else 0.0)
        hi32 = h.astype(np.int64) + (10 if cfg.circle else 0)
        w[c, f0:f0 + block] = ((hi32[:, bb[0]] + hi32[:, bb[1]]) + (hi32[:, bb[2]] + hi32[:, bb[3]])) // 4
        bass[c, f0:f0 + block] = mag[:, cfg.bass_bin].astype(np.int64)

    def tempo_block(job):
        # BassBeatDetector: the ring after frame f holds frames f-255..f (magnitudes do not depend on its rotation)
        c, f0 = job
        Y = rfft_mag(rings[c][f0:f0 + block], False)
        full = np.concatenate([Y, Y[:, 127:95:-1]], axis=1)                        # bins 0..160 (129..160 mirror 127..96)
        peak[c, f0:f0 + block] = _peak_scan(full.astype(np.int64))

    jobs = [(c, f0) for c in range(ch) for f0 in range(0, F, block)]
    with ThreadPoolExecutor(max_workers=workers) as pool:
        list(pool.map(bars_block, jobs))
        rings = [np.lib.stride_tricks.sliding_window_view(
            np.concatenate([np.zeros(TEMPO_RING - 1, np.float32), bass[c].astype(np.float32)]), TEMPO_RING) for c in range(ch)]
        list(pool.map(tempo_block, jobs))
    # AddBeatSample: running 160-sum (beatDetector.c:32-43)
    csw = np.concatenate([np.zeros((ch, BEAT_SAMPLES), np.int64), np.cumsum(w, axis=1)], axis=1)
    ssum = csw[:, BEAT_SAMPLES:] - csw[:, :-BEAT_SAMPLES]
    thr = np.trunc(ssum / BEAT_SAMPLES).astype(np.int64)
    with np.errstate(divide="ignore", invalid="ignore"):
        q = (w.astype(np.float32) / thr.astype(np.float32)) * np.float32(128)
        bv = np.where(np.isfinite(q) & (np.abs(q) < 2.0 ** 31), np.trunc(q), -2.0 ** 31).astype(np.int64)
    return {"bars": bars, "w": w, "thr": thr, "beat_value": bv, "flag": ((thr > 0) & (w > thr)).astype(np.int64),
            "bass": bass, "peak_index": peak, "workers": workers}


def fft_only(cfg, pcm, provider="scipy", workers=None, block=512):
    """Only the windowed real FFTs of every frame (no magnitudes, bins, smoothing or beat stages), blocks of frames over
    all cores: an upper bound for ANY CPU chain built on this FFT.  Returns (seconds, frames, workers)."""
    import os
    import time
    from concurrent.futures import ThreadPoolExecutor
    n, hop = cfg.n, cfg.hop
    ch, S = pcm.shape
    F = (S - n) // hop + 1
    if not workers or workers < 1:
        workers = len(os.sched_getaffinity(0))
    win = (0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n) / n)).astype(np.float32)
    if provider == "torch":
        import torch
        torch.set_num_threads(1)
        twin = torch.from_numpy(win)

        def one(x):
            return torch.fft.rfft(torch.from_numpy(np.ascontiguousarray(x)).float() * twin, dim=1)
    else:
        import scipy.fft as sfft

        def one(x):
            return sfft.rfft(x.astype(np.float32) * win, axis=1)

    def job(cf):
        c, f0 = cf
        frames = np.lib.stride_tricks.as_strided(pcm[c], shape=(F, n), strides=(hop * 2, 2), writeable=False)
        one(frames[f0:f0 + block])

    jobs = [(c, f0) for c in range(ch) for f0 in range(0, F, block)]
    with ThreadPoolExecutor(max_workers=workers) as pool:
        list(pool.map(job, jobs[:workers]))  # warm-up
        t0 = time.perf_counter()
        list(pool.map(job, jobs))
        dt = time.perf_counter() - t0
    return dt, ch * F, workers
