"""Shared comparison rules for the parity tests."""
import numpy as np

REL_TOL = 1e-4  # north_star: bar heights within 1e-4 relative to the frame maximum


PAIR_LEAK = 1e-6  # rounding residue of the louder frame of a pair (2k, 2k+1) that may show in the quieter one (see below)


def assert_bars_close(gpu, ref_f64, what="", pair_leak=0.0):
    """float heights: |gpu - oracle| <= 1e-4 * max(|oracle frame|).

    pair_leak > 0 (tests that feed digital silence): frames 2k and 2k+1 share packed transforms,
    whose untangling X_a = (Z[m] + conj Z[M-m]) / 2 cancels the partner's spectrum only up to its fp32 rounding error
    (measured 2e-7 of the partner's maximum), so a frame that is EXACTLY zero next to a loud one comes out as ~1e-7 of
    its neighbour instead of 0.0; the (int) heights are 0 either way.  The bound then is
    1e-4 * own frame maximum + pair_leak * maximum of the frame pair."""
    gpu = np.asarray(gpu, np.float64)
    ref = np.asarray(ref_f64, np.float64)
    assert gpu.shape == ref.shape, (gpu.shape, ref.shape)
    fmax = np.abs(ref).max(axis=-1, keepdims=True)
    tol = REL_TOL * np.maximum(fmax, 1e-30)
    if pair_leak:
        nf = ref.shape[-2]
        partner = np.minimum(np.arange(nf) ^ 1, nf - 1)
        tol = tol + pair_leak * np.maximum(fmax, np.take(fmax, partner, axis=-2))
    err = np.abs(gpu - ref)
    worst = (err / tol).max() if err.size else 0.0
    assert worst <= 1.0, "%s: worst error %.3g x tolerance (abs %.3g, frame max %.3g)" % (
        what, worst, err.max(), fmax.max())
    return worst


def assert_int_heights(gpu_i32, ref_i32, ref_f64, what=""):
    """int heights: |d| <= 1, and d != 0 only where the float height is within
    1e-4*frame max of an integer boundary."""
    d = gpu_i32.astype(np.int64) - ref_i32.astype(np.int64)
    assert np.abs(d).max(initial=0) <= 1, "%s: int height off by more than 1" % what
    bad = d != 0
    if bad.any():
        fmax = np.abs(ref_f64).max(axis=-1, keepdims=True)
        tol = np.broadcast_to(REL_TOL * fmax, ref_f64.shape)
        dist = np.abs(ref_f64 - np.rint(ref_f64))
        assert (dist[bad] <= tol[bad] + 1e-12).all(), "%s: int mismatch away from an integer boundary" % what
    return float(bad.mean()) if bad.size else 0.0


# ---------------------------------------------------------------------------------------------------------------
# Tempo peak index (beatDetector.c:78-103).  The scan compares NEIGHBOURING integer heights of a 256-point fp32
# transform, so two correct transforms can disagree on a frame: a height that sits next to an integer boundary, or two
# neighbours that are (almost) equal, flips a `diffold >= 0 && diffnew <= 0` decision, and a peak gained or lost early
# shifts the 30-slot peak list.  Every frame where the GPU's index differs from the oracle's must be EXPLAINED that way:
# with every height allowed to move by the rounding error of one fp32 transform (TEMPO_REL of the ring's norm, + 1
# count for the cast) the GPU's index has to be a possible outcome of the reference's scan.  Anything else fails.
TEMPO_REL = 1e-6   # fp32 256-point FFT: per-bin error <~ 3e-7 * ||ring||_2 (measured); 1e-6 leaves a factor 3


def _peak_possible(mag, tau, p, oob=False):
    """Is index p a possible result of the reference's scan when hb[i] = (int)x for some |x - mag[i]| <= tau?
    oob: heightBuffer[160] is whatever the reference's out-of-bounds read finds (beatDetector.c:87)."""
    lo = np.floor(np.maximum(mag - tau, 0.0))
    hi = np.floor(mag + tau)
    if hi.max() >= 2.0 ** 31:          # (int) overflow region (strict_int): treat as unknown
        return True
    if oob:
        lo[160], hi[160] = -2.0 ** 31, 2.0 ** 31
    dn_lo = lo[1:] - hi[:-1]           # diffnew of bin i = hb[i+1] - hb[i], i = 0..159
    dn_hi = hi[1:] - lo[:-1]
    do_lo = np.concatenate(([0.0], dn_lo[:-1]))
    do_hi = np.concatenate(([0.0], dn_hi[:-1]))
    possible = (do_hi >= 0) & (dn_lo <= 0)
    definite = (do_lo >= 0) & (dn_hi <= 0)
    idx = np.arange(160)
    min_rank = np.concatenate(([0], np.cumsum(definite)[:-1]))   # definite peaks before i
    max_rank = np.concatenate(([0], np.cumsum(possible)[:-1]))   # possible peaks before i
    sure = definite & (max_rank < 30) & (idx > 30) & (lo[:160] > 0)  # certainly a candidate for the highest peak
    if p > 30:
        if not (possible[p] and min_rank[p] < 30 and hi[p] > 0):
            return False
        return not (sure & (lo[:160] > hi[p]) & (idx != p)).any()   # nobody is certainly higher
    # p <= 30: no peak above bin 30 qualified, the reference returns peaks[0] (or 0 when there is no peak at all)
    if sure.any():
        return False
    if p == 0:
        return bool(possible[0]) or not definite.any()
    return bool(possible[p]) and not definite[:p].any()


def assert_peaks_explained(O, ocfg, gpu_beat, what="", n_dt=None):
    """gpu_beat: one channel's beat records from frame 0.  The integer stages replayed by the oracle on the GPU's own
    (w, bass) must be bit-exact; the peak index may differ only on frames where the difference is explained (above),
    and movement_per_sec must agree wherever the index does.  Returns (#frames, #differing, all explained)."""
    rep = O.beat_replay(ocfg, gpu_beat["w"], gpu_beat["bass"], first_frame=0)
    for k in ("w", "thr", "beat_value", "flag", "bass"):
        assert (rep[k] == gpu_beat[k]).all(), "%s: %s differs from the oracle's replay" % (what, k)
    same = rep["peak_index"] == gpu_beat["peak_index"]
    if n_dt is None:
        assert (rep["movement_per_sec"][same] == gpu_beat["movement_per_sec"][same]).all(), "%s: movement_per_sec" % what
    # beatDetector.c:105 on the GPU's own index (left to right in float); n_dt: per-frame N*timeDelta when the hop varies
    if n_dt is None:
        n_dt = np.float32(np.float32(ocfg.n) * (np.float32(ocfg.hop) / np.float32(ocfg.sample_rate)))
    assert (gpu_beat["movement_per_sec"] == np.asarray(n_dt, np.float32) * gpu_beat["peak_index"].astype(np.float32)).all(), "%s: movement" % what
    diff = np.nonzero(~same)[0]
    if diff.size:
        mag, norm = O.tempo_mag_replay(gpu_beat["bass"])
        bad = [int(f) for f in diff if not _peak_possible(mag[f], 1.0 + TEMPO_REL * norm[f], int(gpu_beat["peak_index"][f]))]
        assert not bad, "%s: %d of %d differing peak indices are NOT explained by an integer-boundary flip, first frames %s" % (
            what, len(bad), diff.size, bad[:8])
    return same.size, int(diff.size)
