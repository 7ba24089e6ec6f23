"""Shared comparison rules for the parity tests."""
import numpy as np

REL_TOL = 1e-4  # north_star: bar heights within 1e-4 relative to the frame maximum


def assert_bars_close(gpu, ref_f64, what=""):
    """float heights: |gpu - oracle| <= 1e-4 * max(|oracle frame|)."""
    gpu = np.asarray(gpu, np.float64)
    ref = np.asarray(ref_f64, np.float64)
    assert gpu.shape == ref.shape, (gpu.shape, ref.shape)
    fmax = np.abs(ref).max(axis=-1, keepdims=True)
    tol = REL_TOL * np.maximum(fmax, 1e-30)
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
