"""Full-size (BASELINE.json configs[1]) checks through size-independent
properties: shard == full run, channel independence, time-shift consistency,
linearity of the pre-magnitude stages via scaling, and spot frames against the
oracle."""
import ctypes as C

import numpy as np
import pytest

from helpers import assert_peaks_explained

pytestmark = pytest.mark.gpu


def test_c2_full_hour_properties(oracle, lib):
    import torch
    cfg = lib.extended_config(4096, 512, 200)
    S = 48000 * 3600
    pcm = torch.empty((2, S), dtype=torch.int16, device="cuda")
    lib.synth_pcm_device(pcm)
    with lib.BatchedSpectrum(cfg) as sp:
        nf = sp.num_frames(S)
        assert nf == 337493                                  # F per channel for c2
        bars, _, beat = sp.alloc_device_outputs(2, nf, "cuda")
        sp.run_device(pcm, bars, None, beat)
        torch.cuda.synchronize()
        assert torch.isfinite(bars).all()
        # (1) a frame-range shard reproduces the same numbers, beat included (halo recompute)
        b0, b1 = 200000, 200000 + 4099
        pb, _, pbeat = sp.alloc_device_outputs(2, b1 - b0, "cuda")
        sp.run_device(pcm, pb, None, pbeat, b0, b1)
        torch.cuda.synchronize()
        assert torch.equal(pb, bars[:, b0:b1]) and torch.equal(pbeat, beat[:, b0:b1])
        # (2) channels are independent: running channel 1 alone gives the same bars
        ob, _, _ = sp.alloc_device_outputs(1, 5000, "cuda", want_beat=False)
        sp.run_device(pcm[1:2], ob, None, None, 1000, 6000)
        torch.cuda.synchronize()
        assert torch.equal(ob[0], bars[1, 1000:6000])
        # (3) time shift by two hops: frame f of the shifted signal == frame f+2 (frames f, f^1 share a
        #     transform, so an even shift keeps the arithmetic identical); an odd shift agrees to rounding
        sb, _, _ = sp.alloc_device_outputs(2, 3000, "cuda", want_beat=False)
        sp.run_device(pcm[:, 1024:], sb, None, None, 0, 3000)
        torch.cuda.synchronize()
        assert torch.equal(sb, bars[:, 2:3002])
        sp.run_device(pcm[:, 512:], sb, None, None, 0, 3000)
        torch.cuda.synchronize()
        ref1 = bars[:, 1:3001]
        assert ((sb - ref1).abs() <= 2e-6 * ref1.abs().amax(dim=-1, keepdim=True)).all()
        # (4) spot frames against the oracle (the only O(frames) CPU work here)
        ocfg = oracle.Config()
        C.memmove(C.byref(ocfg), C.byref(cfg), C.sizeof(ocfg))
        for f in (0, 1, 168746, 337491, 337492):
            seg = pcm[:, f * 512:f * 512 + 4096].cpu().numpy()
            ref = oracle.stft(ocfg, seg, want_beat=False)["bars_f64"][:, 0]
            got = bars[:, f].cpu().numpy().astype(np.float64)
            assert (np.abs(got - ref) <= 1e-4 * np.abs(ref).max(axis=-1, keepdims=True)).all(), f
        # (5) beat chain over the whole hour replayed bit-exactly from the GPU's own integers
        b = lib.BatchedSpectrum.beat_to_numpy(beat)
        #     (a differing tempo peak index must be explained as an integer-boundary flip: helpers.assert_peaks_explained)
        for ch in range(2):
            assert_peaks_explained(oracle, ocfg, b[ch, :40000], "c2 channel %d" % ch)
        assert b["flag"].sum() > 0


def test_scaling_property(oracle, lib):
    """Pre-magnitude stages are linear and |.| is homogeneous: halving the PCM halves the float bars."""
    pcm = (oracle.synth_pcm(1, 4096 + 512 * 64) // 2 * 2).astype(np.int16)
    half = (pcm // 2).astype(np.int16)
    cfg = lib.extended_config(4096, 512, 200)
    with lib.BatchedSpectrum(cfg) as sp:
        a = sp.run(pcm, want_beat=False)["bars"]
        b = sp.run(half, want_beat=False)["bars"]
    fmax = np.abs(a).max(axis=-1, keepdims=True)
    assert (np.abs(a - 2 * b) <= 2e-6 * fmax).all()
