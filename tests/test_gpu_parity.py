"""GPU parity: libfftlines_cuda (through its C-ABI) against the CPU oracle.

Float stages are held to north_star's tolerance (1e-4 of the frame maximum);
integer stages are replayed by the oracle on the GPU's own integers and must
then be bit-exact (threshold, beatValue, flag, peak index, movement).
"""
import numpy as np
import pytest

from fft_lines_b200 import _capi
from helpers import assert_bars_close, assert_int_heights, assert_peaks_explained

pytestmark = pytest.mark.gpu


def _ocfg(O, F, cfg):
    """Same config for the oracle (independent struct type, same layout)."""
    import ctypes as C
    oc = O.Config()
    C.memmove(C.byref(oc), C.byref(cfg), C.sizeof(oc))
    return oc


def _check(O, F, cfg, pcm, frame_begin=0, frame_end=None, interleaved=False):
    oc = _ocfg(O, F, cfg)
    ref = O.stft(oc, pcm, frame_begin, frame_end)
    with F.BatchedSpectrum(cfg) as sp:
        got = sp.run(pcm.T.copy() if interleaved else pcm, frame_begin, frame_end, interleaved=interleaved)
    assert_bars_close(got["bars"], ref["bars_f64"], "bars")
    assert_int_heights(got["bars_i32"], ref["bars_i32"], ref["bars_f64"], "bars_i32")
    gb, rb = got["beat"], ref["beat"]
    # float-derived integers: w and bass may differ by 1 at integer boundaries
    assert np.abs(gb["w"].astype(np.int64) - rb["w"]).max(initial=0) <= 1
    scale = np.maximum(np.abs(rb["bass"]).max(initial=1), 1)
    assert np.abs(gb["bass"].astype(np.int64) - rb["bass"]).max(initial=0) <= max(1, 2e-6 * scale)
    # integer stages replayed on the GPU's own (w, bass): must be bit exact,
    # except peak_index which also depends on the fp32 256-point transform
    # (helpers.assert_peaks_explained: a differing index must be an integer-boundary flip of the reference's scan)
    if frame_begin == 0:
        for ch in range(pcm.shape[0]):
            assert_peaks_explained(O, oc, gb[ch], "channel %d" % ch)
    return got, ref


def test_parity_mode_c1(oracle, lib):
    """configs[0]: 2048-pt window, 100 bars, stereo sweep+noise (1 s of it here)."""
    pcm = oracle.synth_pcm(2, 48000)
    _check(oracle, lib, lib.reference_config(2048, 512, 100), pcm)


@pytest.mark.parametrize("n", [256, 512, 1024, 2048, 4096, 8192, 16384])
def test_all_sizes_reference_mode(oracle, lib, n):
    pcm = oracle.synth_pcm(2, n + 40 * 256)
    _check(oracle, lib, lib.reference_config(n, 256, min(100, n // 2 + 1)), pcm)


@pytest.mark.parametrize("n,bars", [(1024, 64), (4096, 200), (16384, 512)])
def test_extended_mode(oracle, lib, n, bars):
    """Hann + log bins + Savitzky-Golay (oracle-defined semantics)."""
    pcm = oracle.synth_pcm(2, n + 60 * 512)
    _check(oracle, lib, lib.extended_config(n, 512, bars), pcm)


def test_frame_range_shard_equals_full_run(oracle, lib):
    """Beat outputs of a shard [f0,f1) equal those of a run from frame 0."""
    cfg = lib.extended_config(4096, 512, 200)
    pcm = oracle.synth_pcm(2, 4096 + 700 * 512)
    with lib.BatchedSpectrum(cfg) as sp:
        full = sp.run(pcm)
        part = sp.run(pcm, 301, 650)
    assert (part["bars"] == full["bars"][:, 301:650]).all()
    assert (part["bars_i32"] == full["bars_i32"][:, 301:650]).all()
    for k in ("w", "thr", "beat_value", "flag", "bass", "peak_index", "movement_per_sec"):
        assert (part["beat"][k] == full["beat"][k][:, 301:650]).all(), k
    _check(oracle, lib, cfg, pcm, 301, 650)


def test_interleaved_and_odd_frame_count(oracle, lib):
    pcm = oracle.synth_pcm(2, 2048 + 36 * 300)  # 37 frames: the last pair is half empty
    _check(oracle, lib, lib.reference_config(2048, 300, 100), pcm, interleaved=True)


def test_edge_signals(oracle, lib):
    """silence (0/0 threshold path), full scale, DC, impulse, on-bin sine."""
    n, hop, B = 2048, 512, 100
    S = n + 20 * hop
    t = np.arange(S)
    sigs = np.zeros((6, S), np.int16)
    sigs[1] = 32767
    sigs[2] = -32768
    sigs[3, ::n] = 20000
    sigs[4] = np.round(10000 * np.sin(2 * np.pi * 5 * t / n)).astype(np.int16)
    sigs[5] = np.where((t // 7) % 2 == 0, 32767, -32768)
    cfg = lib.reference_config(n, hop, B)
    got, ref = _check(oracle, lib, cfg, sigs)
    assert (got["bars_i32"][0] == 0).all()
    assert (got["beat"]["beat_value"][0] == np.iinfo(np.int32).min).all()  # 0/0 -> INT32_MIN like x86
    assert (got["beat"]["flag"][0] == 0).all()
    # on-bin sine: |X[5]| = A*N/2 -> 1024 at zoom 1e-4
    assert abs(got["bars"][4, 0, 5] - 1024.0) < 0.2 and got["bars_i32"][4, 0, 4] == 0


def test_circle_and_zoom(oracle, lib):
    pcm = oracle.synth_pcm(1, 2048 + 30 * 512)
    _check(oracle, lib, lib.reference_config(2048, 512, 500, zoom=1e-3, circle=1), pcm)
    _check(oracle, lib, lib.reference_config(2048, 512, 1000, zoom=1e-5), pcm)


def test_synth_pcm_device_bit_exact(oracle, lib):
    import torch
    out = torch.empty((3, 100000), dtype=torch.int16, device="cuda")
    lib.synth_pcm_device(out, first_sample=479000)
    torch.cuda.synchronize()
    ref = oracle.synth_pcm(3, 100000, first_sample=479000)
    assert (out.cpu().numpy() == ref).all()


def test_device_path_matches_host_path(oracle, lib):
    import torch
    cfg = lib.extended_config(4096, 512, 200)
    pcm = oracle.synth_pcm(2, 4096 + 999 * 512)
    with lib.BatchedSpectrum(cfg) as sp:
        host = sp.run(pcm)
        d = torch.from_numpy(pcm).cuda()
        F_ = sp.num_frames(pcm.shape[1])
        bars, i32, beat = sp.alloc_device_outputs(2, F_, "cuda", want_i32=True)
        sp.run_device(d, bars, i32, beat)
        torch.cuda.synchronize()
        calls, tot, bk = sp.timing()
        assert calls == 1 and tot >= bk > 0
        assert sp.launch_count >= 2
    assert (bars.cpu().numpy() == host["bars"]).all()
    assert (i32.cpu().numpy() == host["bars_i32"]).all()
    b = lib.BatchedSpectrum.beat_to_numpy(beat)
    for k in ("w", "thr", "beat_value", "flag", "bass", "peak_index", "movement_per_sec"):
        assert (b[k] == host["beat"][k]).all(), k


BEAT_FIELDS = ("w", "thr", "beat_value", "flag", "bass", "peak_index", "movement_per_sec")


def test_shards_around_beat_blocks_equal_full_run(oracle, lib):
    """The beat stage carries the tempo-ring transform forward inside absolute blocks of 96 frames
    (beatDetector.c:53-67): a shard starting anywhere relative to a block must reproduce the full run bit for bit,
    both through the fast path (4096/512) and the generic one (2048/300)."""
    for cfg, samples in ((lib.extended_config(4096, 512, 200), 4096 + 699 * 512), (lib.extended_config(2048, 300, 100), 2048 + 699 * 300)):
        pcm = oracle.synth_pcm(2, samples)
        with lib.BatchedSpectrum(cfg) as sp:
            full = sp.run(pcm)
            for fb, fe in ((1, 700), (95, 400), (96, 400), (97, 193), (191, 700), (192, 289), (353, 354), (640, 700)):
                part = sp.run(pcm, fb, fe)
                assert (part["bars"] == full["bars"][:, fb:fe]).all(), (cfg.n, fb, fe)
                assert (part["bars_i32"] == full["bars_i32"][:, fb:fe]).all(), (cfg.n, fb, fe)
                for k in BEAT_FIELDS:
                    assert (part["beat"][k] == full["beat"][k][:, fb:fe]).all(), (cfg.n, k, fb, fe)


def test_host_chunks_straddle_beat_blocks(oracle, lib):
    """fftl_stft_run moves a long job in frame-range chunks (about 21 000 frames each here) whose boundaries fall
    inside the beat stage's 96-frame blocks: the chunked host path must equal one device-resident call."""
    import torch
    cfg = lib.extended_config(4096, 512, 200)
    frames = 48000
    d = torch.empty((2, 4096 + (frames - 1) * 512), dtype=torch.int16, device="cuda")
    lib.synth_pcm_device(d)
    torch.cuda.synchronize()
    pcm = d.cpu().numpy()
    with lib.BatchedSpectrum(cfg) as sp:
        host = sp.run(pcm, 1000, frames, want_i32=False)
        bars, _, beat = sp.alloc_device_outputs(2, frames - 1000, "cuda", want_i32=False)
        sp.run_device(d, bars, None, beat, 1000, frames)
        torch.cuda.synchronize()
    assert (bars.cpu().numpy() == host["bars"]).all()
    b = lib.BatchedSpectrum.beat_to_numpy(beat)
    for k in BEAT_FIELDS:
        assert (b[k] == host["beat"][k]).all(), k
    # and the integer chain of a stretch that crosses a chunk boundary is what the oracle replays
    ocfg = oracle.Config()
    import ctypes as C
    C.memmove(C.byref(ocfg), C.byref(cfg), C.sizeof(ocfg))
    with lib.BatchedSpectrum(cfg) as sp:
        head = sp.run(pcm, 0, 1000, want_i32=False)
    stitched = np.concatenate([head["beat"][0], host["beat"][0, :23000]])
    nfr, ndiff = assert_peaks_explained(oracle, ocfg, stitched, "chunked host run")
    assert nfr == 24000


@pytest.mark.parametrize("hop", [441, 480, 768, 1000, 2048, 4096, 5000])
def test_fast_path_any_hop(oracle, lib, hop):
    """N = 4096 with a hop that is not 256/512/1024 runs the fast kernel's own-samples variant: against the
    oracle (parity bar), against the generic kernel (tolerance) and shard == full run (bit exact)."""
    cfg = lib.extended_config(4096, hop, 200)
    pcm = oracle.synth_pcm(2, 4096 + 130 * hop + 17)  # 131 frames, ragged tail
    got, _ = _check(oracle, lib, cfg, pcm)
    with lib.BatchedSpectrum(cfg) as sp:
        before = sp.launch_count
        part = sp.run(pcm, 37, 120)
        fast = sp.fast_launch_count
        assert fast > 0  # the runs above took the fast path
        sp.set_kernel(_capi.KERNEL_GENERIC)
        gen = sp.run(pcm)
        assert sp.fast_launch_count == fast
    assert (part["bars"] == got["bars"][:, 37:120]).all()
    for k in BEAT_FIELDS:
        assert (part["beat"][k] == got["beat"][k][:, 37:120]).all(), k
    fmax = np.abs(gen["bars"]).max(axis=-1, keepdims=True)
    assert (np.abs(gen["bars"] - got["bars"]) <= 2e-6 * fmax).all()
    # reference mode (no window, linear bins) through the same variant
    rcfg = lib.reference_config(4096, hop, 100)
    _check(oracle, lib, rcfg, pcm[:, :4096 + 40 * hop])


@pytest.mark.parametrize("hop", [480, 512])
def test_fast_path_interleaved_pcm(oracle, lib, hop):
    """Interleaved PCM at N = 4096 goes through the fast kernel's own-samples variant; the arithmetic per frame is
    that of the planar run, so the results are bit-identical (and the planar run is checked against the oracle)."""
    cfg = lib.extended_config(4096, hop, 200)
    pcm = oracle.synth_pcm(2, 4096 + 300 * hop + 5)
    planar, _ = _check(oracle, lib, cfg, pcm)
    with lib.BatchedSpectrum(cfg) as sp:
        inter = sp.run(pcm.T.copy(), interleaved=True)
        part = sp.run(pcm.T.copy(), 101, 250, interleaved=True)
    assert (inter["bars"] == planar["bars"]).all()
    assert (inter["bars_i32"] == planar["bars_i32"]).all()
    for k in BEAT_FIELDS:
        assert (inter["beat"][k] == planar["beat"][k]).all(), k
        assert (part["beat"][k] == planar["beat"][k][:, 101:250]).all(), k
    assert (part["bars"] == planar["bars"][:, 101:250]).all()


@pytest.mark.parametrize("n", [1024, 2048, 4096, 8192, 16384])
def test_fast_path_variants_both_sizes(oracle, lib, n):
    """The real-input fast path at the reference's own N = 2048 (global.h:3), at 4096 and at its other sizes: shared-sample variant
    (hop 512), own-sample variant (hop 300), windowed and not, one magnitude row and the full band including
    bin N/2, each against the oracle, against the generic kernel and shard == full."""
    pcm = oracle.synth_pcm(2, n + 150 * 512 + 3)
    cfgs = [lib.extended_config(n, 512, 100),                       # Hann, log bins to 20 kHz, SG
            lib.extended_config(n, 512, 64, fmax=2000.0),           # few magnitude rows
            lib.extended_config(n, 300, 100, sg_half=0),            # own-sample variant, no smoothing
            lib.reference_config(n, 512, min(n // 2 + 1, 4096)),    # every bin including N/2 (bars = N/2 + 1, FFTL_MAX_BARS at most)
            lib.reference_config(n, 512, 100, circle=1)]
    for cfg in cfgs:
        got, _ = _check(oracle, lib, cfg, pcm)
        with lib.BatchedSpectrum(cfg) as sp:
            part = sp.run(pcm, 33, 140)
            fast = sp.fast_launch_count
            # (a bar set whose tables do not fit beside the slots at this CTA count runs the generic kernel anyway)
            assert fast > 0 or cfg.bars > 256
            sp.set_kernel(_capi.KERNEL_GENERIC)
            gen = sp.run(pcm)
            assert sp.fast_launch_count == fast
        assert (part["bars"] == got["bars"][:, 33:140]).all()
        for k in BEAT_FIELDS:
            assert (part["beat"][k] == got["beat"][k][:, 33:140]).all(), k
        # two fp32 transforms with different operation orders: rounding relative to the largest BIN of the frame,
        # which a band-limited bar set need not contain
        fmax = np.abs(gen["bars"]).max(axis=-1, keepdims=True)
        assert (np.abs(gen["bars"] - got["bars"]) <= 1e-5 * fmax).all()


@pytest.mark.parametrize("n,hop", [(4096, 512), (2048, 512), (4096, 480), (1024, 480), (8192, 512), (16384, 512), (16384, 700)])
def test_fast_path_many_channels_few_frames(oracle, lib, n, hop):
    """More CTAs than a channel has tiles: the persistent kernel's tile cursor steps over several channels at once,
    and the last tile of every channel is ragged (9 frames = 2 tiles + 1 frame)."""
    channels = 37
    pcm = oracle.synth_pcm(channels, n + 8 * hop + 11)
    _check(oracle, lib, lib.extended_config(n, hop, 100), pcm)
    _check(oracle, lib, lib.reference_config(n, hop, 100), pcm, 3, 8)


@pytest.mark.parametrize("n,hop,bars", [(4096, 256, 200), (4096, 512, 200), (4096, 1024, 200), (2048, 512, 100), (4096, 512, 64), (4096, 480, 200),
                                        (4096, 441, 200), (4096, 5000, 200), (2048, 480, 100)])
def test_tensor_memory_variant_equals_register_variant(oracle, lib, n, hop, bars):
    """n = 4096 (hop 256 / 512 / 1024) and n = 2048 (hop 512) take the variant of the fast path that parks its per-thread
    constants in tensor memory and writes the heights one tile late (stft_rdif_tm.cuh): the same arithmetic in the same
    order as the register-resident kernel, so bars, (int) heights and beat records are equal bit for bit -- for a full
    run, for a shard that starts and ends inside tiles, for a job shorter than one tile per CTA and for many short
    channels."""
    _tm_equals_registers(oracle, lib, lib.extended_config(n, hop, bars), n, hop)
    # no room for the overlay when the whole band is read (n = 4096: bars up to the Nyquist frequency; n = 2048: one bar
    # per bin including bin n/2): the register variant of the fast path runs
    full = lib.extended_config(n, hop, bars, fmax=24000.0) if n == 4096 else lib.reference_config(n, hop, n // 2 + 1, zoom=1e-3)
    with lib.BatchedSpectrum(full) as sp:
        sp.run(oracle.synth_pcm(1, n + 9 * hop))
        assert sp.fast_launch_count > 0 and sp.tmem_launch_count == 0


@pytest.mark.parametrize("which", ["parity_rows1", "sg_off", "sg_runtime_width", "rows4", "circle"])
def test_tensor_memory_variant_other_instantiations(oracle, lib, which):
    """The other template instantiations of the tensor-memory kernel at n = 4096, hop 512: one magnitude row without a
    window (the reference's chain), Savitzky-Golay off / at a run-time width, four magnitude rows, circle mode."""
    n, hop = 4096, 512
    cfg = {"parity_rows1": lambda: lib.reference_config(n, hop, 100, zoom=1e-3),
           "sg_off": lambda: lib.extended_config(n, hop, 200, sg_half=0),
           "sg_runtime_width": lambda: lib.extended_config(n, hop, 200, sg_half=5, sg_degree=3),
           "rows4": lambda: lib.extended_config(n, hop, 150, fmax=11000.0),
           "circle": lambda: lib.extended_config(n, hop, 200, circle=1)}[which]()
    _tm_equals_registers(oracle, lib, cfg, n, hop, shapes=((2, 517), (3, 5)))


def _tm_equals_registers(oracle, lib, cfg, n, hop, shapes=((2, 1237), (1, 3), (37, 21))):
    for channels, frames in shapes:
        pcm = oracle.synth_pcm(channels, n + (frames - 1) * hop + 5)
        with lib.BatchedSpectrum(cfg) as sp:
            tm = sp.run(pcm)
            assert sp.tmem_launch_count > 0
            n_tm = sp.tmem_launch_count
            part = sp.run(pcm, 1, frames - 1) if frames > 2 else None
            sp.set_kernel(_capi.KERNEL_FAST_REGISTERS)
            reg = sp.run(pcm)
            assert sp.tmem_launch_count == (n_tm + (1 if part is not None else 0)) and sp.fast_launch_count > sp.tmem_launch_count
        assert (tm["bars"] == reg["bars"]).all() and (tm["bars_i32"] == reg["bars_i32"]).all()
        for k in BEAT_FIELDS:
            assert (tm["beat"][k] == reg["beat"][k]).all(), k
        if part is not None:
            assert (part["bars"] == reg["bars"][:, 1:frames - 1]).all()
            assert (part["bars_i32"] == reg["bars_i32"][:, 1:frames - 1]).all()
            for k in BEAT_FIELDS:
                assert (part["beat"][k] == reg["beat"][k][:, 1:frames - 1]).all(), k


def test_tensor_memory_variant_is_race_free_under_repetition(oracle, lib):
    """The tensor-memory kernel hands buffers between phases with one split barrier and four CTA barriers per tile
    (stft_rdif_tm.cuh): a missing one shows up as run-to-run differences (that is how the barrier in front of the next
    tile's pass A was found missing during development).  Forty device-resident runs of a job with ~25 tiles per CTA,
    with a second stream keeping the memory system busy, must all equal the register variant bit for bit."""
    import torch
    cfg = lib.extended_config(4096, 512, 200)
    frames = 22000
    d = torch.empty((2, 4096 + (frames - 1) * 512), dtype=torch.int16, device="cuda")
    lib.synth_pcm_device(d)
    noise_a = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    noise_b = torch.empty_like(noise_a)
    side = torch.cuda.Stream()
    with lib.BatchedSpectrum(cfg) as sp:
        bars, _, beat = sp.alloc_device_outputs(2, frames, "cuda", want_i32=False)
        sp.set_kernel(_capi.KERNEL_FAST_REGISTERS)
        sp.run_device(d, bars, None, beat)
        torch.cuda.synchronize()
        ref_bars, ref_beat = bars.clone(), beat.clone()
        sp.set_kernel(_capi.KERNEL_AUTO)
        for it in range(40):
            bars.zero_()
            beat.zero_()
            if it % 2:
                with torch.cuda.stream(side):
                    noise_b.copy_(noise_a, non_blocking=True)
            sp.run_device(d, bars, None, beat)
            torch.cuda.synchronize()
            assert torch.equal(bars, ref_bars), "run %d" % it
            assert torch.equal(beat, ref_beat), "run %d" % it
        assert sp.tmem_launch_count == 40


def test_tensor_memory_variant_without_beat_outputs(oracle, lib):
    """want_beat=False (no band-energy / bass side outputs, aux pointers null): the tensor-memory kernel skips its band
    energy and bass steps; the bars are the bits of the run with beat outputs."""
    cfg = lib.extended_config(4096, 512, 200)
    pcm = oracle.synth_pcm(3, 4096 + 300 * 512 + 77)
    with lib.BatchedSpectrum(cfg) as sp:
        a = sp.run(pcm)
        n0 = sp.tmem_launch_count
        b = sp.run(pcm, want_beat=False)
        assert sp.tmem_launch_count > n0
        c = sp.run(pcm, 5, 290, want_beat=False, want_i32=False)
    assert (a["bars"] == b["bars"]).all() and (a["bars_i32"] == b["bars_i32"]).all()
    assert (c["bars"] == a["bars"][:, 5:290]).all()
