"""Frame-range shards on the GPU (BASELINE configs[4]): every rank uploads ONLY its own sample span
(sharding.sample_span) and runs fftl_stft_run_device_span with the job's absolute frame numbers; the shards put
together must equal the single-GPU run bit for bit, beat records included."""
import numpy as np
import pytest

from fft_lines_b200 import _capi

pytestmark = pytest.mark.gpu

BEAT_FIELDS = ("w", "thr", "beat_value", "flag", "bass", "peak_index", "movement_per_sec")


def _full_device_run(lib, sp, pcm):
    import torch
    d = torch.from_numpy(pcm).cuda()
    nf = sp.num_frames(pcm.shape[1])
    bars, i32, beat = sp.alloc_device_outputs(pcm.shape[0], nf, "cuda", want_i32=True)
    sp.run_device(d, bars, i32, beat)
    torch.cuda.synchronize()
    return bars.cpu().numpy(), i32.cpu().numpy(), lib.BatchedSpectrum.beat_to_numpy(beat)


@pytest.mark.parametrize("n,hop,bars,mode", [(4096, 512, 200, "x"), (2048, 512, 100, "p"), (2048, 300, 100, "x"),
                                             (16384, 512, 512, "x"), (8192, 512, 256, "x"), (512, 128, 64, "x")])
@pytest.mark.parametrize("tail", [0, 77])
def test_span_only_shards_equal_single_run(oracle, lib, n, hop, bars, mode, tail):
    """world = 2, 3 and 8 ranks, each holding only sample_span(...): the gathered result is the full run.
    tail != 0: (S - n) % hop != 0, so the phantom partner of an odd job's last frame is partly real audio."""
    import torch
    sh = lib.sharding
    frames = 1203 if n <= 4096 else 403  # odd: the last frame has no partner
    S = n + (frames - 1) * hop + tail
    pcm = oracle.synth_pcm(2, S)
    cfg = lib.extended_config(n, hop, bars) if mode == "x" else lib.reference_config(n, hop, bars)
    with lib.BatchedSpectrum(cfg) as sp:
        assert sp.num_frames(S) == frames
        fbars, fi32, fbeat = _full_device_run(lib, sp, pcm)
        for world in (2, 3, 8):
            for rank in range(world):
                fb, fe = sh.shard_frames(frames, world, rank)
                s0, s1 = sh.sample_span(fb, fe, n, hop, True, S)
                assert (s0, s1) == sp.sample_span(S, fb, fe, True)
                span = torch.from_numpy(np.ascontiguousarray(pcm[:, s0:s1])).cuda()  # nothing else is on this "rank"
                b2, e2, b, i, bt = sh.run_shard(sp, span, s0, S, world, rank, want_i32=True)
                torch.cuda.synchronize()
                assert (b2, e2) == (fb, fe)
                assert (b.cpu().numpy() == fbars[:, fb:fe]).all(), (world, rank)
                assert (i.cpu().numpy() == fi32[:, fb:fe]).all(), (world, rank)
                got = lib.BatchedSpectrum.beat_to_numpy(bt)
                for k in BEAT_FIELDS:
                    assert (got[k] == fbeat[k][:, fb:fe]).all(), (k, world, rank)


def test_span_must_cover_what_the_frames_read(oracle, lib):
    import torch
    cfg = lib.extended_config(4096, 512, 200)
    S = 4096 + 999 * 512
    pcm = oracle.synth_pcm(1, S)
    with lib.BatchedSpectrum(cfg) as sp:
        s0, s1 = sp.sample_span(S, 500, 700)
        assert s0 == (480 - 256) * 512 and s1 == 699 * 512 + 4096
        assert sp.sample_span(S, 500, 700, with_beat=False) == (500 * 512, s1)
        assert sp.sample_span(S, 500, 701)[1] == 701 * 512 + 4096       # even last frame: the partner too
        assert sp.sample_span(S, 500, 1000)[1] == S                       # clipped to the job
        bars, _, beat = sp.alloc_device_outputs(1, 200, "cuda")
        short = torch.from_numpy(np.ascontiguousarray(pcm[:, s0 + 1:s1])).cuda()
        with pytest.raises(lib.FftLinesError) as ei:
            sp.run_device_span(short, s0 + 1, S, bars, None, beat, 500, 700)
        assert ei.value.code == _capi.FFTL_ERR_ARG
        short = torch.from_numpy(np.ascontiguousarray(pcm[:, s0:s1 - 1])).cuda()
        with pytest.raises(lib.FftLinesError):
            sp.run_device_span(short, s0, S, bars, None, beat, 500, 700)
        # without beat output the halo is not needed
        nb = torch.from_numpy(np.ascontiguousarray(pcm[:, 500 * 512:s1])).cuda()
        sp.run_device_span(nb, 500 * 512, S, bars, None, None, 500, 700)
        torch.cuda.synchronize()


@pytest.mark.parametrize("n,hop", [(4096, 512), (4096, 480), (2048, 300), (512, 200)])
def test_host_path_equals_device_path_on_ragged_odd_jobs(oracle, lib, n, hop):
    """(S - n) % hop != 0 and an odd frame count: the chunked host call uploads the phantom partner of the last
    frame exactly as far as the job has samples, so fftl_stft_run == fftl_stft_run_device bit for bit."""
    frames = 1001
    S = n + (frames - 1) * hop + hop - 1
    pcm = oracle.synth_pcm(2, S)
    cfg = lib.extended_config(n, hop, 100)
    with lib.BatchedSpectrum(cfg) as sp:
        assert sp.num_frames(S) == frames
        host = sp.run(pcm)
        dbars, di32, dbeat = _full_device_run(lib, sp, pcm)
    assert (host["bars"] == dbars).all()
    assert (host["bars_i32"] == di32).all()
    for k in BEAT_FIELDS:
        assert (host["beat"][k] == dbeat[k]).all(), k
