"""Frame-range sharding: host-side split/halo logic and the
world_size-2 gather over gloo.  The compute inside the workers is the CPU
oracle (test stand-in only): what is under test is that independently
computed shards, gathered, equal the single-run result."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_frames_partition(lib):
    sh = lib.sharding
    for total, world in ((337493, 8), (10, 3), (7, 8), (0, 2), (1000, 1)):
        spans = [sh.shard_frames(total, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c and b >= a
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.shard_frames(10, 2, 2)


def test_sample_span_includes_beat_halo(lib):
    sh = lib.sharding
    assert sh.sample_span(0, 10, 4096, 512) == (0, 9 * 512 + 4096)
    assert sh.sample_span(1000, 2000, 4096, 512) == ((960 - 256) * 512, 1999 * 512 + 4096)
    assert sh.sample_span(1001, 2000, 4096, 512, with_beat=False) == (1000 * 512, 1999 * 512 + 4096)
    assert sh.sample_span(100, 200, 4096, 512) == (0, 199 * 512 + 4096)
    assert sh.sample_span(5, 5, 4096, 512) == (0, 0)
    # an even last frame shares its packed sub-transforms with its partner: the partner's samples belong to the span,
    # as far as the job has them
    assert sh.sample_span(1000, 2001, 4096, 512) == ((960 - 256) * 512, 2001 * 512 + 4096)
    assert sh.sample_span(1000, 2001, 4096, 512, total_samples=2000 * 512 + 4096 + 100) == ((960 - 256) * 512, 2000 * 512 + 4096 + 100)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from fft_lines_b200 import sharding as sh
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = O.extended_config(1024, 256, 64)
    S = 1024 + 256 * 600
    total = O.num_frames(cfg, S)
    b, e = sh.shard_frames(total, world, rank)
    s0, s1 = sh.sample_span(b, e, cfg.n, cfg.hop, True, S)
    # each rank only materialises its own sample span (plus halo), like a GPU shard would
    local_pcm = O.synth_pcm(2, s1 - s0, first_sample=s0)
    f_off = s0 // cfg.hop
    r = O.stft(cfg, local_pcm, b - f_off, e - f_off)
    bars = sh.gather_frames(torch.from_numpy(r["bars"]), total)
    beat = sh.gather_frames(torch.from_numpy(r["beat"].view(np.int32).reshape(2, e - b, 8)), total, dst=0)
    if rank == 0:
        np.save(os.path.join(tmp, "bars.npy"), bars.numpy())
        np.save(os.path.join(tmp, "beat.npy"), beat.numpy())
    dist.destroy_process_group()


def test_two_rank_gloo_shards_equal_single_run(oracle, tmp_path):
    import torch.multiprocessing as mp
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    cfg = oracle.extended_config(1024, 256, 64)
    pcm = oracle.synth_pcm(2, 1024 + 256 * 600)
    full = oracle.stft(cfg, pcm)
    bars = np.load(tmp_path / "bars.npy")
    beat = np.load(tmp_path / "beat.npy").view(oracle.BEAT_DTYPE).reshape(2, -1)
    assert (bars == full["bars"]).all()
    for k in ("w", "thr", "beat_value", "flag", "bass", "peak_index", "movement_per_sec"):
        assert (beat[k] == full["beat"][k]).all(), k


def test_first_frame_needed_matches_library(lib):
    """The Python mirror agrees with fftl_stft_first_frame_needed (host arithmetic only: runs without a GPU)."""
    from fft_lines_b200 import _capi, sharding as sh
    L = _capi.lib()
    for fb in (0, 1, 95, 96, 97, 255, 256, 257, 351, 352, 353, 1000, 1001, 10 ** 9 + 7):
        for wb in (0, 1):
            assert L.fftl_stft_first_frame_needed(fb, wb) == sh.first_frame_needed(fb, bool(wb)), (fb, wb)
    assert sh.first_frame_needed(400) == 384 - 256 and sh.first_frame_needed(100) == 0
