"""Frame-range sharding of one STFT job across GPUs.

Frames are independent up to the bar heights, and the beat stages need only two
integers per frame of a bounded history (at most 96 + 256 frames back), which
``fftl_stft_run*`` recomputes itself for any ``frame_begin``.  So a shard is just a ``[begin, end)`` frame range: no
data-path collective, only an optional final gather of the outputs.
"""


def shard_frames(total_frames, world_size, rank):
    """Contiguous, balanced split: the first ``total % world`` ranks get one more frame."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size/rank")
    base, extra = divmod(int(total_frames), world_size)
    begin = rank * base + min(rank, extra)
    end = begin + base + (1 if rank < extra else 0)
    return begin, end


BEAT_BLOCK = 96   # frames per block of the beat stage (csrc/kernels.cuh BEAT_FB)
TEMPO_RING = 256  # beatDetector.c:53-59


def first_frame_needed(frame_begin, with_beat=True):
    """Mirror of ``fftl_stft_first_frame_needed``: the first frame whose samples a run from
    ``frame_begin`` reads (the tempo-ring state is carried forward from the start of the
    frame's 96-frame block, which itself needs 256 frames of history)."""
    fb = max(0, int(frame_begin))
    if with_beat:
        fb = max(0, fb // BEAT_BLOCK * BEAT_BLOCK - TEMPO_RING)
    return fb & ~1


def sample_span(frame_begin, frame_end, n, hop, with_beat=True):
    """Samples a shard must hold: its frames plus the beat warm-up halo before it."""
    if frame_end <= frame_begin:
        return 0, 0
    return first_frame_needed(frame_begin, with_beat) * hop, (frame_end - 1) * hop + n


def gather_frames(local, total_frames, group=None, dst=None):
    """All-gather (or gather to ``dst``) per-rank ``[channels, frames_r, ...]`` tensors along
    the frame axis with torch.distributed.  Works on CPU (gloo) and GPU (nccl) tensors.
    Returns the full tensor (on ``dst`` only when given, else on every rank)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_frames(total_frames, world, r) for r in range(world)]
    # collectives need equal shapes: pad every shard to the largest (they differ by at most one frame)
    fmax = max(e - b for b, e in sizes)
    shape = list(local.shape)
    shape[1] = fmax
    padded = torch.zeros(shape, dtype=local.dtype, device=local.device)
    padded[:, :local.shape[1]] = local
    parts = [torch.empty(shape, dtype=local.dtype, device=local.device) for _ in range(world)]
    if dst is None:
        dist.all_gather(parts, padded, group=group)
    else:
        dist.gather(padded, parts if rank == dst else None, dst=dst, group=group)
        if rank != dst:
            return None
    return torch.cat([p[:, :e - b] for p, (b, e) in zip(parts, sizes)], dim=1)
