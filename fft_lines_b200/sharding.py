"""Frame-range sharding of one STFT job across GPUs.

Frames are independent up to the bar heights, and the beat stages need only two
integers per frame of a bounded history (at most 96 + 256 frames back), which
``fftl_stft_run*`` recomputes itself for any ``frame_begin``.  So a shard is just a ``[begin, end)`` frame range: no
data-path collective, only an optional final gather of the outputs.  A rank that holds only its own samples
(``sample_span``) runs ``fftl_stft_run_device_span`` with the job's absolute frame numbers (``run_shard`` below).
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


def sample_span(frame_begin, frame_end, n, hop, with_beat=True, total_samples=None):
    """Samples ``[first, end)`` a shard must hold (mirror of ``fftl_stft_sample_span``): its frames, the beat warm-up
    halo before them and, when the last frame is even, its pair partner (frames 2k and 2k+1 share packed
    sub-transforms, so the partner's samples decide the low bits), clipped to the job's ``total_samples``."""
    if frame_end <= frame_begin:
        return 0, 0
    end = ((frame_end - 1) | 1) * hop + n
    if total_samples is not None:
        end = min(end, int(total_samples))
    return first_frame_needed(frame_begin, with_beat) * hop, end


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


def run_shard(sp, pcm_span, span_first_sample, total_samples, world_size, rank, want_i32=False, want_beat=True, stream=None):
    """This rank's frame range of one job through ``fftl_stft_run_device_span``.  ``pcm_span``: torch int16 CUDA
    [channels, span] holding samples ``sample_span(...)`` of the job.  Returns (frame_begin, frame_end, bars, i32, beat)."""
    total_frames = sp.num_frames(total_samples)
    fb, fe = shard_frames(total_frames, world_size, rank)
    bars, i32, beat = sp.alloc_device_outputs(pcm_span.shape[0], fe - fb, pcm_span.device, want_i32=want_i32, want_beat=want_beat)
    sp.run_device_span(pcm_span, span_first_sample, total_samples, bars, i32, beat, fb, fe, stream=stream)
    return fb, fe, bars, i32, beat
