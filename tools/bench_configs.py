"""Secondary measurements for the BASELINE.json configs that are not the bench line
(c1, c3, c4, c5-style chunked run).  Writes one JSON object per config.

    python tools/bench_configs.py > gpurun_out/configs.json
"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import fft_lines_b200 as F

dev = torch.device("cuda", 0)
out = []


def timed_device(sp, pcm, want_beat=True, reps=10):
    ch, S = pcm.shape
    nf = sp.num_frames(S)
    bars, _, beat = sp.alloc_device_outputs(ch, nf, dev, want_beat=want_beat)
    for _ in range(3):
        sp.run_device(pcm, bars, None, beat)
    torch.cuda.synchronize()
    sp.timing()
    for _ in range(reps):
        sp.run_device(pcm, bars, None, beat)
    torch.cuda.synchronize()
    calls, tot, bk = sp.timing()
    return ch * nf, tot / calls, bk / calls


# ---- c1: the reference's own configuration, 10 s stereo, N=2048, 100 bars, parity mode
pcm = torch.empty((2, 480000), dtype=torch.int16, device=dev)
F.synth_pcm_device(pcm)
with F.BatchedSpectrum(F.reference_config(2048, 512, 100)) as sp:
    frames, ms, bms = timed_device(sp, pcm)
out.append({"config": "c1: 10 s stereo 48 kHz, 2048-pt, hop 512 (assumed), 100 bars, reference mode", "frames": frames,
            "ms": ms, "bars_kernel_ms": bms, "frames_per_s": frames / ms * 1e3, "kernel": "stft_bars_kernel<2048> (generic)"})

# ---- c3: 1024 concurrent mono streams, one 1024-sample frame each per batch, 64 bars -> latency per batch
pcm = torch.empty((1024, 1024), dtype=torch.int16, device=dev)
F.synth_pcm_device(pcm)
cfg3 = F.extended_config(1024, 1024, 64, sg_half=3, sg_degree=2)
with F.BatchedSpectrum(cfg3) as sp:
    bars, _, beat = sp.alloc_device_outputs(1024, 1, dev)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(5):
            sp.run_device(pcm, bars, None, beat, stream=st)
        st.synchronize()
        # (a) plain launches, synchronised per batch (what a real-time caller sees from Python)
        lat = []
        for _ in range(2000):
            t0 = time.perf_counter()
            sp.run_device(pcm, bars, None, beat, stream=st)
            st.synchronize()
            lat.append((time.perf_counter() - t0) * 1e6)
        lat_plain = np.array(lat[200:])
        # (b) CUDA graph replay
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            sp.run_device(pcm, bars, None, beat, stream=st)
        lat = []
        for _ in range(2000):
            t0 = time.perf_counter()
            g.replay()
            st.synchronize()
            lat.append((time.perf_counter() - t0) * 1e6)
        lat_graph = np.array(lat[200:])
        # (c) device time per batch: 200 back-to-back replays between two events
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(200):
            g.replay()
        e1.record(st)
        st.synchronize()
        dev_us = e0.elapsed_time(e1) * 1e3 / 200
    # (d) through the C-ABI host call with pinned buffers (H2D 2 MB, D2H 0.3 MB per batch)
    h_pcm = F.PinnedArray((1024, 1024), np.int16)
    h_pcm.array[:] = pcm.cpu().numpy()
    h_bars = F.PinnedArray((1024, 1, 64), np.float32)
    h_beat = F.PinnedArray((1024, 1), F.BEAT_DTYPE)
    o = {"bars": h_bars.array, "beat": h_beat.array}
    for _ in range(20):
        sp.run(h_pcm.array, want_i32=False, out=o)
    lat = []
    for _ in range(1000):
        t0 = time.perf_counter()
        sp.run(h_pcm.array, want_i32=False, out=o)
        lat.append((time.perf_counter() - t0) * 1e6)
    lat_host = np.array(lat[100:])
pct = lambda a: {"p50_us": float(np.percentile(a, 50)), "p99_us": float(np.percentile(a, 99))}
out.append({"config": "c3: 1024 concurrent mono streams, 1024-pt Hann, 64 log bars + SG + beat, one frame per stream per batch",
            "frames_per_batch": 1024, "device_us_per_batch_graph_back_to_back": dev_us,
            "launch_sync_python": pct(lat_plain), "cuda_graph_replay_sync_python": pct(lat_graph),
            "c_abi_host_buffers_pinned": pct(lat_host), "frames_per_s_device": 1024 / dev_us * 1e6,
            "note": "latency bound: 1024 frames are ~1 us of work at the roofline; kernels: stft_bars_kernel<1024> + beat_post_kernel"})

# ---- c4: 16384-pt window, 512 bars, stereo (1 min), hop 512 and 4096
pcm = torch.empty((2, 48000 * 60), dtype=torch.int16, device=dev)
F.synth_pcm_device(pcm)
for hop in (512, 4096):
    with F.BatchedSpectrum(F.extended_config(16384, hop, 512)) as sp:
        frames, ms, bms = timed_device(sp, pcm, reps=5)
    out.append({"config": "c4: 16384-pt Hann, hop %d, 512 log bars + SG + beat, 60 s stereo" % hop, "frames": frames, "ms": ms,
                "bars_kernel_ms": bms, "frames_per_s": frames / ms * 1e3, "kernel": "stft_bars_kernel<16384> (generic)",
                "roofline_frac_nominal_fp32": (frames / (bms * 1e-3)) * 5 * 16384 * 14 / 74.45e12})

# ---- c5-style: 8 channels, chunked hours generated on the device and processed chunk by chunk (nothing is resident
# beyond one chunk): 8 x 30 min here; the 1000 h job is the same loop run longer / sharded by frame range over GPUs
CH, CHUNK_S, CHUNKS = 8, 48000 * 300, 6
cfg5 = F.extended_config(4096, 512, 200)
with F.BatchedSpectrum(cfg5) as sp:
    halo = 4096 - 512
    buf = torch.empty((CH, CHUNK_S + halo), dtype=torch.int16, device=dev)
    nfc = sp.num_frames(CHUNK_S + halo)
    bars, _, _ = sp.alloc_device_outputs(CH, nfc, dev, want_beat=False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    total = 0
    for c in range(CHUNKS):
        F.synth_pcm_device(buf, first_sample=c * CHUNK_S)      # stands in for the decoder / file reader
        sp.run_device(buf, bars, None, None)
        total += CH * nfc
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    sp.timing()
    sp.run_device(buf, bars, None, None)
    torch.cuda.synchronize()
    _, tot, bk = sp.timing()
out.append({"config": "c5-style: 8 channels x 30 min, 4096-pt Hann, hop 512, 200 bars, generated and processed in 5-min chunks on one GPU",
            "frames": total, "wall_s_including_on_device_generation": dt, "frames_per_s_including_generation": total / dt,
            "frames_per_s_stft_only": CH * nfc / (tot * 1e-3), "note": "frame-range sharding across GPUs: bench.py --gpus N"})
print(json.dumps(out, indent=1))
