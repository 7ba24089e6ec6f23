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
            "ms": ms, "bars_kernel_ms": bms, "frames_per_s": frames / ms * 1e3, "kernel": "stft_bars_rdif_kernel<2048> (fast path, hop 512)"})

# ---- c1 scaled up: the reference's configuration on 1 h of stereo audio (batch throughput of the N = 2048 fast path)
pcm = torch.empty((2, 48000 * 3600), dtype=torch.int16, device=dev)
F.synth_pcm_device(pcm)
for label, cfg in (("reference mode (no window, bar i = bin i)", F.reference_config(2048, 512, 100)),
                   ("extended mode (Hann, 100 log bars, SG)", F.extended_config(2048, 512, 100))):
    with F.BatchedSpectrum(cfg) as sp:
        frames, ms, bms = timed_device(sp, pcm)
    out.append({"config": "c1 x 360: 1 h stereo 48 kHz, 2048-pt, hop 512, 100 bars + beat, " + label, "frames": frames, "ms": ms,
                "bars_kernel_ms": bms, "frames_per_s": frames / ms * 1e3, "bars_kernel_frames_per_s": frames / bms * 1e3,
                "roofline_frac_nominal_fp32": frames / bms * 1e3 * 5 * 2048 * 11 / 74.45e12,
                "kernel": "stft_bars_rdif_kernel<2048> (fast path, hop 512)"})
del pcm

# ---- c3: 1024 concurrent mono streams, 1024-pt window, 64 bars: one tick (512 new samples per stream) per batch
# through the streaming front end (device-resident sliding windows + beat state)
STREAMS, COUNT = 1024, 512
src = torch.empty((STREAMS, COUNT * 64), dtype=torch.int16, device=dev)
F.synth_pcm_device(src)
cfg3 = F.extended_config(1024, COUNT, 64, sg_half=3, sg_degree=2)
with F.StreamingSpectrum(cfg3, STREAMS) as ss:
    fresh = src[:, :COUNT].contiguous()
    bars = torch.empty((STREAMS, 64), dtype=torch.float32, device=dev)
    beat = torch.empty((STREAMS, 8), dtype=torch.int32, device=dev)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(5):
            ss.push_device(fresh, bars, None, beat, stream=st)
        st.synchronize()
        lat = []
        for _ in range(2000):
            t0 = time.perf_counter()
            ss.push_device(fresh, bars, None, beat, stream=st)
            st.synchronize()
            lat.append((time.perf_counter() - t0) * 1e6)
        lat_plain = np.array(lat[200:])
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            ss.push_device(fresh, bars, None, beat, stream=st)
        lat = []
        for _ in range(2000):
            t0 = time.perf_counter()
            g.replay()
            st.synchronize()
            lat.append((time.perf_counter() - t0) * 1e6)
        lat_graph = np.array(lat[200:])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(200):
            g.replay()
        e1.record(st)
        st.synchronize()
        dev_us = e0.elapsed_time(e1) * 1e3 / 200
    h_fresh = F.PinnedArray((STREAMS, COUNT), np.int16)
    h_fresh.array[:] = fresh.cpu().numpy()
    o = {"bars": F.PinnedArray((STREAMS, 64), np.float32).array, "beat": F.PinnedArray((STREAMS,), F.BEAT_DTYPE).array}
    for _ in range(20):
        ss.push(h_fresh.array, want_i32=False, out=o)
    lat = []
    for _ in range(1000):
        t0 = time.perf_counter()
        ss.push(h_fresh.array, want_i32=False, out=o)
        lat.append((time.perf_counter() - t0) * 1e6)
    lat_host = np.array(lat[100:])
pct = lambda a: {"p50_us": float(np.percentile(a, 50)), "p99_us": float(np.percentile(a, 99))}
out.append({"config": "c3: 1024 concurrent mono streams, 1024-pt Hann, 64 log bars + SG + beat, one tick (512 new samples) per stream per batch",
            "api": "fftl_stream_push[_device] (sliding windows and beat state resident on the GPU)",
            "frames_per_batch": STREAMS, "device_us_per_batch_graph_back_to_back": dev_us,
            "launch_sync_python": pct(lat_plain), "cuda_graph_replay_sync_python": pct(lat_graph),
            "c_abi_host_buffers_pinned": pct(lat_host), "frames_per_s_device": STREAMS / dev_us * 1e6,
            "h2d_bytes_per_batch": STREAMS * COUNT * 2, "d2h_bytes_per_batch": STREAMS * 64 * 4 + STREAMS * 32,
            "note": "latency bound: 1024 frames are ~1 us of work at the roofline; one kernel per tick (stream_tick_kernel<1024>: window move, 4-point-per-thread transform, bars and beat stage of two streams per CTA); tools/tick_bench.py times both forms"})

# ---- c4: 16384-pt window, 512 bars, stereo (1 min), hop 512 and 4096
pcm = torch.empty((2, 48000 * 60), dtype=torch.int16, device=dev)
F.synth_pcm_device(pcm)
for hop in (512, 4096):
    with F.BatchedSpectrum(F.extended_config(16384, hop, 512)) as sp:
        frames, ms, bms = timed_device(sp, pcm, reps=5)
    out.append({"config": "c4: 16384-pt Hann, hop %d, 512 log bars + SG + beat, 60 s stereo" % hop, "frames": frames, "ms": ms,
                "bars_kernel_ms": bms, "frames_per_s": frames / ms * 1e3, "kernel": "stft_bars_rdif_kernel<16384> (fast path, own-samples variant)",
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
