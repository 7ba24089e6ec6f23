#!/usr/bin/env python
"""bench.py -- STFT frames/s -> bars on N B200s (BASELINE.json's metric).

A *step* is one pass of the hot path over one batch: 1 h of synthetic stereo
48 kHz int16 PCM per GPU -> 4096-pt Hann STFT, hop 512 -> 200 log-frequency
bars with Savitzky-Golay smoothing + band-energy beat test + tempo estimate
(BASELINE.json configs[1]).  Frames shard by range/stream with no data-path
collective, so N GPUs run N such batches ("weak" scaling).

  value     frames/s, inputs and outputs resident in HBM (CUDA events, max over ranks)
  e2e       same metric through the C-ABI host call fftl_stft_run: pinned host
            PCM in, host bars + beat records out, copies inside the timed region
  roofline  the dominant kernel (stft_bars) against its FP32 compute roofline
            (5*N*log2 N flop per frame) and its HBM byte roofline
  sharded   BASELINE configs[4] in small: ONE 8-channel job cut into frame ranges (sharding.shard_frames), every
            rank generating and holding only its own sample span and running fftl_stft_run_device_span; strong-
            scaling efficiency against the same job on one GPU, every shard compared bit for bit with the
            single-GPU result, and the NCCL gather of the outputs timed separately
  sustained the same step back to back for >= 3 s with the clock sampler running
  latency   BASELINE configs[2]: one tick of 1024 streams x 1024 points through the streaming front end
  cpu_baseline  the CPU port of the reference chain on this box's host cores (all cores and one thread) and the
            same chain on library FFTs (pocketfft, torch.fft)

`--impl reference` times the reference's CPU implementation of the same
configuration instead (oracle/ port: the reference has no Hann/log/SG stages
and is compiled for N=2048, so the unmodified files cannot run this config).
"""
import argparse
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FFT, HOP, BARS, CHANNELS, FS = 4096, 512, 200, 2, 48000
SECONDS = 3600
WORKLOAD = "c2: 1 h synthetic stereo 48 kHz, 4096-pt Hann, hop 512, 200 log bars + SG + beat detect"
FLOPS_PER_FRAME = 5.0 * N_FFT * math.log2(N_FFT)          # 245 760
BYTES_PER_FRAME = HOP * 2 + BARS * 4 + 8                   # 1832
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12       # 74.45


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index, power=False):
        self.index = index
        self.power = [] if power else None
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nvml = None

    def _loop(self):
        nv = self._nvml
        names = {}
        for k in dir(nv):
            if k.startswith("nvmlClocksEventReason") or k.startswith("nvmlClocksThrottleReason"):
                v = getattr(nv, k)
                if isinstance(v, int) and v:
                    names.setdefault(v, k.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", ""))
        getr = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
            nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                if self.power is not None:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
                if getr:
                    bits = getr(self._h)
                    for v, k in names.items():
                        if bits & v and k not in ("None", "All", "GpuIdle"):
                            self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.01)

    def start(self):
        if self._nvml:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        if self.samples:
            s = sorted(self.samples)
            out["sm_mhz"] = s[len(s) // 2]
            out["sm_min_mhz"] = s[0]
            out["samples"] = len(s)
        if self.power:
            pw = sorted(self.power)
            out["power_w_median"], out["power_w_max"] = pw[len(pw) // 2], pw[-1]
        return out


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def cpu_leg(threads, target_seconds=12.0):
    """CPU port of the chain (oracle/cpu_port.c) on a bounded sample of the same workload."""
    from oracle import oracle as O
    O.build()
    cfg = O.extended_config(N_FFT, HOP, BARS, sample_rate=float(FS))
    if threads is None:
        threads = O.lib().port_max_threads()
    # calibrate on 20 s of audio, then size the sample for ~target_seconds
    cal = O.synth_pcm(CHANNELS, FS * 20)
    sec, _ = O.port_run(cfg, cal, threads=threads, want_i32=False)
    frames = CHANNELS * O.num_frames(cfg, cal.shape[1])
    rate = frames / sec
    want = min(SECONDS, max(30, int(target_seconds * rate / CHANNELS * HOP / FS)))
    pcm = O.synth_pcm(CHANNELS, FS * want)
    return O, cfg, pcm, threads, want


def cpu_baselines():
    """Three CPU figures for the same chain on this box (bounded samples, ~25 s in all): the fp32 port of the
    reference chain on all cores (the reference-shaped arm: scalar code, this repo's FFT shim), the same port on ONE
    thread (how the reference actually runs, fft_lines.c:161-288), and the chain on library FFTs on all cores."""
    O, cfg, pcm, threads, want = cpu_leg(None, target_seconds=8.0)
    frames = CHANNELS * O.num_frames(cfg, pcm.shape[1])
    sec, _ = O.port_run(cfg, pcm, threads=threads, want_i32=False)
    out = {"value": frames / sec, "unit": "frames/s", "cores": threads, "kind": "port",
           "fft_provider": "fftwf-shim (this repo's scalar radix-4 FFT: no host FFTW in this image)",
           "sample": "first %d s of the same workload (%d frames), fp32 port of the reference chain" % (want, frames),
           "seconds": sec, "others": []}
    n1 = max(10, want // threads)
    p1 = pcm[:, :FS * n1]
    f1 = CHANNELS * O.num_frames(cfg, p1.shape[1])
    sec1, _ = O.port_run(cfg, p1, threads=1, want_i32=False)
    out["others"].append({"value": f1 / sec1, "unit": "frames/s", "cores": 1, "kind": "port", "fft_provider": "fftwf-shim",
                          "sample": "first %d s (%d frames), one thread as the reference runs" % (n1, f1), "seconds": sec1})
    try:
        from oracle import numpy_chain as NC
        for prov, label in (("scipy", "pocketfft (scipy.fft.rfft)"), ("torch", "torch.fft.rfft (CPU)")):
            NC.run(O, cfg, pcm[:, :FS * 5], provider=prov)  # warm-up: imports, plans, thread pool
            t0 = time.perf_counter()
            r = NC.run(O, cfg, pcm, provider=prov)
            dt = time.perf_counter() - t0
            out["others"].append({"value": frames / dt, "unit": "frames/s", "cores": r["workers"], "kind": "port",
                                  "fft_provider": label, "seconds": dt,
                                  "sample": "first %d s (%d frames): batched real FFT + the same Hann / log-bin / SG / beat "
                                            "tail in numpy, blocks of frames over all cores (oracle/numpy_chain.py)" % (want, frames)})
        for prov, label in (("scipy", "pocketfft (scipy.fft.rfft)"), ("torch", "torch.fft.rfft (CPU)")):
            NC.fft_only(cfg, pcm[:, :FS * 20], provider=prov)  # warm-up
            dt, nfr, wk = NC.fft_only(cfg, pcm, provider=prov)
            out["others"].append({"value": nfr / dt, "unit": "frames/s", "cores": wk, "kind": "fft-only upper bound",
                                  "fft_provider": label, "seconds": dt,
                                  "sample": "windowed real FFTs of the same %d frames only: no magnitudes, bins, smoothing or beat "
                                            "stages -- what any CPU chain on this FFT could at most reach" % nfr})
    except Exception as e:  # a missing library must not cost the bench line
        out["others"].append({"fft_provider": "library FFT line unavailable", "error": str(e)})
    best = max([out] + [o for o in out["others"] if "value" in o and o.get("kind") == "port"], key=lambda o: o["value"])
    out["best"] = {"value": best["value"], "fft_provider": best["fft_provider"], "cores": best["cores"]}
    return out



SHARD_CHANNELS, SHARD_SECONDS = 8, 7200   # configs[4] in small: 8 channels x 2 h (1000 h is the same loop run longer)


def _bit_hash(t):
    """Order-sensitive 64-bit checksum of a tensor's raw 32-bit words (wraps mod 2^64), computed on its device."""
    import torch
    w = t.contiguous().view(torch.int32).reshape(-1).to(torch.int64)
    idx = torch.arange(w.numel(), device=w.device, dtype=torch.int64)
    return int(((w & 0xffffffff) * (idx % 1000003 + 1)).sum().item())


def sharded_leg(F, sp, dev, rank, world, barrier, max_over_ranks, steps):
    """One fixed 8-channel job, frame-range sharded over the ranks: no data-path collective.
    Every rank generates ONLY its own sample span on its device and runs it with the job's absolute frame numbers.
    Rank 0 also runs the whole job alone: the 1-GPU time of the same job and the single-GPU result every shard is
    compared with bit for bit (order-sensitive checksums of whole shards + the frames either side of every cut)."""
    import torch
    import torch.distributed as dist
    from fft_lines_b200 import sharding as sh
    S = FS * SHARD_SECONDS
    total_frames = sp.num_frames(S)
    stream = torch.cuda.current_stream(dev)
    EDGE = 8

    def timed(fn, k):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(k):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / k

    # ---- the same job on one GPU (rank 0 alone; the other ranks wait at the barrier)
    t1 = 0.0
    full_hash, full_edges = None, None
    if rank == 0:
        pcm = torch.empty((SHARD_CHANNELS, S), dtype=torch.int16, device=dev)
        F.synth_pcm_device(pcm)
        fbars, _, fbeat = sp.alloc_device_outputs(SHARD_CHANNELS, total_frames, dev)
        t1 = timed(lambda: sp.run_device(pcm, fbars, None, fbeat), steps)
        del pcm
        full_hash, full_edges = [], []
        for r in range(world):
            b, e = sh.shard_frames(total_frames, world, r)
            full_hash.append((_bit_hash(fbars[:, b:e]), _bit_hash(fbeat[:, b:e])))
            full_edges.append((fbars[:, b:b + EDGE].clone(), fbars[:, e - EDGE:e].clone(), fbeat[:, b:b + EDGE].clone(), fbeat[:, e - EDGE:e].clone()))
        del fbars, fbeat
        torch.cuda.empty_cache()
    barrier()
    t1 = max_over_ranks(t1)

    # ---- this rank's shard: its sample span only
    fb, fe = sh.shard_frames(total_frames, world, rank)
    s0, s1 = sp.sample_span(S, fb, fe, True)
    assert (s0, s1) == sh.sample_span(fb, fe, N_FFT, HOP, True, S)
    span = torch.empty((SHARD_CHANNELS, s1 - s0), dtype=torch.int16, device=dev)
    F.synth_pcm_device(span, first_sample=s0)
    bars, _, beat = sp.alloc_device_outputs(SHARD_CHANNELS, fe - fb, dev)
    barrier()
    tn = max_over_ranks(timed(lambda: sp.run_device_span(span, s0, S, bars, None, beat, fb, fe), steps))
    mine = torch.tensor([_bit_hash(bars), _bit_hash(beat)], dtype=torch.int64, device=dev)
    edges = [bars[:, :EDGE].contiguous(), bars[:, -EDGE:].contiguous(), beat[:, :EDGE].contiguous(), beat[:, -EDGE:].contiguous()]
    ok_hash = ok_edges = True
    if world > 1:
        hashes = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(hashes, mine)
        gathered = []
        for t in edges:
            parts = [torch.empty_like(t) for _ in range(world)] if rank == 0 else None
            dist.gather(t, parts, dst=0)
            gathered.append(parts)
    else:
        hashes, gathered = [mine], [[t] for t in edges]
    if rank == 0:
        for r in range(world):
            ok_hash &= (int(hashes[r][0]), int(hashes[r][1])) == full_hash[r]
            ok_edges &= all(torch.equal(gathered[q][r], full_edges[r][q]) for q in range(4))
    # ---- optional final gather of the outputs over NCCL (not part of the path; timed on its own)
    gather = None
    if world > 1:
        def g():
            return sh.gather_frames(bars, total_frames)
        g()
        torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        allb = g()
        e1.record(stream)
        torch.cuda.synchronize()
        gms = max_over_ranks(e0.elapsed_time(e1))
        nbytes = allb.numel() * 4
        gather = {"what": "all-gather of the float bars onto every rank (NCCL, sharding.gather_frames)", "ms": gms,
                  "bytes": int(nbytes), "gb_per_s_per_rank": nbytes / (gms * 1e-3) / 1e9}
        del allb
    frames = SHARD_CHANNELS * total_frames
    halo = int(max_over_ranks(float(SHARD_CHANNELS * (fb - s0 // HOP) if fe > fb else 0)))
    rec = {"job": "%d channels x %d s, 4096-pt Hann, hop 512, 200 bars + beat: %d frames, frame-range sharded over %d GPU(s)"
                  % (SHARD_CHANNELS, SHARD_SECONDS, frames, world),
           "api": "fftl_stft_run_device_span (each rank holds only sharding.sample_span of the PCM, generated on its own device)",
           "frames": frames, "ms_one_gpu": t1, "ms_sharded": tn, "frames_per_s": frames / (tn * 1e-3),
           "speedup": t1 / tn, "efficiency": t1 / (world * tn), "scaling": "strong",
           "halo_frames_recomputed_per_rank_max": int(halo), "collective_in_path": None,
           "bit_exact_vs_one_gpu": bool(ok_hash), "rank_boundaries_bit_exact": bool(ok_edges),
           "boundary_check": "first and last %d frames of every shard (bars and beat records) torch.equal to the single-GPU run; "
                             "order-sensitive 64-bit checksums of every whole shard equal" % EDGE,
           "gather": gather}
    del span, bars, beat
    torch.cuda.empty_cache()
    return rec


def sustained_leg(sp, pcm, bars, beat, dev, local_rank, frames_per_step, world, barrier, max_over_ranks, seconds=3.0):
    """The step back to back for >= `seconds` with the clock sampler running: what the kernel holds, not a burst."""
    import torch
    stream = torch.cuda.current_stream(dev)
    sampler = ClockSampler(physical_gpu_index(local_rank), power=True)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    n = 0
    while True:
        for _ in range(50):
            sp.run_device(pcm, bars, None, beat)
        n += 50
        torch.cuda.synchronize()
        if time.perf_counter() - t0 >= seconds:
            break
    e1.record(stream)
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    per_step = max_over_ranks(ms / n)
    return {"seconds": ms * 1e-3, "steps": n, "ms_per_step": per_step, "value": world * frames_per_step / (per_step * 1e-3),
            "unit": "frames/s", "clocks": clocks}


def latency_leg(F, dev):
    """BASELINE configs[2]: 1024 concurrent mono streams, 1024-point window, 64 bars; one tick (512 new samples per
    stream) per batch through the streaming front end (windows and beat state resident on the device)."""
    import numpy as np
    import torch
    STREAMS, COUNT, B = 1024, 512, 64
    cfg3 = F.extended_config(1024, COUNT, B)
    src = torch.empty((STREAMS, COUNT * 8), dtype=torch.int16, device=dev)
    F.synth_pcm_device(src)
    pct = lambda a: {"p50_us": float(np.percentile(a, 50)), "p99_us": float(np.percentile(a, 99))}
    with F.StreamingSpectrum(cfg3, STREAMS, device=dev.index) as ss:
        fresh = src[:, :COUNT].contiguous()
        bars = torch.empty((STREAMS, B), dtype=torch.float32, device=dev)
        beat = torch.empty((STREAMS, 8), dtype=torch.int32, device=dev)
        st = torch.cuda.Stream(dev)
        with torch.cuda.stream(st):
            l0 = ss.launch_count
            for _ in range(5):
                ss.push_device(fresh, bars, None, beat, stream=st)
            st.synchronize()
            per_tick = (ss.launch_count - l0) // 5
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                ss.push_device(fresh, bars, None, beat, stream=st)
            lat = []
            for _ in range(1500):
                t0 = time.perf_counter()
                g.replay()
                st.synchronize()
                lat.append((time.perf_counter() - t0) * 1e6)
            lat_graph = np.array(lat[300:])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(500):
                g.replay()
            e1.record(st)
            st.synchronize()
            dev_us = e0.elapsed_time(e1) * 1e3 / 500
            # eight ticks per graph: the graph-launch gap (about 2 us between replays) is spread over eight ticks
            g8 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g8, stream=st):
                for _ in range(8):
                    ss.push_device(fresh, bars, None, beat, stream=st)
            for _ in range(20):
                g8.replay()
            e0.record(st)
            for _ in range(100):
                g8.replay()
            e1.record(st)
            st.synchronize()
            dev_us8 = e0.elapsed_time(e1) * 1e3 / 800
        # the three-launch form of a push (window move, batched bars kernel, beat stage) timed the same way, for comparison
        dev_us_three = None
        try:
            with F.StreamingSpectrum(cfg3, STREAMS, device=dev.index) as s3:
                s3.set_tick_kernel(F.TICK_THREE_LAUNCHES)
                with torch.cuda.stream(st):
                    for _ in range(5):
                        s3.push_device(fresh, bars, None, beat, stream=st)
                    st.synchronize()
                    g3 = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g3, stream=st):
                        s3.push_device(fresh, bars, None, beat, stream=st)
                    for _ in range(50):
                        g3.replay()
                    e0.record(st)
                    for _ in range(500):
                        g3.replay()
                    e1.record(st)
                    st.synchronize()
                    dev_us_three = e0.elapsed_time(e1) * 1e3 / 500
                    del g3
        except Exception as ex:  # a comparison figure only: never fails the bench
            sys.stderr.write("latency: three-launch comparison skipped (%s)\n" % ex)
        h_fresh = F.PinnedArray((STREAMS, COUNT), np.int16)
        h_fresh.array[:] = fresh.cpu().numpy()
        hb, hbt = F.PinnedArray((STREAMS, B), np.float32), F.PinnedArray((STREAMS,), F.BEAT_DTYPE)
        o = {"bars": hb.array, "beat": hbt.array}
        for _ in range(20):
            ss.push(h_fresh.array, want_i32=False, out=o)
        lat = []
        for _ in range(800):
            t0 = time.perf_counter()
            ss.push(h_fresh.array, want_i32=False, out=o)
            lat.append((time.perf_counter() - t0) * 1e6)
        lat_host = np.array(lat[100:])
        h_fresh.free(); hb.free(); hbt.free()
    return {"workload": "c3: 1024 concurrent mono streams, 1024-pt Hann, 64 log bars + SG + beat, one tick (512 new samples per stream) per batch",
            "frames_per_batch": STREAMS, "kernels_per_tick": int(per_tick),
            "device_us_per_batch": dev_us, "device_us_per_batch_8_ticks_per_graph": dev_us8, "device_us_per_batch_three_launches": dev_us_three, "frames_per_s_device": STREAMS / dev_us * 1e6,
            "tick_kernel": "stream_tick_kernel<1024> (one kernel per tick: csrc/stream_tick.cuh)" if per_tick == 1 else "window move + bars kernel + beat stage",
            "graph_replay_plus_sync": pct(lat_graph), "host_buffer_call": pct(lat_host),
            "note": "device = CUDA-graph replays back to back (events); graph_replay_plus_sync = one replay + stream sync seen "
                    "from the host; host_buffer_call = fftl_stream_push with pinned host buffers (H2D + tick + D2H + sync)"}


def run_reference(args, rank, world):
    """`--impl reference`: the reference's CPU path for this config on the host cores."""
    if rank != 0:
        return
    O, cfg, pcm, threads, want = cpu_leg(None, target_seconds=4.0)
    frames = CHANNELS * O.num_frames(cfg, pcm.shape[1])
    for _ in range(max(args.warmup, 0)):
        O.port_run(cfg, pcm, threads=threads, want_i32=False)
    t = 0.0
    for _ in range(args.steps):
        sec, _ = O.port_run(cfg, pcm, threads=threads, want_i32=False)
        t += sec
    v = frames * args.steps / t
    sample = "%d s of the 1 h stereo workload per step (%d frames), fp32 port + fftwf-shim" % (want, frames)
    line = {
        "impl": "reference", "metric": "stft_frames_per_s", "value": v, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample,
                         "fft_provider": "fftwf-shim (no host FFTW in this image)"},
        "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seconds", type=int, default=SECONDS, help=argparse.SUPPRESS)
    ap.add_argument("--no-cpu-baseline", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-e2e", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--no-extras", action="store_true", help=argparse.SUPPRESS)  # profiling runs: skip sharded / sustained / latency
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import fft_lines_b200 as F

    if F.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; libfftlines_cuda has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = None
    if world > 1 and not os.environ.get("FFTL_NO_AFFINITY"):
        # Run this rank on the CPUs next to its GPU before any pinned buffer exists: the e2e leg streams
        # >1 GB per step between host and device, and first-touch then puts the pinned pages on the GPU's
        # own NUMA node instead of wherever torchrun happened to start the process.
        try:
            import pynvml
            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
            numa = "rank bound to its GPU's CPU set (nvmlDeviceSetCpuAffinity): %d cpus" % len(os.sched_getaffinity(0))
        except Exception as e:  # affinity is an optimisation, never a requirement
            numa = "not bound: %s" % e
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    S = FS * args.seconds
    cfg = F.extended_config(N_FFT, HOP, BARS, sample_rate=float(FS))
    sp = F.BatchedSpectrum(cfg, local_rank)
    nf = sp.num_frames(S)
    frames_per_step = CHANNELS * nf
    # every rank gets its own hour of audio (weak scaling): sample offset = rank * S
    pcm = torch.empty((CHANNELS, S), dtype=torch.int16, device=dev)
    F.synth_pcm_device(pcm, first_sample=rank * S)
    bars, _, beat = sp.alloc_device_outputs(CHANNELS, nf, dev)
    stream = torch.cuda.current_stream(dev)
    fp32_peak = F.measure_fp32_peak(5)

    # ---------------- device-resident: `value` ----------------
    for _ in range(max(args.warmup, 3)):
        sp.run_device(pcm, bars, None, beat)
    barrier()
    sp.timing()
    l0 = sp.launch_count
    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        sp.run_device(pcm, bars, None, beat)
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = sp.launch_count - l0
    bars_kernel_name = ("stft_bars_rdif_tm_kernel<4096> (fast path, per-thread constants in tensor memory; csrc/stft_rdif_tm.cuh)" if sp.tmem_launch_count > 0 else
                        "stft_bars_rdif_kernel<4096> (fast path, register-resident constants)" if sp.fast_launch_count > 0 else "stft_bars_kernel<4096> (generic)")
    calls, tot_ms, bars_ms = sp.timing()
    bars_kernel_ms = max_over_ranks(bars_ms / max(calls, 1))
    value = world * frames_per_step * args.steps / (ms_total * 1e-3)
    checksum = float(bars[:, ::997].double().sum().item())  # the result is real and finite
    assert math.isfinite(checksum) and checksum > 0

    # ---------------- end to end through the C-ABI host call ----------------
    e2e = None
    if not args.no_e2e:
        h_pcm = F.PinnedArray((CHANNELS, S), np.int16)
        h_bars = F.PinnedArray((CHANNELS, nf, BARS), np.float32)
        h_beat = F.PinnedArray((CHANNELS, nf), F.BEAT_DTYPE)
        torch.cuda.synchronize()
        pin_t = torch.from_numpy(h_pcm.array)
        pin_t.copy_(pcm)  # device -> pinned host, outside the timed region
        torch.cuda.synchronize()
        out = {"bars": h_bars.array, "beat": h_beat.array}
        e2e_steps = max(2, min(args.steps, 5))
        for _ in range(2):
            sp.run(h_pcm.array, want_i32=False, out=out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            sp.run(h_pcm.array, want_i32=False, out=out)  # returns when outputs are in host memory
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        dev_bars = bars.cpu().numpy()
        assert np.array_equal(h_bars.array, dev_bars), "host path and device path disagree"
        e2e = {"value": world * frames_per_step * e2e_steps / dt, "unit": "frames/s",
               "h2d_bytes_per_step": int(CHANNELS * S * 2),
               "d2h_bytes_per_step": int(CHANNELS * nf * BARS * 4 + CHANNELS * nf * F.BEAT_DTYPE.itemsize),
               "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
               "api": "fftl_stft_run (host int16 PCM -> host float bars + beat records, pinned buffers)"}
        if numa:
            e2e["host_affinity"] = numa
        # What the host <-> device links of this box allow for exactly these buffers, with every rank copying at once:
        # the step's input H2D and its outputs D2H as two bare concurrent copies (no kernels, no chunking).  The
        # end-to-end path cannot beat this; `frac_of_link_ceiling` says how close it gets at this N.
        s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        hb_t, hbeat_t = torch.from_numpy(h_bars.array), torch.from_numpy(h_beat.array.view(np.int32).reshape(CHANNELS, nf, -1))

        def raw_copies():
            with torch.cuda.stream(s_in):
                pcm.copy_(pin_t, non_blocking=True)
            with torch.cuda.stream(s_out):
                hb_t.copy_(bars, non_blocking=True)
                hbeat_t.copy_(beat, non_blocking=True)
        raw_copies()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            raw_copies()
        barrier()
        dt_raw = max_over_ranks(time.perf_counter() - t0) / 3
        moved = e2e["h2d_bytes_per_step"] + e2e["d2h_bytes_per_step"]
        e2e["link_ceiling"] = {"ms_per_step": 1e3 * dt_raw, "value": world * frames_per_step / dt_raw, "unit": "frames/s",
                               "aggregate_gb_per_s": world * moved / dt_raw / 1e9, "per_gpu_gb_per_s": moved / dt_raw / 1e9,
                               "what": "this step's H2D and D2H as two bare concurrent pinned copies on every rank at once"}
        e2e["frac_of_link_ceiling"] = e2e["value"] / e2e["link_ceiling"]["value"]
        h_pcm.free(); h_bars.free(); h_beat.free()

    # ---------------- sustained / sharded job / streaming latency ----------------
    sustained = sharded = latency = None
    if not args.no_extras:
        sustained = sustained_leg(sp, pcm, bars, beat, dev, local_rank, frames_per_step, world, barrier, max_over_ranks)
    del pcm, bars, beat
    torch.cuda.empty_cache()
    if not args.no_extras:
        sharded = sharded_leg(F, sp, dev, rank, world, barrier, max_over_ranks, max(3, min(args.steps, 10)))
        if rank == 0:
            latency = latency_leg(F, dev)
        barrier()

    # ---------------- CPU baselines beside it (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baselines()

    if rank == 0:
        peaks, how = measured_peaks()
        frames_per_launch = frames_per_step + CHANNELS * 0  # warm-up frames: frame_begin = 0, none
        flops = FLOPS_PER_FRAME * frames_per_launch
        achieved_tf = flops / (bars_kernel_ms * 1e-3) / 1e12
        achieved_gbs = BYTES_PER_FRAME * frames_per_launch / (bars_kernel_ms * 1e-3) / 1e9
        # dram bytes per launch come from an `ncu --set full` capture (never measured under the bench itself); the
        # record is stamped with a hash of the kernel sources it was taken from, and a stale one is not reported
        traffic, traffic_note = None, "no capture"
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            from fft_lines_b200.build import kernel_source_hash
            with open(tp) as f:
                tj = json.load(f)
            if tj.get("kernel_source_hash") == kernel_source_hash():
                traffic, traffic_note = tj.get("stft_bars_dram_bytes_per_launch"), "from %s" % tj.get("source")
            else:
                traffic_note = "stale: %s was captured from other kernel sources (hash %s, now %s)" % (
                    tj.get("source"), tj.get("kernel_source_hash"), kernel_source_hash())
        roof = {
            "bound": "fp32", "kernel": bars_kernel_name,
            "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp32_peak,
            "peak_source": "FP32 FMA probe run in this process (not in MEASURED_PEAKS.json); nominal %.2f" % NOMINAL_FP32_TFLOPS,
            "frac_of_nominal": achieved_tf / NOMINAL_FP32_TFLOPS,
            "flops_per_frame": FLOPS_PER_FRAME, "frames_per_launch": frames_per_launch,
            "kernel_ms": bars_kernel_ms,
            "hbm": {"achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved_gbs / peaks["hbm_gbs"], "bytes_per_frame": BYTES_PER_FRAME, "of": how},
            "traffic": traffic, "traffic_note": traffic_note,
        }
        line = {
            "metric": "stft_frames_per_s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n": N_FFT, "hop": HOP, "bars": BARS, "channels": CHANNELS,
                       "seconds": args.seconds, "frames_per_step_per_gpu": frames_per_step,
                       "sharding": "one batch per GPU, no data-path collective",
                       "l2": "input %d MB per step > 126 MB L2, no flush needed" % (CHANNELS * S * 2 // 1000000)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches) * world, "roofline": roof, "cpu_baseline": cpu,
            "sharded": sharded, "sustained": sustained, "latency": latency,
        }
        print(json.dumps(line), flush=True)
    sp.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
