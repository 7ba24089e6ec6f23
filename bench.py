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
  cpu_baseline  the CPU port of the reference chain on this box's host cores

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

    def __init__(self, index):
        self.index = index
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
            out["samples"] = len(s)
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
        h_pcm.free(); h_bars.free(); h_beat.free()

    # ---------------- CPU baseline beside it (rank 0, N=1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        O, ocfg, cpcm, threads, want = cpu_leg(None)
        cframes = CHANNELS * O.num_frames(ocfg, cpcm.shape[1])
        sec, _ = O.port_run(ocfg, cpcm, threads=threads, want_i32=False)
        cpu = {"value": cframes / sec, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "first %d s of the same workload (%d frames), fp32 port of the reference chain + fftwf-shim" % (want, cframes),
               "seconds": sec}

    if rank == 0:
        peaks, how = measured_peaks()
        frames_per_launch = frames_per_step + CHANNELS * 0  # warm-up frames: frame_begin = 0, none
        flops = FLOPS_PER_FRAME * frames_per_launch
        achieved_tf = flops / (bars_kernel_ms * 1e-3) / 1e12
        achieved_gbs = BYTES_PER_FRAME * frames_per_launch / (bars_kernel_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get("stft_bars_dram_bytes_per_launch")
        roof = {
            "bound": "fp32", "kernel": "stft_bars_rdif_kernel<4096> (fast path; generic fallback: stft_bars_kernel<N>)",
            "achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp32_peak,
            "peak_source": "FP32 FMA probe run in this process (not in MEASURED_PEAKS.json); nominal %.2f" % NOMINAL_FP32_TFLOPS,
            "frac_of_nominal": achieved_tf / NOMINAL_FP32_TFLOPS,
            "flops_per_frame": FLOPS_PER_FRAME, "frames_per_launch": frames_per_launch,
            "kernel_ms": bars_kernel_ms,
            "hbm": {"achieved": achieved_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved_gbs / peaks["hbm_gbs"], "bytes_per_frame": BYTES_PER_FRAME, "of": how},
            "traffic": traffic,
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
        }
        print(json.dumps(line), flush=True)
    sp.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
