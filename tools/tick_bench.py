"""Developer timing probe of the c3 streaming tick (1024 streams x 1024-pt): device time per tick by CUDA-graph replay."""
import sys
import time
sys.path.insert(0, ".")
import torch
import fft_lines_b200 as F

STREAMS, COUNT, B = 1024, 512, 64
dev = torch.device("cuda")
src = torch.empty((STREAMS, COUNT), dtype=torch.int16, device=dev)
F.synth_pcm_device(src)
cfg3 = F.extended_config(1024, COUNT, B)
for name, which in (("one kernel per tick", F.TICK_AUTO), ("three launches per tick", F.TICK_THREE_LAUNCHES)):
    with F.StreamingSpectrum(cfg3, STREAMS) as ss:
        ss.set_tick_kernel(which)
        bars = torch.empty((STREAMS, B), dtype=torch.float32, device=dev)
        beat = torch.empty((STREAMS, 8), dtype=torch.int32, device=dev)
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for _ in range(5):
                ss.push_device(src, bars, None, beat, stream=st)
            st.synchronize()
            # TICKS pushes per graph: the host's share of a replay (one cudaGraphLaunch) is spread over them
            for TICKS in (1, 8):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=st):
                    for _ in range(TICKS):
                        ss.push_device(src, bars, None, beat, stream=st)
                for _ in range(50):
                    g.replay()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                st.synchronize()
                e0.record(st)
                h0 = time.perf_counter()
                for _ in range(1000 // TICKS):
                    g.replay()
                h1 = time.perf_counter()   # the host's share: how fast it can submit replays at all
                e1.record(st)
                st.synchronize()
                print("%s: %.2f us per tick (device, graph of %d tick(s) replayed back to back; host submits one replay per %.2f us)"
                      % (name, e0.elapsed_time(e1) * 1e3 / (1000 // TICKS * TICKS), TICKS, (h1 - h0) * 1e6 / (1000 // TICKS)))

# the floor of this way of measuring: a graph of one (or eight) empty-ish kernel(s) replayed back to back
one = torch.zeros(1, device=dev)
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for TICKS in (1, 8):
        one.add_(1.0)
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(TICKS):
                one.add_(1.0)
        for _ in range(50):
            g.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(1000 // TICKS):
            g.replay()
        e1.record(st)
        st.synchronize()
        print("one-element kernel: %.2f us per launch (graph of %d replayed back to back)" % (e0.elapsed_time(e1) * 1e3 / (1000 // TICKS * TICKS), TICKS))
