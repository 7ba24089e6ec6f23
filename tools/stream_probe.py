"""Developer probe: a few ticks of the c3 streaming front end (for an ncu launch list)."""
import sys
sys.path.insert(0, ".")
import torch
import fft_lines_b200 as F
STREAMS, COUNT = 1024, 512
dev = torch.device("cuda")
src = torch.empty((STREAMS, COUNT), dtype=torch.int16, device=dev)
F.synth_pcm_device(src)
cfg3 = F.extended_config(1024, COUNT, 64, sg_half=3, sg_degree=2)
with F.StreamingSpectrum(cfg3, STREAMS) as ss:
    bars = torch.empty((STREAMS, 64), dtype=torch.float32, device=dev)
    beat = torch.empty((STREAMS, 8), dtype=torch.int32, device=dev)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(8):
            ss.push_device(src, bars, None, beat, stream=st)
        st.synchronize()
print("ok")
