import sys
sys.path.insert(0, ".")
import torch, numpy as np
import fft_lines_b200 as F
cfg = F.extended_config(4096, 512, 200)
S = 48000 * 3600
pcm = torch.empty((2, S), dtype=torch.int16, device="cuda")
F.synth_pcm_device(pcm)
sp = F.BatchedSpectrum(cfg)
nf = sp.num_frames(S)
bars, _, beat = sp.alloc_device_outputs(2, nf, "cuda")
sp.run_device(pcm, bars, None, beat)
torch.cuda.synchronize()
b0, b1 = 200000, 200000 + 4099
pb, _, pbeat = sp.alloc_device_outputs(2, b1 - b0, "cuda")
sp.run_device(pcm, pb, None, pbeat, b0, b1)
torch.cuda.synchronize()
d = (pb != bars[:, b0:b1])
print("bars neq", int(d.sum()), d.nonzero()[:5].tolist())
a = F.BatchedSpectrum.beat_to_numpy(pbeat); b = F.BatchedSpectrum.beat_to_numpy(beat[:, b0:b1].contiguous())
for k in a.dtype.names:
    ne = (a[k] != b[k])
    print(k, int(ne.sum()), np.argwhere(ne)[:5].tolist())
