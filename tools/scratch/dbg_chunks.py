import sys, numpy as np, torch
sys.path.insert(0, ".")
import fft_lines_b200 as lib
cfg = lib.extended_config(4096, 512, 200)
frames = 48000
d = torch.empty((2, 4096 + (frames - 1) * 512), dtype=torch.int16, device="cuda")
lib.synth_pcm_device(d)
torch.cuda.synchronize()
pcm = d.cpu().numpy()
with lib.BatchedSpectrum(cfg) as sp:
    host = sp.run(pcm, 1000, frames, want_i32=False)
    bars, _, beat = sp.alloc_device_outputs(2, frames - 1000, "cuda", want_i32=False)
    sp.run_device(d, bars, None, beat, 1000, frames)
    torch.cuda.synchronize()
    b2 = torch.empty_like(bars)
    sp.run_device(d, b2, None, beat, 1000, frames)
    torch.cuda.synchronize()
print("device twice equal:", bool((bars == b2).all()))
hb = host["bars"]; db = bars.cpu().numpy()
bad = np.argwhere((hb != db).any(axis=2))
print("mismatching (ch, frame) count", len(bad))
print(bad[:40].tolist())
if len(bad):
    c, f = bad[0]
    print("host", hb[c, f, :8], "dev", db[c, f, :8])
    fr = np.unique(bad[:, 1]); print("frames rel", fr[:60].tolist())
