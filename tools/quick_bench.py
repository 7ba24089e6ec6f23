"""Developer timing probe (not the contract bench): device-resident c2-shaped run."""
import sys
import time

import torch

sys.path.insert(0, ".")
import fft_lines_b200 as F

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
bars = int(sys.argv[2]) if len(sys.argv) > 2 else 200
secs = int(sys.argv[3]) if len(sys.argv) > 3 else 3600
mode = sys.argv[4] if len(sys.argv) > 4 else "x"
ch = 2
S = 48000 * secs
cfg = F.extended_config(n, 512, bars) if mode == "x" else F.reference_config(n, 512, bars)
pcm = torch.empty((ch, S), dtype=torch.int16, device="cuda")
F.synth_pcm_device(pcm)
torch.cuda.synchronize()
sp = F.BatchedSpectrum(cfg)
nf = sp.num_frames(S)
out_bars, _, beat = sp.alloc_device_outputs(ch, nf, "cuda")
for _ in range(3):
    sp.run_device(pcm, out_bars, None, beat)
torch.cuda.synchronize()
sp.timing()
K = 5
t0 = time.time()
for _ in range(K):
    sp.run_device(pcm, out_bars, None, beat)
torch.cuda.synchronize()
wall = (time.time() - t0) / K
calls, tot, bk = sp.timing()
# order-independent fingerprint of the outputs: A/B builds that claim bit-identical results must print the same one
fp = int(out_bars.view(torch.int32).to(torch.int64).sum().item()) & 0xffffffffffff
fp ^= (int(beat.to(torch.int64).sum().item()) & 0xffffffffffff) << 1   # the beat records too
print("n=%d bars=%d mode=%s frames=%d  fp %012x  total %.3f ms  bars-kernel %.3f ms  wall %.3f ms  -> %.1f M frames/s (bars kernel %.1f M/s)"
      % (n, bars, mode, ch * nf, fp, tot / calls, bk / calls, wall * 1e3, ch * nf / (tot / calls) / 1e3, ch * nf / (bk / calls) / 1e3))
