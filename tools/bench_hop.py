"""quick_bench with a hop argument: python tools/bench_hop.py <hop> [interleaved]"""
import sys, time, torch
sys.path.insert(0, ".")
import fft_lines_b200 as F
hop = int(sys.argv[1]); inter = len(sys.argv) > 2 and sys.argv[2] == 'i'
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
bars, ch = (200 if n == 4096 else 100), 2
S = 48000 * 3600
cfg = F.extended_config(n, hop, bars)
pcm = torch.empty((ch, S), dtype=torch.int16, device="cuda")
F.synth_pcm_device(pcm)
if inter:
    pcm = pcm.t().contiguous().t()   # [ch][S] view of interleaved storage
torch.cuda.synchronize()
sp = F.BatchedSpectrum(cfg)
nf = sp.num_frames(S)
out_bars, _, beat = sp.alloc_device_outputs(ch, nf, "cuda")
for _ in range(3):
    sp.run_device(pcm, out_bars, None, beat)
torch.cuda.synchronize(); sp.timing()
for _ in range(5):
    sp.run_device(pcm, out_bars, None, beat)
torch.cuda.synchronize()
calls, tot, bk = sp.timing()
print("hop=%d inter=%s frames=%d bars-kernel %.3f ms -> %.1f M frames/s (tmem launches %d)" % (hop, inter, ch * nf, bk / calls, ch * nf / (bk / calls) / 1e3, sp.tmem_launch_count))
