"""Developer probe: end-to-end fftl_stft_run timing (pinned host buffers) for the c2 workload."""
import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import fft_lines_b200 as F
S = 48000 * 3600
cfg = F.extended_config(4096, 512, 200)
sp = F.BatchedSpectrum(cfg)
nf = sp.num_frames(S)
d = torch.empty((2, S), dtype=torch.int16, device="cuda"); F.synth_pcm_device(d)
h_pcm = F.PinnedArray((2, S), np.int16); torch.from_numpy(h_pcm.array).copy_(d); torch.cuda.synchronize()
h_bars = F.PinnedArray((2, nf, 200), np.float32); h_beat = F.PinnedArray((2, nf), F.BEAT_DTYPE)
out = {"bars": h_bars.array, "beat": h_beat.array}
for _ in range(2): sp.run(h_pcm.array, want_i32=False, out=out)
t0 = time.perf_counter()
K = 5
for _ in range(K): sp.run(h_pcm.array, want_i32=False, out=out)
dt = (time.perf_counter() - t0) / K
print("e2e %.2f ms/step -> %.1f M frames/s, %.1f GB/s aggregate" % (dt * 1e3, 2 * nf / dt / 1e6, (2 * S * 2 + 2 * nf * 200 * 4 + 2 * nf * 32) / dt / 1e9))
