"""Developer probe: pinned H2D / D2H bandwidth alone and concurrently (the ceiling of the e2e path)."""
import torch, time
a = torch.empty(691_000_000, dtype=torch.uint8).pin_memory(); b = torch.empty(562_000_000, dtype=torch.uint8).pin_memory()
da = torch.empty_like(a, device="cuda"); db = torch.empty_like(b, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, K=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(K):
        if h2d:
            with torch.cuda.stream(s1): da.copy_(a, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): b.copy_(db, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / K
for _ in range(2): run(True, True)
t = run(True, False); print("H2D alone  %.2f ms  %.1f GB/s" % (t * 1e3, a.numel() / t / 1e9))
t = run(False, True); print("D2H alone  %.2f ms  %.1f GB/s" % (t * 1e3, b.numel() / t / 1e9))
t = run(True, True); print("both       %.2f ms  %.1f GB/s aggregate" % (t * 1e3, (a.numel() + b.numel()) / t / 1e9))
