import sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import fft_lines_b200 as lib
from oracle import oracle
from helpers import assert_bars_close, assert_int_heights
rng0 = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
fails = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 200):
    n = int(rng0.choice([256, 512, 1024, 2048, 4096])); seed = int(rng0.integers(0, 2**31 - 1)); shape = str(rng0.choice(["uniform", "sparse", "loud", "tiny", "square"]))
    rng = np.random.default_rng(seed)
    frames, hop = 6, n // 4
    S = n + (frames - 1) * hop
    if shape == "uniform": x = rng.integers(-32768, 32768, S)
    elif shape == "sparse": x = np.where(rng.random(S) < 0.01, rng.integers(-32768, 32768, S), 0)
    elif shape == "loud": x = rng.choice([-32768, 32767], S)
    elif shape == "tiny": x = rng.integers(-2, 3, S)
    else: x = np.where((np.arange(S) // int(rng.integers(1, 40))) % 2 == 0, 12345, -12345)
    pcm = x.astype(np.int16)[None, :]
    B = n // 2 + 1
    cfg = lib.reference_config(n, hop, B, zoom=1e-3)
    with lib.BatchedSpectrum(cfg) as sp:
        got = sp.run(pcm, want_beat=False)
    ref = np.empty((1, frames, B))
    for f in range(frames):
        X = oracle.dft_naive(pcm[0, f * hop:f * hop + n].astype(np.float64))
        ref[0, f] = np.abs(X[:B]) * float(np.float32(1e-3))
    try:
        assert_bars_close(got["bars"], ref, "n=%d %s" % (n, shape))
        assert_int_heights(got["bars_i32"], ref.astype(np.int64).astype(np.int32), ref, "n=%d %s" % (n, shape))
    except AssertionError as e:
        fails += 1
        print("FAIL n=%d seed=%d shape=%s: %s" % (n, seed, shape, str(e)[:300]))
print("done, fails", fails)
