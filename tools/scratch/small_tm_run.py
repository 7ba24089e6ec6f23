"""Small run of the tensor-memory fast path for compute-sanitizer (racecheck / memcheck / synccheck)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import fft_lines_b200 as lib
from fft_lines_b200 import _capi
from oracle import oracle
for n, hop, bars in ((4096, 512, 200), (2048, 512, 100)):
    cfg = lib.extended_config(n, hop, bars)
    pcm = oracle.synth_pcm(2, n + 1500 * hop + 3)
    with lib.BatchedSpectrum(cfg) as sp:
        a = sp.run(pcm)
        assert sp.tmem_launch_count > 0
        sp.set_kernel(_capi.KERNEL_FAST_REGISTERS)
        b = sp.run(pcm)
    print(n, "equal:", bool((a["bars"] == b["bars"]).all()), "frames", a["bars"].shape)
