"""Developer helper: build A/B variants of the library with extra -D flags into build_ab/ (git-ignored; travels to the
GPU box with the snapshot) and print the shell loop that times them with tools/quick_bench.py.

    python tools/ab_build.py base= v1=-DFFTL_X_BEATWARP v2=-DFFTL_X_BEATWARP,-DFFTL_X_PREFETCH
"""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fft_lines_b200 import build as B

out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build_ab")
os.makedirs(out_dir, exist_ok=True)
jobs = []
for arg in sys.argv[1:]:
    name, _, flags = arg.partition("=")
    jobs.append((name, [f for f in flags.split(",") if f]))


def one(job):
    name, flags = job
    return B.build(force=True, extra=flags, out=os.path.join(out_dir, "lib_%s.so" % name), verbose=False)


with ThreadPoolExecutor(max_workers=2) as pool:
    for path in pool.map(one, jobs):
        print(path)
