"""Developer helper: static SASS mnemonic counts of one kernel of libfftlines_cuda.so (cuobjdump -sass), as the
markdown table kept in profiles/.   python tools/sass_summary.py 'stft_bars_rdif_tm_kernel<4096, 2, true, 3, 7>'"""
import collections
import os
import re
import subprocess
import sys

lib = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fft_lines_b200", "libfftlines_cuda.so")
want = sys.argv[1]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
counts, name, regs = collections.Counter(), None, None
for block in txt.split("Function : ")[1:]:
    mangled = block.split("\n", 1)[0].strip()
    dem = subprocess.run(["c++filt", mangled], capture_output=True, text=True).stdout.strip()
    if want not in dem:
        continue
    name = dem
    for line in block.splitlines():
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m:
            counts[m.group(1)] += 1
    break
if name is None:
    sys.exit("kernel not found: " + want)
print("`%s`: %d instructions" % (name.split("(")[0], sum(counts.values())))
print()
print("| mnemonic | static count |")
print("|---|---|")
for k, v in sorted(counts.items(), key=lambda kv: -kv[1]):
    print("| %s | %d |" % (k, v))
