"""Developer helper: split an `ncu --page source --csv` dump into barrier-delimited phases."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
kern = []
cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "ins": []}
        kern.append(cur)
        continue
    if cur is None:
        continue
    if cur["hdr"] is None:
        cur["hdr"] = r
        continue
    if len(r) >= len(cur["hdr"]):
        cur["ins"].append(r)
seen = set()
for k in kern:
    if k["name"] in seen:
        continue
    seen.add(k["name"])
    hdr = k["hdr"]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    phase = 0
    acc = collections.defaultdict(collections.Counter)
    ninst = collections.Counter()
    ops = collections.defaultdict(collections.Counter)
    for r in k["ins"]:
        s = r[ix["Source"]].strip()
        op = re.sub(r"^@!?U?P\d+\s+", "", s).split()[0].split(".")[0]
        e = int(r[ix["Instructions Executed"]])
        for st in stalls:
            acc[phase][st] += int(r[ix[st]] or 0)
        ninst[phase] += e
        ops[phase][op] += e
        if op == "BAR":
            phase += 1
    tot_i = sum(ninst.values())
    tot_s = sum(sum(a.values()) for a in acc.values())
    print("==", k["name"][:70], "warp-instr", tot_i, "samples", tot_s)
    for ph in sorted(acc):
        t = sum(acc[ph].values())
        print(" phase %d: instr %5.1f%%  samples %5.1f%%  stalls %s" % (
            ph, 100.0 * ninst[ph] / tot_i, 100.0 * t / max(tot_s, 1),
            [(a.replace("stall_", ""), round(100 * b / max(t, 1))) for a, b in acc[ph].most_common(6)]))
        print("          ops %s" % [(a, round(100.0 * b / tot_i, 1)) for a, b in ops[ph].most_common(10)])
