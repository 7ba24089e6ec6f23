"""Turns the ncu artefacts of one gpurun call into the tracked summaries under profiles/.

    python tools/summarize_profile.py r01 gpurun_out/launches_r01.csv gpurun_out/prof_r01.ncu-rep
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

tag, launches_csv, rep = sys.argv[1:4]
out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list: per-kernel share of the step (cold-cache, serialised: compare shares, not absolutes)
rows = [r for r in csv.reader(l for l in open(launches_csv) if not l.startswith("=="))]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
agg = collections.OrderedDict()
lines = []
for r in rows[1:]:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = r[ix["Kernel Name"]].split("(")[0]
    val = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = val / 1000.0 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1000.0)
    agg.setdefault(name, []).append(us)
    lines.append((r[ix["ID"]], name, us))
total = sum(sum(v) for v in agg.values())
with open(os.path.join(out_dir, "%s_launches.md" % tag), "w") as f:
    f.write("# %s: ncu launch list (`--metrics gpu__time_duration.sum --clock-control none`) of `python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e`\n\n" % tag)
    f.write("Per-launch times under ncu are cold-cache and serialised: the SHARE of the step is what is comparable.\n\n")
    f.write("| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write("| `%s` | %d | %.1f | %.1f | %.1f%% |\n" % (k, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / total))
    f.write("\nFirst 40 launches in order:\n\n| id | kernel | us |\n|---|---|---|\n")
    for i, n, us in lines[:40]:
        f.write("| %s | `%s` | %.1f |\n" % (i, n, us))

# ---- full capture: key counters per kernel
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
summary = {}
with open(os.path.join(out_dir, "%s_ncu_full.md" % tag), "w") as f:
    f.write("# %s: `ncu --set full --clock-control none --import-source on` of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e`\n\n" % tag)
    seen = set()
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        name = d["Kernel Name"][0].split("(")[0]
        if name in seen:
            continue
        seen.add(name)
        f.write("## `%s`\n\n| counter | value | unit |\n|---|---|---|\n" % d["Kernel Name"][0][:100])
        entry = {}
        for k in keys:
            if k in d:
                f.write("| %s | %s | %s |\n" % (k, d[k][0], d[k][1]))
                entry[k] = d[k]
        f.write("\n")
        summary[name] = entry
    f.write("FP32 pipe: `sm__pipe_fma_cycles_active` is the FMA-pipe busy fraction; tensor pipes are unused by design (the FFT is not a dense contraction).\n")

# traffic per launch of the dominant kernel, read by bench.py
for name, e in summary.items():
    if "rdif" in name or "stft_bars" in name:
        def mb(x):
            v, u = x
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        t = mb(e["dram__bytes_read.sum"]) + mb(e["dram__bytes_write.sum"])
        with open(os.path.join(out_dir, "traffic.json"), "w") as f:
            sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            from fft_lines_b200.build import kernel_source_hash
            json.dump({"stft_bars_dram_bytes_per_launch": t, "source": "profiles/%s_ncu_full.md" % tag, "kernel": name,
                       "kernel_source_hash": kernel_source_hash()}, f, indent=1)
        break
print("wrote profiles/%s_*" % tag)
