"""One ncu capture (`--set full --import-source on`) of one kernel -> a tracked summary under profiles/: key counters
plus the barrier-delimited phase table of tools/phase_report.py (share of instructions / of warp samples, top stalls).

    python tools/summarize_kernel.py gpurun_out/r02_n8192.ncu-rep profiles/r02_ncu_n8192.md "<command line that was profiled>"
"""
import csv
import io
import os
import subprocess
import sys

rep, out, title = sys.argv[1:4]
here = os.path.dirname(os.path.abspath(__file__))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
src_csv = rep + ".source.csv"
with open(src_csv, "w") as f:
    f.write(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout)
phases = subprocess.run([sys.executable, os.path.join(here, "phase_report.py"), src_csv], capture_output=True, text=True).stdout
os.remove(src_csv)
with open(out, "w") as f:
    f.write("# `ncu --set full --clock-control none --import-source on` of `%s`\n\n" % title)
    seen = set()
    for vals in rows[2:]:
        d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
        name = d["Kernel Name"][0]
        if name in seen:
            continue
        seen.add(name)
        f.write("## `%s`\n\n| counter | value | unit |\n|---|---|---|\n" % name[:110])
        for k in keys:
            if k in d:
                f.write("| %s | %s | %s |\n" % (k, d[k][0], d[k][1]))
        f.write("\n")
    f.write("## Phases (SASS between consecutive `BAR` instructions, in SASS order; share of executed warp instructions, share of "
            "warp-state samples, leading stall reasons in % of the phase's samples, leading opcodes in % of all instructions)\n\n```\n")
    f.write(phases)
    f.write("```\n")
print("wrote", out)
