#!/usr/bin/env python
"""Summarise an .ncu-rep (via `ncu -i ... --page raw --csv`) into a short text file for profiles/."""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
KEYS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__inst_executed.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_ld.sum",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_local_op_st.sum",
]
with open(out, "w") as f:
    f.write(f"# summary of {rep} (ncu --set full --clock-control none); one column per captured launch\n")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            f.write(f"{k:75s} [{units[i]}] " + " | ".join(r[i] for r in data) + "\n")
    f.write("\n# warp stall sampling (smsp__pcsamp_warps_issue_stalled_*), first launch, share of samples\n")
    items = []
    for i, h in enumerate(hdr):
        if "pcsamp_warps_issue_stalled" in h and not h.endswith("_not_issued"):
            try:
                items.append((float(data[0][i].replace(",", "")), h))
            except ValueError:
                pass
    tot = sum(v for v, _ in items) or 1
    for v, h in sorted(items, reverse=True)[:10]:
        f.write(f"{h.replace('smsp__pcsamp_warps_issue_stalled_', ''):30s} {v / tot:6.1%}\n")
def _num(key):
    if key not in hdr:
        return None
    i = hdr.index(key)
    v = float(data[0][i].replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(units[i], 1)
    return v * scale


if len(sys.argv) > 3:  # also write the per-launch DRAM traffic bench.py reports as roofline.traffic
    import json

    json.dump({"dram_bytes_read": _num("dram__bytes_read.sum"), "dram_bytes_write": _num("dram__bytes_write.sum"),
               "fmaheavy_pct_of_peak": _num("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"),
               "alu_pct": _num("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
               "source": out}, open(sys.argv[3], "w"), indent=1)
print(open(out).read())
