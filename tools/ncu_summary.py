"""Writes a text summary of one ncu report (key raw metrics + per-source-line attribution of
issued instructions / stall samples) for profiles/.  Dev container only (reads .ncu-rep files).
usage: python tools/ncu_summary.py report.ncu-rep kernel_substring out.txt"""
import csv
import subprocess
import sys

rep, kname, out = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
vals = rows[2]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
lines = [f"ncu report: {rep}", ""]
for h, u, v in zip(hdr, units, vals):
    if h in KEYS or ("issue_stalled" in h and "per_issue_active" in h and float(v or 0) >= 0.05):
        lines.append(f"{h:78s} {v} {u}")
lines.append("")
lines.append("per-source-line attribution (tools/ncu_lines.py):")
cubin = sys.argv[4] if len(sys.argv) > 4 else "/tmp/cub/engine.sm_100a.cubin"
att = subprocess.run([sys.executable, "tools/ncu_lines.py", rep, cubin, kname, "40"], capture_output=True, text=True).stdout
lines.append(att)
open(out, "w").write("\n".join(lines))
print("\n".join(lines[:40]))
