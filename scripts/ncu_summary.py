#!/usr/bin/env python
"""Per-launch summary of an ncu --set full report: duration, tensor / XU pipe activity, DRAM bytes, issue activity, registers.
usage: ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "us"), ("sm__cycles_elapsed.avg.per_second", "GHz"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"), ("lts__t_sector_hit_rate.pct", "l2hit%"), ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"),
        ("launch__block_size", "block")]
print(f"# {rep}")
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    name = r[col["Kernel Name"]].split("(")[0]
    parts = []
    for key, label in want:
        if key in col:
            unit = rows[1][col[key]]
            v = r[col[key]]
            try:
                v = f"{float(v.replace(',', '')):.4g}"
            except ValueError:
                pass
            parts.append(f"{label}={v}{unit if label.startswith('dram_') or label == 'us' else ''}")
    print(name[:70], " ".join(parts))
