"""Warp-stall breakdown of one kernel from an `ncu --set full --import-source on` report: the per-issue stall ratios, the pipe
utilisations and the instructions that collect the most stall samples.  usage: python scripts/ncu_stalls.py report.ncu-rep [kernel substring]"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
sub = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if sub in d["Kernel Name"]:
        break
print("kernel:", d["Kernel Name"][:120])
print("duration_us:", d.get("gpu__time_duration.sum"), " cycles:", d.get("sm__cycles_elapsed.max"))
keys = ["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__warps_active.avg.per_cycle_active", "dram__bytes_read.sum", "dram__bytes_write.sum"]
for k in keys:
    print(f"  {k}: {d.get(k)}")
print("warp stalls per issued instruction (smsp__average_warps_issue_stalled_*_per_issue_active):")
for k, v in sorted(((k, float(v)) for k, v in d.items() if "average_warps_issue_stalled" in k and v), key=lambda kv: -kv[1]):
    print(f"  {k.split('stalled_')[1].replace('_per_issue_active.ratio', ''):22s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[start + 1:] if len(r) == len(hdr)]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print(f"stall samples: {tot} over {len(data)} SASS instructions; by reason:")
agg = collections.Counter()
for r in data:
    for h in stalls:
        agg[h] += int(r[ix[h]])
for h, n in agg.most_common(10):
    print(f"  {h:24s} {n:6d}  {100.0 * n / tot:5.1f} %")
print("top 25 instructions by samples:  [index] SASS | samples | executed | reasons")
for i, r in sorted(sorted(enumerate(data), key=lambda t: -int(t[1][ix["# Samples"]]))[:25]):
    s = {h[6:]: int(r[ix[h]]) for h in stalls if int(r[ix[h]]) > 0}
    print(f"  [{i:4d}] {r[ix['Source']][:64]:64s} {r[ix['# Samples']]:>6s} {r[ix['Instructions Executed']]:>8s} {s}")
mix = collections.Counter()
mx = max(int(r[ix["Instructions Executed"]]) for r in data)
for r in data:
    n = int(r[ix["Instructions Executed"]])
    t = r[ix["Source"]].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    mix[op] += n
tot_i = sum(mix.values())
print("executed warp instructions by opcode (all warps):", ", ".join(f"{k} {100.0 * v / tot_i:.1f}%" for k, v in mix.most_common(14)))
