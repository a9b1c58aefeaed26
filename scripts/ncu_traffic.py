"""DRAM traffic per launch of the dominant kernel family from an `ncu --set full` report -> profiles/top_kernel_traffic.json (read by
bench.py for roofline.traffic).  usage: ncu_traffic.py report.ncu-rep family kernel_regex out.json"""
import csv, io, json, re, subprocess, sys
rep, family, rx, out = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot, n = 0.0, 0
for r in rows[2:]:
    if len(r) < len(hdr) or not re.search(rx, r[col["Kernel Name"]]):
        continue
    b = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        b += float(r[col[k]].replace(",", "")) * mult[units[col[k]]]
    tot += b
    n += 1
d = json.load(open(out)) if len(sys.argv) > 5 and sys.argv[5] == "merge" else {}
d[family] = tot / max(n, 1)
d["_source"] = f"{rep.split('/')[-1]}: mean dram__bytes_read.sum + dram__bytes_write.sum over {n} captured launches matching /{rx}/ (ncu --set full --clock-control none, scripts/ncu_step.py)"
json.dump(d, open(out, "w"), indent=1)
print(d)
