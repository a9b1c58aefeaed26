"""Per-kernel counts of the Blackwell-specific SASS instructions in the built library (cuobjdump -sass): UTCHMMA (tcgen05.mma, .2CTA =
cta_group::2), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UTMAREDG (TMA load / store / reduce-add), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), plus MUFU.EX2 and the packed fp32 forms.  usage: python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "video_depth_normal_v2_b200", "libvdn_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
keys = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UTMALDG.MULTICAST", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "MUFU.EX2", "FFMA2", "FADD2", "FMNMX3", "USETMAXREG"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        per[cur]["_total"] += 1
        base = op.split(".")[0]
        if base in ("UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "FFMA2", "FADD2", "FMNMX3", "USETMAXREG"):
            per[cur][base] += 1
        if op.startswith("UTCHMMA") and ".2CTA" in op:
            per[cur]["UTCHMMA.2CTA"] += 1
        if op.startswith("UTMALDG") and "MULTICAST" in op:
            per[cur]["UTMALDG.MULTICAST"] += 1
        if op.startswith("MUFU.EX2"):
            per[cur]["MUFU.EX2"] += 1
tot = collections.Counter()
print(f"# {os.path.relpath(lib, ROOT)}: {len(per)} kernels (sm_100a SASS); columns: " + ", ".join(keys) + ", total instructions")
rows = []
for name, c in per.items():
    for k in keys:
        tot[k] += c[k]
    if any(c[k] for k in keys[:9]):
        d = demangle(name)
        d = re.sub(r"\((int|bool)\)", "", d)
        d = re.sub(r"\(.*", "", d).replace("void ", "").replace("vdn::", "")
        rows.append((d, c))
for d, c in sorted(rows):
    print(f"{d[:70]:70s} " + " ".join(f"{c[k]:5d}" for k in keys) + f" {c['_total']:7d}")
print("TOTAL".ljust(70) + " " + " ".join(f"{tot[k]:5d}" for k in keys))
ldd = subprocess.run(["ldd", lib], capture_output=True, text=True).stdout
print("# ldd:", ", ".join(sorted(l.split()[0] for l in ldd.splitlines() if "=>" in l or "ld-linux" in l)))
