"""Lane utilisation per source function: thread instructions / warp instructions (ncu source-page CSV dump).
Usage: python tools/ncu_lanes.py dump.csv [source.cu]"""
import csv, re, sys, os
src = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "urmap_b200", "csrc", "urmb_kernels.cu")
lines = open(src).read().split("\n")
func_at, cur = {}, "?"
pat = re.compile(r"^(?:template.*>\s*)?(?:static\s+)?(?:__device__|__global__|__host__).*?\b([A-Za-z_0-9]+)\s*\(")
for i, l in enumerate(lines, 1):
    m = pat.match(l)
    if m and not l.rstrip().endswith(";"):
        cur = m.group(1)
    func_at[i] = cur
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Line No")
i_inst, i_thr = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
agg = {}
for r in rows:
    if len(r) <= i_thr or not r[0].isdigit():
        continue
    try:
        a = agg.setdefault(func_at.get(int(r[0]), "?"), [0, 0])
        a[0] += int(r[i_inst]); a[1] += int(r[i_thr])
    except ValueError:
        pass
T = sum(a[0] for a in agg.values())
for f, (i, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{100*i/T:5.1f}% inst  {t/max(i,1):5.1f} lanes  {f}")
