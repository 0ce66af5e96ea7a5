"""Per-function stall-reason breakdown from an ncu source-page CSV dump.
Usage: python tools/ncu_stalls.py dump.csv [source.cu]"""
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
cols = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
i_inst = hdr.index("Instructions Executed")
agg, tot = {}, {}
for r in rows:
    if len(r) < len(hdr) or not r[0].isdigit():
        continue
    f = func_at.get(int(r[0]), "?")
    a = agg.setdefault(f, {})
    for h, i in cols.items():
        try:
            v = int(r[i])
        except ValueError:
            continue
        a[h] = a.get(h, 0) + v
        tot[h] = tot.get(h, 0) + v
    try:
        a["inst"] = a.get("inst", 0) + int(r[i_inst])
    except ValueError:
        pass
T = sum(tot.values())
print("total samples", T, {k: f"{100*v/T:.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v > 0.01 * T})
top = sorted(tot, key=lambda k: -tot[k])[:6]
print("func".ljust(22), "samples%", " ".join(t[6:].rjust(10) for t in top))
for f, a in sorted(agg.items(), key=lambda kv: -sum(v for k, v in kv[1].items() if k != "inst"))[:30]:
    s = sum(v for k, v in a.items() if k != "inst")
    print(f.ljust(22), f"{100*s/T:7.1f}%", " ".join(f"{100*a.get(t,0)/T:9.1f}%" for t in top))
