"""Aggregate an ncu source-page CSV dump (see tools/ncu_lines.py) by enclosing __device__/__global__ function of
urmap_b200/csrc/urmb_kernels.cu.  Usage: python tools/ncu_funcs.py dump.csv [source.cu]"""
import csv, re, sys, os
src = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "urmap_b200", "csrc", "urmb_kernels.cu")
lines = open(src).read().split("\n")
func_at = {}
cur = "?"
pat = re.compile(r"^(?:template.*>\s*)?(?:static\s+)?(?:__device__|__global__|__host__).*?\b([A-Za-z_0-9]+)\s*\(")
for i, l in enumerate(lines, 1):
    m = pat.match(l)
    if m and not l.rstrip().endswith(";"):
        cur = m.group(1)
    func_at[i] = cur
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if r and r[0] == "Line No")
i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
agg = {}
tot = tots = 0
for r in rows:
    if len(r) <= i_inst or not r[0].isdigit():
        continue
    try:
        inst, s = int(r[i_inst]), int(r[i_samp])
    except ValueError:
        continue
    f = func_at.get(int(r[0]), "?")
    a = agg.setdefault(f, [0, 0])
    a[0] += inst
    a[1] += s
    tot += inst
    tots += s
print(f"total warp instructions {tot}  stall samples {tots}")
for f, (inst, s) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{100 * inst / max(tot, 1):5.1f}% inst {100 * s / max(tots, 1):5.1f}% samples  {f}")
