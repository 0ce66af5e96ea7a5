"""Code bytes (SASS instructions x 16 B) and executed warp instructions per source function, from an ncu source-page CSV
dump made with --print-source cuda,sass.  Usage: python tools/ncu_codesize.py dump.csv [source.cu]"""
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
i0 = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
f = "?"
size, inst = {}, {}
for r in rows[i0 + 1:]:
    if r[0].isdigit():
        f = func_at.get(int(r[0]), "?")
    elif len(r) > 7 and r[2].startswith("0x"):
        size[f] = size.get(f, 0) + 16
        try:
            inst[f] = inst.get(f, 0) + int(r[7])
        except ValueError:
            pass
T, TI = sum(size.values()), sum(inst.values())
print(f"total code {T/1024:.1f} KB, executed {TI}")
for k, v in sorted(size.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{v/1024:7.1f} KB {100*inst.get(k,0)/max(TI,1):5.1f}% inst  {k}")
