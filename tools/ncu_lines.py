"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` dump by CUDA source line.
Usage: python tools/ncu_lines.py dump.csv [top_n] [samples]   (third argument: rank by stall samples instead of instructions)"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
hdr = next(r for r in rows if r and r[0] == "Line No")
i_inst, i_samp = hdr.index("Instructions Executed"), hdr.index("# Samples")
lines, tot, tots = [], 0, 0
for r in rows:
    if len(r) <= i_inst or not r[0].isdigit():
        continue
    try:
        inst, s = int(r[i_inst]), int(r[i_samp])
    except ValueError:
        continue
    lines.append((inst, s, int(r[0]), r[1][:100]))
    tot += inst
    tots += s
print(f"total warp instructions {tot}  stall samples {tots}")
by_samples = len(sys.argv) > 3 and sys.argv[3] == "samples"
for inst, s, ln, src in sorted(lines, key=lambda t: (t[1], t[0]) if by_samples else t, reverse=True)[:top]:
    print(f"{100 * inst / max(tot, 1):5.1f}% inst {100 * s / max(tots, 1):5.1f}% samples  L{ln}: {src}")
