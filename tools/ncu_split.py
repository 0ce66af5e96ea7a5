"""Split a multi-kernel `ncu --page source --csv --print-source cuda,sass` dump (one section per kernel and source file)
into one CSV per kernel holding the urmb_kernels.cu section, ready for tools/ncu_stalls.py / ncu_lines.py / ncu_codesize.py.
Usage: python tools/ncu_split.py dump.csv[.gz] out_dir   -> out_dir/<kernel>_src.csv, prints the kernel names"""
import csv, gzip, io, os, re, sys
path, out = sys.argv[1], sys.argv[2]
os.makedirs(out, exist_ok=True)
raw = gzip.open(path, "rt").read() if path.endswith(".gz") else open(path).read()
rows = list(csv.reader(io.StringIO(raw)))
func, fpath, cur, seen = None, None, None, {}
def flush():
    global cur
    if cur and func and fpath and fpath.endswith("urmb_kernels.cu"):
        name = re.sub(r"^.*::", "", func.split("(")[0])
        k = seen.get(name, 0)
        seen[name] = k + 1
        fn = os.path.join(out, f"{name}{'' if k == 0 else '_' + str(k)}_src.csv")
        with open(fn, "w", newline="") as f:
            csv.writer(f).writerows(cur)
        print(name, len(cur) - 1, "lines ->", fn)
    cur = None
for r in rows:
    if r and r[0] == "File Path":
        flush()
        fpath = r[1] if len(r) > 1 else None
    elif r and r[0] == "Function Name":
        func = r[1] if len(r) > 1 else None
    elif r and r[0] == "Line No":
        cur = [r]
    elif cur is not None and r and (r[0].isdigit() or (len(r) > 2 and r[2].startswith("0x"))):
        cur.append(r)   # a source line, or one of the SASS instructions listed under it
flush()
