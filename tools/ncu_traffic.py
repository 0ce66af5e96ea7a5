"""DRAM traffic per read pair of a kernel from `ncu --set full` raw CSV dumps -> profiles/ncu_traffic.json (read by bench.py
for roofline.traffic).  Usage: python tools/ncu_traffic.py <pairs in the profiled launch> name=raw.csv [name=raw.csv ...]"""
import csv, json, os, sys
pairs = int(sys.argv[1])
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out_path = os.path.join(os.path.dirname(__file__), "..", "profiles", "ncu_traffic.json")
try:
    out = json.load(open(out_path))
except Exception:
    out = {}
for arg in sys.argv[2:]:
    name, path = arg.split("=", 1)
    rows = list(csv.reader(open(path)))
    h, u, v = rows[0], rows[1], rows[2]
    tot = 0.0
    for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = h.index(key)
        tot += float(v[i]) * UNIT[u[i]]
    t = float(v[h.index("gpu__time_duration.sum")])
    out[name] = {"dram_bytes_per_pair": tot / pairs, "profiled_pairs": pairs, "profiled_launch_ms": t if u[h.index("gpu__time_duration.sum")] == "ms" else None,
                 "source": os.path.basename(path)}
json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
