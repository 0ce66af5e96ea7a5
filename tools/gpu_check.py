"""GPU parity + timing check against the oracle on config-1 style data (run under gpurun)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import oracle_py as O
from urmap_b200 import engine, synth

WORK = os.environ.get("URMB_WORK", "/tmp/urmb_work")
os.makedirs(WORK, exist_ok=True)


def paths(res, runs):
    return [tuple(runs[r["path_off"]:r["path_off"] + r["path_runs"]].tolist()) for r in res]


def compare(tag, ro, runs_o, rg, runs_g, limit=5):
    fields = ("db_pos", "score", "best", "second", "mapq", "flags", "hit_count", "hsp_count")
    bad = np.zeros(len(ro), dtype=bool)
    for f in fields:
        bad |= ro[f] != rg[f]
    po, pg = paths(ro, runs_o), paths(rg, runs_g)
    for i in range(len(ro)):
        if po[i] != pg[i]:
            bad[i] = True
    idx = np.nonzero(bad)[0]
    for i in idx[:limit]:
        print(f"  {tag} MISMATCH read {i}: oracle {ro[i]} {po[i]} | gpu {rg[i]} {pg[i]}")
    print(f"{tag}: n={len(ro)} mismatches={len(idx)}")
    return len(idx)


def main():
    n_se = int(os.environ.get("N_SE", 20000))
    t0 = time.time()
    g = synth.make_genome(5_000_000, n_contigs=3, seed=12345, n_runs=[(1, 0.5, 2000)])
    fa = os.path.join(WORK, "ref.fa")
    ufi = os.path.join(WORK, "ref.ufi")
    g.write_fasta(fa)
    O.build(ref=True)
    O.run_reference(["-make_ufi", fa, "-output", ufi])
    print("index built", time.time() - t0)
    oix = O.Index(ufi)
    hix = engine.HostIndex(ufi)
    ctx = engine.Context(0)
    ctx.set_index(hix)
    print("index uploaded", time.time() - t0)
    total_bad = 0
    for tag, sub, ind, n in (("se1", 0.01, 0.001, n_se), ("se5", 0.05, 0.01, n_se // 2)):
        reads, _ = synth.sim_se(g, n, 150, sub, ind, seed=777)
        b = O.ReadBatch.from_arrays(reads)
        ro, runs_o = O.map_se(oix, b, threads=os.cpu_count())
        t = time.time()
        rg, runs_g = ctx.map_se(b.seqs, b.offs)
        dt = time.time() - t
        tm = ctx.timing(0)
        print(f"{tag}: gpu e2e {dt*1e3:.1f} ms  {n/dt:.0f} reads/s  kernels {tm}")
        total_bad += compare(tag, ro, runs_o, rg, runs_g)
    for tag, sub, ind, n in (("pe1", 0.01, 0.001, n_se), ("pe5", 0.05, 0.01, n_se // 4)):
        r1, r2, _ = synth.sim_pe(g, n, 150, sub, ind, seed=778)
        b1, b2 = O.ReadBatch.from_arrays(r1), O.ReadBatch.from_arrays(r2)
        o1, o2, runs_o = O.map_pe(oix, b1, b2, threads=os.cpu_count())
        t = time.time()
        g1, g2, runs_g = ctx.map_pe(b1.seqs, b1.offs, b2.seqs, b2.offs)
        dt = time.time() - t
        tm = ctx.timing(0)
        print(f"{tag}: gpu e2e {dt*1e3:.1f} ms  {2*n/dt:.0f} reads/s  kernels {tm}")
        total_bad += compare(tag + "/1", o1, runs_o, g1, runs_g)
        total_bad += compare(tag + "/2", o2, runs_o, g2, runs_g)
    print("TOTAL MISMATCHES", total_bad)
    return 1 if total_bad else 0


if __name__ == "__main__":
    sys.exit(main())
