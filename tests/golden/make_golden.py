"""Regenerates the golden fixtures in tests/golden/ with the UNMODIFIED reference binary
(oracle/_ref/urmap, built from /root/reference/src by oracle/Makefile).

    python tests/golden/make_golden.py

Inputs are small and deterministic (seeded numpy); they are committed next to the SAMs so the tests do not
depend on numpy's random stream staying stable.  The cases cover: 3 contigs (PADGAP / SetMappedPos), an N run,
soft-masked (lower-case) genome blocks, injected repeats and tandem repeats (rows > 1, MAPQ ties), reads with
N and lower-case bases, 100/150/250-bp reads, 1 % and 5 %/1 % error rates, SE and PE, default and -veryfast.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import numpy as np

from oracle import oracle_py as O
from urmap_b200 import synth


def main():
    g = synth.make_genome(150_000, n_contigs=3, seed=4242, repeat_frac=0.08, n_runs=[(1, 0.4, 300)], tandem=6,
                          lower_frac=0.05)
    # a 3 kb segmental duplication at 2 % divergence (multi-hit rows, MAPQ 0 ties)
    rng = np.random.default_rng(99)
    dup = g.asc[20_000:23_000].copy()
    m = rng.random(len(dup)) < 0.02
    dup[m] = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=int(m.sum()))]
    g.asc[120_000:123_000] = dup
    fa = os.path.join(HERE, "ref.fa")
    g.write_fasta(fa)
    gu = synth.Genome(g.names, g.lens, g.asc & 0xDF | (g.asc & 0x40))  # reads are simulated from upper case
    gu.asc = np.where(g.asc >= 97, g.asc - 32, g.asc).astype(np.uint8)

    def fq(path, parts, suffix=b""):
        with open(path, "wb") as f:
            k = 0
            for reads, names in parts:
                for i in range(len(names)):
                    s = reads[i].tobytes()
                    f.write(b"@" + names[i] + b".%d" % k + suffix + b"\n" + s + b"\n+\n" + b"I" * len(s) + b"\n")
                    k += 1

    a, an = synth.sim_se(gu, 300, 150, 0.01, 0.001, seed=1)
    b, bn = synth.sim_se(gu, 150, 150, 0.05, 0.01, seed=2)
    c, cn = synth.sim_se(gu, 60, 250, 0.01, 0.001, seed=3)
    d, dn = synth.sim_se(gu, 60, 100, 0.02, 0.002, seed=4)
    # sprinkle N and lower-case bases into some reads
    a = a.copy()
    a[5, 70] = ord("N")
    a[6, 10:14] = ord("N")
    a[7, :] = np.where(a[7] < 97, a[7] + 32, a[7])
    a[8, 100:120] += 32
    rnd = np.frombuffer(b"ACGT", dtype=np.uint8)[np.random.default_rng(5).integers(0, 4, size=(10, 150))]
    fq(os.path.join(HERE, "se.fq"), [(a, an), (b, bn), (c, cn), (d, dn), (rnd, [b"rnd%d" % i for i in range(10)])])
    p1, p2, pn = synth.sim_pe(gu, 250, 150, 0.01, 0.001, seed=6)
    q1, q2, qn = synth.sim_pe(gu, 100, 150, 0.05, 0.01, seed=7)
    s1, s2, sn = synth.sim_pe(gu, 40, 250, 0.02, 0.002, seed=8)
    # pairs whose second mate is heavily damaged in its first half (forces mate rescue / ScanPair)
    t1, t2, tn = synth.sim_pe(gu, 40, 150, 0.01, 0.001, seed=9)
    t2 = t2.copy()
    t2[:, :70] = rnd[:1, :70]
    fq(os.path.join(HERE, "pe_1.fq"), [(p1, pn), (q1, qn), (s1, sn), (t1, tn)], b"/1")
    fq(os.path.join(HERE, "pe_2.fq"), [(p2, pn), (q2, qn), (s2, sn), (t2, tn)], b"/2")

    ufi = os.path.join(HERE, "ref.ufi")
    O.run_reference(["-make_ufi", fa, "-output", ufi])
    ufi3 = "/tmp/golden_ref_maxix3.ufi"
    runs = [
        ("se.sam", ["-map", "se.fq"]),
        ("se_veryfast.sam", ["-map", "se.fq", "-veryfast"]),
        ("pe.sam", ["-map2", "pe_1.fq", "-reverse", "pe_2.fq"]),
        ("pe_veryfast.sam", ["-map2", "pe_1.fq", "-reverse", "pe_2.fq", "-veryfast"]),
    ]
    for out, args in runs:
        args = [os.path.join(HERE, x) if x.endswith(".fq") else x for x in args]
        O.run_reference(args + ["-ufi", ufi, "-samout", os.path.join(HERE, out), "-threads", "1"])
        n = sum(1 for l in open(os.path.join(HERE, out), "rb") if not l.startswith(b"@"))
        print(out, n, "records")
    print("ufi bytes", os.path.getsize(ufi))


if __name__ == "__main__":
    main()
