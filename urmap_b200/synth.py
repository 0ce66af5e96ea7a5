"""Deterministic synthetic genomes and reads for the BASELINE.json configs.

Everything is seeded numpy, vectorised (no per-base Python loops) so 10^6..10^7 reads are
practical.  Used by tests/, bench.py and tools/; nothing here touches the GPU.

Read error model: per-base substitution to a *different* base with
probability ``sub``; per-base indel with probability ``indel`` (half deletions, half
insertions of a random base after the base); reads are truncated to exactly ``read_len``.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.full(256, ord("N"), dtype=np.uint8)
for _a, _b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
    _COMP[_a] = _b

HUMAN_MB = [248, 242, 198, 190, 182, 171, 159, 145, 138, 134, 135, 133, 114, 107, 102, 90, 83, 80,
            59, 64, 47, 51, 156, 57]


class Genome:
    """Concatenated ASCII genome + contig directory (no padding between contigs)."""

    def __init__(self, names, lens, asc):
        self.names = list(names)
        self.lens = np.asarray(lens, dtype=np.int64)
        self.offs = np.concatenate([[0], np.cumsum(self.lens)]).astype(np.int64)
        self.asc = asc  # uint8, len == sum(lens)

    def write_fasta(self, path, cols=60):
        with open(path, "wb") as f:
            for ci, nm in enumerate(self.names):
                f.write(b">" + nm.encode() + b"\n")
                seq = self.asc[self.offs[ci]:self.offs[ci + 1]]
                n = len(seq) // cols * cols
                if n:
                    a = np.empty((n // cols, cols + 1), dtype=np.uint8)
                    a[:, :cols] = seq[:n].reshape(-1, cols)
                    a[:, cols] = 10
                    a.tofile(f)
                if n < len(seq):
                    f.write(seq[n:].tobytes() + b"\n")


def make_genome(total_len, n_contigs=1, seed=12345, repeat_frac=0.0, n_runs=(), human_ratios=False,
                tandem=0, segdup=0, lower_frac=0.0):
    """iid ACGT background + optional injected repeat families / tandem repeats / seg-dups / N runs.

    repeat_frac : fraction of bases covered by copies of 300 bp-6 kb elements at 0-15 % divergence.
    tandem      : number of tandem-repeat arrays (period 2-60 bp x 5-200 copies).
    segdup      : number of segmental duplications (10-100 kb, 1-5 % divergence).
    n_runs      : iterable of (contig_index, start_fraction, length) runs of 'N'.
    lower_frac  : fraction of the genome written in lower case (soft-masked) in 1 kb blocks.
    """
    rng = np.random.default_rng(seed)
    G = int(total_len)
    if human_ratios:
        mb = HUMAN_MB[:n_contigs] if n_contigs <= 24 else HUMAN_MB
        tot = sum(mb)
        lens = [int(G * m / tot) for m in mb]
        lens[-1] += G - sum(lens)
        names = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"]
        names = names[:len(lens)]
    else:
        base = G // n_contigs
        lens = [base] * n_contigs
        lens[-1] += G - sum(lens)
        names = [f"ctg{i + 1}" for i in range(n_contigs)]
    g = rng.integers(0, 4, size=G, dtype=np.uint8)
    if repeat_frac > 0:
        target = int(repeat_frac * G)
        covered = 0
        while covered < target:
            L = int(rng.integers(300, 6001))
            K = int(rng.integers(20, 2001))
            K = max(2, min(K, (target - covered) // L + 1))
            elem = rng.integers(0, 4, size=L, dtype=np.uint8)
            pos = rng.integers(0, G - L, size=K)
            div = rng.uniform(0, 0.15, size=K)
            for p, d in zip(pos, div):
                c = elem.copy()
                m = rng.random(L) < d
                c[m] = (c[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) % 4
                g[p:p + L] = c
            covered += L * K
    for _ in range(int(tandem)):
        period = int(rng.integers(2, 61))
        copies = int(rng.integers(5, 201))
        unit = rng.integers(0, 4, size=period, dtype=np.uint8)
        arr = np.tile(unit, copies)
        m = rng.random(len(arr)) < 0.02
        arr[m] = (arr[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) % 4
        p = int(rng.integers(0, G - len(arr)))
        g[p:p + len(arr)] = arr
    for _ in range(int(segdup)):
        L = int(rng.integers(10_000, 100_001))
        L = min(L, G // 8)
        src = int(rng.integers(0, G - L))
        dst = int(rng.integers(0, G - L))
        c = g[src:src + L].copy()
        d = rng.uniform(0.01, 0.05)
        m = rng.random(L) < d
        c[m] = (c[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) % 4
        g[dst:dst + L] = c
    asc = _ACGT[g]
    del g
    offs = np.concatenate([[0], np.cumsum(lens)])
    for ci, frac, ln in n_runs:
        s = int(offs[ci] + lens[ci] * frac)
        asc[s:s + int(ln)] = ord("N")
    if lower_frac > 0:
        nblk = int(lower_frac * G / 1000)
        for p in rng.integers(0, max(1, G - 1000), size=nblk):
            asc[p:p + 1000] |= 0x20
    return Genome(names, lens, asc)


def _mutate(frag, sub, indel, read_len, rng):
    """frag: (n, L) uint8 ASCII (upper-case ACGT/N). Returns (n, read_len) uint8."""
    n, L = frag.shape
    x = rng.random((n, L), dtype=np.float32)
    is_sub = x < sub
    is_del = (x >= sub) & (x < sub + indel / 2)
    is_ins = (x >= sub + indel / 2) & (x < sub + indel)
    # substitution to a different base (only meaningful for ACGT; others left alone)
    code = np.full(256, 255, dtype=np.uint8)
    for i, c in enumerate(b"ACGT"):
        code[c] = i
    cf = code[frag]
    shift = rng.integers(1, 4, size=(n, L), dtype=np.uint8)
    subbed = _ACGT[(cf + shift) & 3]
    base = np.where(is_sub & (cf < 4), subbed, frag)
    emit = np.ones((n, L), dtype=np.int8)
    emit[is_del] = 0
    emit[is_ins] = 2
    tot = emit.sum(axis=1, dtype=np.int64)
    if tot.min() < read_len:
        raise ValueError("fragment padding too small for deletion rate")
    flat_src = np.repeat(np.arange(n * L, dtype=np.int64), emit.ravel())
    first = np.ones(len(flat_src), dtype=bool)
    first[1:] = flat_src[1:] != flat_src[:-1]
    row_start = np.concatenate([[0], np.cumsum(tot)[:-1]])
    idx = row_start[:, None] + np.arange(read_len)[None, :]
    src = flat_src[idx]
    out = base.ravel()[src]
    ins_mask = ~first[idx]
    ins_bases = _ACGT[rng.integers(0, 4, size=(n, read_len), dtype=np.uint8)]
    out = np.where(ins_mask, ins_bases, out)
    return np.ascontiguousarray(out)


def revcomp_rows(a):
    return np.ascontiguousarray(_COMP[a[:, ::-1]])


def _pick_positions(genome, n, span, rng):
    w = genome.lens.astype(np.float64)
    ci = rng.choice(len(w), size=n, p=w / w.sum())
    maxp = genome.lens[ci] - span
    if (maxp <= 0).any():
        raise ValueError("contig shorter than fragment span")
    p = (rng.random(n) * maxp).astype(np.int64)
    return ci, p


def sim_se(genome, n, read_len=150, sub=0.01, indel=0.001, seed=777, pad=40):
    """Single-end reads; strand 50/50. Returns (seqs (n,read_len) uint8, names list[bytes])."""
    rng = np.random.default_rng(seed)
    L = read_len + pad
    ci, p = _pick_positions(genome, n, L, rng)
    g0 = genome.offs[ci] + p
    frag = genome.asc[g0[:, None] + np.arange(L)[None, :]]
    minus = rng.random(n) < 0.5
    # minus-strand reads: take the rev-comp of the fragment first so that the read's first base
    # corresponds to the fragment's last base
    frag = np.where(minus[:, None], revcomp_rows(frag), frag)
    reads = _mutate(frag, sub, indel, read_len, rng)
    names = [b"r%d_%s_%d_%s" % (i, genome.names[c].encode(), pp + 1, b"-" if m else b"+")
             for i, (c, pp, m) in enumerate(zip(ci, p, minus))]
    return reads, names


def sim_pe(genome, n, read_len=150, sub=0.01, indel=0.001, seed=778, pad=40,
           ins_mean=400, ins_sd=50, ins_lo=200, ins_hi=800):
    """FR pairs; which mate is forward is randomised. Returns (r1, r2, names)."""
    rng = np.random.default_rng(seed)
    L = read_len + pad
    ins = np.clip(rng.normal(ins_mean, ins_sd, size=n), max(ins_lo, L), ins_hi).astype(np.int64)
    ci, p = _pick_positions(genome, n, ins_hi + pad, rng)
    g0 = genome.offs[ci] + p
    ar = np.arange(L)[None, :]
    left = genome.asc[g0[:, None] + ar]
    right = genome.asc[(g0 + ins - L)[:, None] + ar]
    a = _mutate(left, sub, indel, read_len, rng)
    b = _mutate(revcomp_rows(right), sub, indel, read_len, rng)
    flip = rng.random(n) < 0.5
    r1 = np.where(flip[:, None], b, a)
    r2 = np.where(flip[:, None], a, b)
    names = [b"p%d_%s_%d_%d_%s" % (i, genome.names[c].encode(), pp + 1, il, b"-" if fl else b"+")
             for i, (c, pp, il, fl) in enumerate(zip(ci, p, ins, flip))]
    return np.ascontiguousarray(r1), np.ascontiguousarray(r2), names


def write_fastq(path, reads, names, suffix=b"", qual=b"I"):
    n, L = reads.shape
    q = qual * L
    with open(path, "wb") as f:
        buf = []
        for i in range(n):
            buf.append(b"@" + names[i] + suffix + b"\n" + reads[i].tobytes() + b"\n+\n" + q + b"\n")
            if len(buf) >= 65536:
                f.write(b"".join(buf))
                buf = []
        f.write(b"".join(buf))


def parse_sam(path_or_bytes):
    """Returns (header_lines, {(qname, mate): fields_tuple}); mate = flag & 0xC0."""
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    hdr, recs = [], {}
    for line in data.split(b"\n"):
        if not line:
            continue
        if line.startswith(b"@"):
            hdr.append(line)
            continue
        f = line.split(b"\t")
        recs[(f[0], int(f[1]) & 0xC0)] = tuple(f)
    return hdr, recs


def compare_sam(a, b, ignore_pg=True):
    """Record-level comparison. Returns dict(total, identical, only_a, only_b, diffs[list of keys])."""
    ha, ra = parse_sam(a)
    hb, rb = parse_sam(b)
    if ignore_pg:
        ha = [h for h in ha if not h.startswith(b"@PG")]
        hb = [h for h in hb if not h.startswith(b"@PG")]
    keys = set(ra) | set(rb)
    diffs = [k for k in keys if ra.get(k) != rb.get(k)]
    return {
        "total": len(keys),
        "identical": len(keys) - len(diffs),
        "header_equal": ha == hb,
        "only_a": len(set(ra) - set(rb)),
        "only_b": len(set(rb) - set(ra)),
        "diffs": sorted(diffs)[:50],
    }
