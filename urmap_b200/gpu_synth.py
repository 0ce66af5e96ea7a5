"""Synthetic genome / read generation on the GPU with torch (bench plumbing: the human-scale workload of
BASELINE.json configs[1..2] is far too big to synthesise on the host in a bench that must finish in minutes).
Same models as urmap_b200/synth.py."""
from __future__ import annotations

import numpy as np
import torch

from .synth import HUMAN_MB

PADGAP = 32  # ufindex.h:95


def contig_layout(total_len, n_contigs=24, human_ratios=True):
    G = int(total_len)
    if human_ratios:
        mb = HUMAN_MB[:n_contigs]
        tot = sum(mb)
        lens = [int(G * m / tot) for m in mb]
        lens[-1] += G - sum(lens)
        names = ([f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"])[:len(lens)]
    else:
        base = G // n_contigs
        lens = [base] * n_contigs
        lens[-1] += G - sum(lens)
        names = [f"ctg{i + 1}" for i in range(n_contigs)]
    offsets, o = [], 0
    for i, L in enumerate(lens):
        offsets.append(o)
        o += L + (PADGAP if i + 1 < len(lens) else 0)
    return names, lens, offsets, o  # o = SeqDataSize


def _g_to_seq(pos, lens, offsets):
    """Positions in the pad-free concatenation of the contigs -> sequence-data offsets (PADGAP between contigs)."""
    starts = np.concatenate([[0], np.cumsum(np.asarray(lens, dtype=np.int64))])[:-1]
    ci = np.searchsorted(starts, pos, side="right") - 1
    return np.asarray(offsets, dtype=np.int64)[ci] + (pos - starts[ci])


def _inject_segdups(g, G, frac, rng, gen, device):
    """Segmental duplications (BASELINE.json configs[4]): a 5-100 kb source segment copied to 1-5 other places, each copy
    with 0.5-5 % substitutions, half of the copies reverse-complemented.  Returns [(start, length)] of every copy and
    source in pad-free coordinates."""
    regions, covered, target = [], 0, int(frac * G)
    while covered < target:
        L = int(min(rng.integers(5_000, 100_001), max(1000, G // 64)))
        K = int(rng.integers(1, 6))
        src = int(rng.integers(0, G - L))
        seg = g[src:src + L].clone()
        regions.append((src, L))
        for _ in range(K):
            dst = int(rng.integers(0, G - L))
            c = seg
            if rng.random() < 0.5:
                c = (3 - seg).flip(0)
            m = torch.rand(L, device=device, generator=gen) < float(rng.uniform(0.005, 0.05))
            shift = torch.randint(1, 4, (L,), dtype=torch.uint8, device=device, generator=gen)
            g[dst:dst + L] = torch.where(m, (c + shift) & 3, c)
            regions.append((dst, L))
            covered += L
    return regions


def _inject_tandems(g, G, frac, rng, gen, device):
    """Tandem repeats (BASELINE.json configs[4]): arrays of a 2-60 bp unit, 200-5000 bp long, 0-10 % substitutions per
    array; all arrays are written by one scatter (non-overlapping by construction)."""
    target = int(frac * G)
    if target <= 0:
        return []
    n = max(1, target // 2600)
    T = rng.integers(200, 5001, size=n).astype(np.int64)
    U = rng.integers(2, 61, size=n).astype(np.int64)
    pos = np.sort(rng.integers(0, G - 5001, size=n)).astype(np.int64)
    keep = np.ones(n, dtype=bool)
    last_end = -1
    for i in range(n):
        if pos[i] < last_end:
            keep[i] = False
        else:
            last_end = pos[i] + T[i]
    T, U, pos = T[keep], U[keep], pos[keep]
    n = len(T)
    ustart = np.concatenate([[0], np.cumsum(U)])[:-1]
    units = torch.randint(0, 4, (int(U.sum()),), dtype=torch.uint8, device=device, generator=gen)
    Tt = torch.as_tensor(T, device=device)
    rid = torch.repeat_interleave(torch.arange(n, device=device), Tt)
    first = torch.as_tensor(np.concatenate([[0], np.cumsum(T)])[:-1], device=device)
    j = torch.arange(int(T.sum()), device=device) - first[rid]
    val = units[torch.as_tensor(ustart, device=device)[rid] + j % torch.as_tensor(U, device=device)[rid]]
    div = torch.as_tensor(rng.uniform(0.0, 0.10, size=n), device=device, dtype=torch.float32)[rid]
    m = torch.rand(val.shape, device=device, generator=gen) < div
    shift = torch.randint(1, 4, val.shape, dtype=torch.uint8, device=device, generator=gen)
    g[torch.as_tensor(pos, device=device)[rid] + j] = torch.where(m, (val + shift) & 3, val)
    return list(zip(pos.tolist(), T.tolist()))


def make_seqdata(total_len, device, seed=12345, n_contigs=24, human_ratios=True, repeat_frac=0.10, n_runs=3,
                 segdup_frac=0.0, tandem_frac=0.0, regions_out=None):
    """Returns (seqdata uint8 tensor [SeqDataSize + pad] upper-case ASCII with '-' pads, names, lens, offsets).
    segdup_frac / tandem_frac add segmental duplications and tandem repeats (the repeat-rich reference of
    BASELINE.json configs[4]); regions_out (a list) receives (sequence-data offset, length) of every such region so that
    reads can be drawn from them (sim_se / sim_pe `regions`).  With both at 0 the genome is the one earlier rounds used."""
    names, lens, offsets, sds = contig_layout(total_len, n_contigs, human_ratios)
    G = int(total_len)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    g = torch.randint(0, 4, (G,), dtype=torch.uint8, device=device, generator=gen)
    rng = np.random.default_rng(seed)
    covered, target = 0, int(repeat_frac * G)
    while covered < target:
        L = int(rng.integers(300, 6001))
        K = int(rng.integers(20, 2001))
        K = max(2, min(K, (target - covered) // L + 1))
        pos = np.sort(rng.integers(0, G - L, size=K))
        keep = np.ones(K, dtype=bool)
        last = -10**18
        for i, p in enumerate(pos):  # drop overlapping copies so that the scatter below is deterministic
            if p - last < L:
                keep[i] = False
            else:
                last = p
        pos = pos[keep]
        K = len(pos)
        elem = torch.randint(0, 4, (L,), dtype=torch.uint8, device=device, generator=gen)
        div = torch.rand((K, 1), device=device, generator=gen) * 0.15
        copies = elem[None, :].repeat(K, 1)
        m = torch.rand((K, L), device=device, generator=gen) < div
        shift = torch.randint(1, 4, (K, L), dtype=torch.uint8, device=device, generator=gen)
        copies = torch.where(m, (copies + shift) & 3, copies)
        idx = torch.as_tensor(pos, device=device)[:, None] + torch.arange(L, device=device)[None, :]
        g[idx.reshape(-1)] = copies.reshape(-1)
        covered += L * K
        del idx, copies, m, shift
    regions = []
    if segdup_frac > 0:
        regions += _inject_segdups(g, G, segdup_frac, rng, gen, device)
    if tandem_frac > 0:
        regions += _inject_tandems(g, G, tandem_frac, rng, gen, device)
    if regions_out is not None and regions:
        r = np.asarray(regions, dtype=np.int64)
        regions_out.extend(zip(_g_to_seq(r[:, 0], lens, offsets).tolist(), r[:, 1].tolist()))
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    seq = torch.empty(sds + 4096, dtype=torch.uint8, device=device)
    seq[sds:] = 0
    src = 0
    CH = 1 << 28
    for i, L in enumerate(lens):
        o = offsets[i]
        for c0 in range(0, L, CH):
            c1 = min(L, c0 + CH)
            seq[o + c0:o + c1] = lut[g[src + c0:src + c1].long()]
        if i + 1 < len(lens):
            seq[o + L:o + L + PADGAP] = ord("-")
        src += L
    del g
    for k in range(n_runs):
        ci = (0, 4, 9)[k % 3] % len(lens)
        s = offsets[ci] + lens[ci] // 3
        seq[s:s + min(50_000, lens[ci] // 10)] = ord("N")
    return seq, names, lens, offsets, sds


_COMP = None


def _comp_lut(device):
    global _COMP
    if _COMP is None or _COMP.device != torch.device(device):
        t = torch.full((256,), ord("N"), dtype=torch.uint8)
        for a, b in zip(b"ACGTN", b"TGCAN"):
            t[a] = b
        _COMP = t.to(device)
    return _COMP


def _mutate(frag, sub, indel, read_len, gen):
    n, L = frag.shape
    dev = frag.device
    x = torch.rand((n, L), device=dev, generator=gen)
    is_sub = x < sub
    is_del = (x >= sub) & (x < sub + indel / 2)
    is_ins = (x >= sub + indel / 2) & (x < sub + indel)
    code = torch.full((256,), 255, dtype=torch.uint8, device=dev)
    for i, c in enumerate(b"ACGT"):
        code[c] = i
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    cf = code[frag.long()]
    shift = torch.randint(1, 4, (n, L), dtype=torch.uint8, device=dev, generator=gen)
    subbed = lut[((cf + shift) & 3).long()]
    base = torch.where(is_sub & (cf < 4), subbed, frag)
    emit = torch.ones((n, L), dtype=torch.int64, device=dev)
    emit[is_del] = 0
    emit[is_ins] = 2
    tot = emit.sum(dim=1)
    assert int(tot.min()) >= read_len
    flat_src = torch.repeat_interleave(torch.arange(n * L, device=dev), emit.reshape(-1))
    first = torch.ones(flat_src.numel(), dtype=torch.bool, device=dev)
    first[1:] = flat_src[1:] != flat_src[:-1]
    row_start = torch.cumsum(tot, 0) - tot
    idx = row_start[:, None] + torch.arange(read_len, device=dev)[None, :]
    src = flat_src[idx]
    out = base.reshape(-1)[src]
    ins_mask = ~first[idx]
    ins_bases = lut[torch.randint(0, 4, (n, read_len), device=dev, generator=gen)]
    return torch.where(ins_mask, ins_bases, out).contiguous()


def _pick(m, span, lens_t, offs_t, w, gen, device, regions, enrich):
    """Fragment starts (sequence-data offsets) of m fragments of `span` bases: uniform over the genome, except that a
    fraction `enrich` starts inside (or up to span/2 before) one of `regions` [(offset, length)] chosen by length."""
    ci = torch.multinomial(w / w.sum(), m, replacement=True, generator=gen)
    p = (torch.rand(m, device=device, generator=gen, dtype=torch.float64) * (lens_t[ci] - span)).long()
    g0 = offs_t[ci] + p
    if regions is not None and enrich > 0:
        rs, rl = regions
        ri = torch.multinomial(rl.double() / rl.double().sum(), m, replacement=True, generator=gen)
        q = rs[ri] + (torch.rand(m, device=device, generator=gen, dtype=torch.float64) * rl[ri].double()).long() - span // 2
        cq = torch.clamp(torch.searchsorted(offs_t, q, right=True) - 1, 0, len(lens_t) - 1)
        q = torch.minimum(torch.maximum(q, offs_t[cq]), offs_t[cq] + lens_t[cq] - span)
        use = torch.rand(m, device=device, generator=gen) < enrich
        g0 = torch.where(use, q, g0)
    return g0


def regions_tensor(regions, device):
    r = torch.as_tensor(np.asarray(regions, dtype=np.int64), device=device)
    return r[:, 0].contiguous(), r[:, 1].contiguous()


def sim_pe(seq, lens, offsets, n, device, read_len=150, sub=0.01, indel=0.001, seed=778, pad=40, ins_mean=400,
           ins_sd=50, ins_lo=200, ins_hi=800, chunk=1 << 18, return_truth=False, regions=None, enrich=0.0):
    """FR pairs (which mate is forward is randomised). Returns (r1, r2) uint8 tensors [n, read_len] on device; with
    return_truth also (pos1, pos2, plus1): the sequence-data offset each mate was drawn from (start of its alignment
    on the plus strand, up to the read's own indels) and whether mate 1 is the forward one."""
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    L = read_len + pad
    w = torch.tensor(lens, dtype=torch.float64, device=device)
    offs_t = torch.tensor(offsets, dtype=torch.int64, device=device)
    lens_t = torch.tensor(lens, dtype=torch.int64, device=device)
    comp = _comp_lut(device)
    ar = torch.arange(L, device=device)[None, :]
    o1, o2 = [], []
    t1, t2, tf = [], [], []
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        ins = torch.clamp(torch.normal(float(ins_mean), float(ins_sd), (m,), device=device, generator=gen),
                          max(ins_lo, L), ins_hi).long()
        if regions is None:   # the draw order of earlier rounds (same reads for the same seed)
            ci = torch.multinomial(w / w.sum(), m, replacement=True, generator=gen)
            maxp = lens_t[ci] - (ins_hi + pad)
            p = (torch.rand(m, device=device, generator=gen, dtype=torch.float64) * maxp).long()
            g0 = offs_t[ci] + p
        else:
            g0 = _pick(m, ins_hi + pad, lens_t, offs_t, w, gen, device, regions, enrich)
        left = seq[g0[:, None] + ar]
        right = seq[(g0 + ins - L)[:, None] + ar]
        a = _mutate(left, sub, indel, read_len, gen)
        b = _mutate(comp[right.flip(1).long()], sub, indel, read_len, gen)
        flip = torch.rand(m, device=device, generator=gen) < 0.5
        o1.append(torch.where(flip[:, None], b, a))
        o2.append(torch.where(flip[:, None], a, b))
        if return_truth:
            pa, pb = g0, g0 + ins - read_len
            t1.append(torch.where(flip, pb, pa))
            t2.append(torch.where(flip, pa, pb))
            tf.append(~flip)
    if return_truth:
        return (torch.cat(o1).contiguous(), torch.cat(o2).contiguous(), torch.cat(t1), torch.cat(t2), torch.cat(tf))
    return torch.cat(o1).contiguous(), torch.cat(o2).contiguous()


def sim_se(seq, lens, offsets, n, device, read_len=150, sub=0.01, indel=0.001, seed=777, pad=40, chunk=1 << 18,
           regions=None, enrich=0.0):
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    L = read_len + pad
    w = torch.tensor(lens, dtype=torch.float64, device=device)
    offs_t = torch.tensor(offsets, dtype=torch.int64, device=device)
    lens_t = torch.tensor(lens, dtype=torch.int64, device=device)
    comp = _comp_lut(device)
    ar = torch.arange(L, device=device)[None, :]
    out = []
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        g0 = _pick(m, L, lens_t, offs_t, w, gen, device, regions, enrich)
        frag = seq[g0[:, None] + ar]
        minus = torch.rand(m, device=device, generator=gen) < 0.5
        frag = torch.where(minus[:, None], comp[frag.flip(1).long()], frag)
        out.append(_mutate(frag, sub, indel, read_len, gen))
    return torch.cat(out).contiguous()
