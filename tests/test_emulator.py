"""Runs the CUDA kernel SOURCE (urmap_b200/csrc/urmb_kernels.cu) under the lock-step warp emulator of
tests/emu/ and compares it with the oracle.  This is a debugging / regression harness for machines without a
GPU: it also aborts on warp divergence at a collective and exposes intra-warp read/write races.  The real
parity gate is tests/test_gpu_parity.py (-m gpu)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

FIELDS = ("db_pos", "score", "best", "second", "mapq", "flags", "hit_count", "hsp_count")


def _paths(res, runs):
    return [tuple(runs[r["path_off"]:r["path_off"] + r["path_runs"]].tolist()) for r in res]


def _same(ro, uo, re_, ue):
    for f in FIELDS:
        assert (ro[f] == re_[f]).all(), (f, np.nonzero(ro[f] != re_[f])[0][:5])
    assert _paths(ro, uo) == _paths(re_, ue)


@pytest.mark.parametrize("method", [6, 7])
def test_emulated_se(oracle, golden_oix, golden_dir, method):
    import emu_py
    b = oracle.ReadBatch.from_fastq(os.path.join(golden_dir, "se.fq"))
    sel = np.r_[0:40, 300:340, 450:470, 510:530, 570:580]  # 150 bp 1 %, 5 %, 250 bp, 100 bp, random
    seqs = np.concatenate([b.seqs[b.offs[i]:b.offs[i + 1]] for i in sel])
    offs = np.concatenate([[0], np.cumsum([b.offs[i + 1] - b.offs[i] for i in sel])]).astype(np.uint32)
    ro, uo = oracle.map_se(golden_oix, oracle.ReadBatch(seqs, offs), method=method)
    re_, ue, cnt = emu_py.emu_map(golden_oix, oracle.RESULT_DTYPE, seqs, offs, len(sel), False, method=method)
    assert cnt[1] == 0
    _same(ro, uo, re_, ue)


@pytest.mark.parametrize("pe_method", [4, 5])
def test_emulated_pe(oracle, golden_oix, golden_dir, pe_method):
    import emu_py
    b1 = oracle.ReadBatch.from_fastq(os.path.join(golden_dir, "pe_1.fq"))
    b2 = oracle.ReadBatch.from_fastq(os.path.join(golden_dir, "pe_2.fq"))
    sel = np.r_[0:30, 250:280, 350:360, 390:430]  # 1 %, 5 %, 250 bp, damaged-mate (mate rescue) pairs

    def take(b):
        s = np.concatenate([b.seqs[b.offs[i]:b.offs[i + 1]] for i in sel])
        o = np.concatenate([[0], np.cumsum([b.offs[i + 1] - b.offs[i] for i in sel])]).astype(np.uint32)
        return s, o

    s1, o1 = take(b1)
    s2, o2 = take(b2)
    r1, r2, uo = oracle.map_pe(golden_oix, oracle.ReadBatch(s1, o1), oracle.ReadBatch(s2, o2), pe_method=pe_method)
    seqs = np.concatenate([s1, s2])
    offs = np.concatenate([o1, o2[1:] + o1[-1]]).astype(np.uint32)
    br = 4 if pe_method == 5 else -1
    re_, ue, cnt = emu_py.emu_map(golden_oix, oracle.RESULT_DTYPE, seqs, offs, len(sel), True, pe_method=pe_method,
                                  band_radius=br)
    assert cnt[1] == 0
    _same(np.concatenate([r1, r2]), uo, re_, ue)


@pytest.mark.parametrize("cap,rounds", [(0, 0), (1000, 0), (1000, 2), (1000, 6)])
def test_emulated_rescue_rounds(oracle, golden_oix, golden_dir, cap, rounds, monkeypatch):
    """Mate rescue continued from the saved states -- in place (the default), or in rounds of (window scans, batched
    full-window DPs) with a last round for the stragglers (URMB_RESCUE_ROUNDS) -- against the legacy kernel that searches the
    pair again from scratch (rescue pool capacity 0): all equal the oracle."""
    import emu_py
    monkeypatch.setenv("URMB_EMU_RESCUE_CAP", str(cap))
    monkeypatch.setenv("URMB_RESCUE_ROUNDS", str(rounds))
    b1 = oracle.ReadBatch.from_fastq(os.path.join(golden_dir, "pe_1.fq"))
    b2 = oracle.ReadBatch.from_fastq(os.path.join(golden_dir, "pe_2.fq"))
    sel = np.r_[255:270, 390:430]   # 5 % reads and damaged-mate pairs

    def take(b):
        s = np.concatenate([b.seqs[b.offs[i]:b.offs[i + 1]] for i in sel])
        o = np.concatenate([[0], np.cumsum([b.offs[i + 1] - b.offs[i] for i in sel])]).astype(np.uint32)
        return s, o

    s1, o1 = take(b1)
    s2, o2 = take(b2)
    r1, r2, uo, st = oracle.map_pe(golden_oix, oracle.ReadBatch(s1, o1), oracle.ReadBatch(s2, o2), want_stats=True)
    assert st["scan_calls"] > 20 and st["dp_cells_scan"] > 0
    seqs = np.concatenate([s1, s2])
    offs = np.concatenate([o1, o2[1:] + o1[-1]]).astype(np.uint32)
    re_, ue, cnt = emu_py.emu_map(golden_oix, oracle.RESULT_DTYPE, seqs, offs, len(sel), True)
    assert cnt[1] == 0 and cnt[2] > 20              # no overflow; pairs that needed mate rescue
    if cap:
        assert cnt[5] == 0                          # nothing left to the legacy kernel
        assert (cnt[6] > 10) == (rounds > 0)        # DPs run by the rounds' DP kernel (in place otherwise)
    else:
        assert cnt[5] == cnt[2] and cnt[6] == 0
    _same(np.concatenate([r1, r2]), uo, re_, ue)


def _tab_lines(contigs, labels, L1, L2, r1, r2, s1, s2):
    """State2::OutputTab2 (outputtab2.cpp:6-119) restated in Python over result records (test helper)."""
    def chrpos(pos):
        for lab, length, off in contigs:   # UFIndex::PosToCoord, ufindex.cpp:701-727
            if off <= pos < off + length:
                return lab, pos - off
        return "", 0xFFFFFFFF

    def pos1(h, fwd):
        lab, c = chrpos(h[1])
        return "%s:%u(%s)/%s" % (lab, (c + 1) & 0xFFFFFFFF, "+" if h[2] else "-", "1" if fwd else "2")

    def pairpos(h1, h2):
        if not h1[0] and not h2[0]:
            return "*"
        if h1[0] and not h2[0]:
            return pos1(h1, True)
        if h2[0] and not h1[0]:
            return pos1(h2, False)
        (l1, c1), (l2, c2) = chrpos(h1[1]), chrpos(h2[1])
        if l1 == l2 and h1[2] != h2[2]:
            return "%s:%u-%u" % (l1, (c1 + 1) & 0xFFFFFFFF, (c2 + 1) & 0xFFFFFFFF)
        return pos1(h1, True) + "," + pos1(h2, False)

    def tl(h1, h2, a, b):
        v = (h2[1] + b - h1[1]) if h1[1] <= h2[1] else (h1[1] + a - h2[1])
        return 0 if v < 0 or v > 1000 else v

    out = []
    for i, lab in enumerate(labels):
        t1 = (bool(r1[i]["flags"] & 2), int(r1[i]["db_pos"]), bool(r1[i]["flags"] & 1), int(r1[i]["score"]))
        t2 = (bool(r2[i]["flags"] & 2), int(r2[i]["db_pos"]), bool(r2[i]["flags"] & 1), int(r2[i]["score"]))
        u1 = (bool(s1[i]["flags"] & 2), int(s1[i]["db_pos"]), bool(s1[i]["flags"] & 1), int(s1[i]["score"]))
        u2 = (bool(s2[i]["flags"] & 2), int(s2[i]["db_pos"]), bool(s2[i]["flags"] & 1), int(s2[i]["score"]))
        f = [lab, pairpos(t1, t2), "%u,%u" % (r1[i]["mapq"], r2[i]["mapq"]), pairpos(u1, u2) if u1[0] else "*"]
        if t1[0] and t2[0] and u1[0] and u2[0]:
            a, b = tl(t1, t2, L1[i], L2[i]), tl(u1, u2, L1[i], L2[i])
            x, y = t1[3] + t2[3], u1[3] + u2[3]
            f.append(("TL=%u;" % a if a == b else "TL/%u,%u;" % (a, b)) + ("Score=%d;" % x if x == y else "Score/%d,%d;" % (x, y)))
        out.append("\t".join(f))
    return out


def test_emulated_second_pair_matches_reference_tabbedout(oracle, golden_oix, golden_dir, built_lib):
    """The second pair the kernels export (urmb_second, -tabbedout) against the reference's own tabbed output
    (tests/golden/pe.tab, written by the reference binary): all pairs with a second pair plus some without."""
    import emu_py
    from urmap_b200 import engine
    want = open(os.path.join(golden_dir, "pe.tab")).read().split("\n")[:-1]
    with_second = [i for i, l in enumerate(want) if l.split("\t")[3] != "*"]
    assert len(with_second) >= 9
    sel = sorted(set(with_second) | set(range(0, 12)) | set(range(392, 400)))
    b1 = oracle.ReadBatch.from_fastq(os.path.join(golden_dir, "pe_1.fq"))
    b2 = oracle.ReadBatch.from_fastq(os.path.join(golden_dir, "pe_2.fq"))

    def take(b):
        s = np.concatenate([b.seqs[b.offs[i]:b.offs[i + 1]] for i in sel])
        o = np.concatenate([[0], np.cumsum([b.offs[i + 1] - b.offs[i] for i in sel])]).astype(np.uint32)
        return s, o

    s1, o1 = take(b1)
    s2, o2 = take(b2)
    seqs = np.concatenate([s1, s2])
    offs = np.concatenate([o1, o2[1:] + o1[-1]]).astype(np.uint32)
    n = len(sel)
    second = np.zeros(2 * n, dtype=emu_py.SECOND_DTYPE)
    res, _, cnt = emu_py.emu_map(golden_oix, oracle.RESULT_DTYPE, seqs, offs, n, True, second=second)
    assert cnt[1] == 0
    hix = engine.HostIndex(os.path.join(golden_dir, "ref.ufi"))
    labels = [want[i].split("\t")[0] for i in sel]
    got = _tab_lines(hix.contigs, labels, np.diff(o1), np.diff(o2), res[:n], res[n:], second[:n], second[n:])
    assert got == [want[i] for i in sel]


def test_emulated_repeat_rich_rows(oracle, built_lib, tmp_path):
    """Repeat-rich genome (35 % repeats, tandem runs): most k-mers of a read own a non-BOTH1 slot, so the pending-row
    stages (plus and minus lists gathered together, rows <= 2 now / longer rows deferred, search1pepend.cpp:53-110 and
    search1m6.cpp:149-245) and the seed lists built from the probe rows carry the search.  Index from the drop-in's own
    host builder (byte-identical to the reference's, tests/test_cli.py)."""
    import subprocess
    import emu_py
    from urmap_b200 import synth
    g = synth.make_genome(300_000, n_contigs=2, seed=77, repeat_frac=0.35, n_runs=[(0, 0.3, 300)], tandem=6)
    fa, ufi = str(tmp_path / "rr.fa"), str(tmp_path / "rr.ufi")
    g.write_fasta(fa)
    cli = os.path.join(ROOT, "urmap_b200", "bin", "urmap_b200")
    r = subprocess.run([cli, "-make_ufi", fa, "-output", ufi, "-quiet"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ix = oracle.Index(ufi)
    try:
        reads, names = synth.sim_se(g, 120, 150, 0.03, 0.003, seed=3)
        fq = str(tmp_path / "se.fq")
        synth.write_fastq(fq, reads, names)
        b = oracle.ReadBatch.from_fastq(fq)
        for method in (6, 7):
            ro, uo = oracle.map_se(ix, b, method=method)
            re_, ue, cnt = emu_py.emu_map(ix, oracle.RESULT_DTYPE, b.seqs, b.offs, b.n, False, method=method)
            assert cnt[1] == 0
            _same(ro, uo, re_, ue)
        r1, r2, names = synth.sim_pe(g, 100, 150, 0.03, 0.003, seed=13)
        f1, f2 = str(tmp_path / "p1.fq"), str(tmp_path / "p2.fq")
        synth.write_fastq(f1, r1, names, b"/1")
        synth.write_fastq(f2, r2, names, b"/2")
        b1, b2 = oracle.ReadBatch.from_fastq(f1), oracle.ReadBatch.from_fastq(f2)
        o1, o2, uo, st = oracle.map_pe(ix, b1, b2, pe_method=4, want_stats=True)
        seqs = np.concatenate([b1.seqs, b2.seqs])
        offs = np.concatenate([b1.offs, b2.offs[1:] + b1.offs[-1]]).astype(np.uint32)
        re_, ue, cnt = emu_py.emu_map(ix, oracle.RESULT_DTYPE, seqs, offs, b1.n, True, pe_method=4)
        assert cnt[1] == 0
        assert st["row_calls"] > 20 * st["reads"] and st["row_hops"] > st["row_calls"]   # the case is what it claims to be
        _same(np.concatenate([o1, o2]), uo, re_, ue)
    finally:
        ix.close()


def test_emulated_maxix_above_32(oracle, built_lib, tmp_path):
    """An index built with -maxix 100 (ufindexio.cpp:135-136): rows of 33 .. 100 positions (tandem arrays) go through the
    deferred-row stage with a row stride of MaxIx instead of 32."""
    import subprocess
    import emu_py
    from urmap_b200 import synth
    g = synth.make_genome(300_000, n_contigs=2, seed=78, repeat_frac=0.35, n_runs=[(0, 0.3, 300)], tandem=12)
    fa, ufi, ufi32 = str(tmp_path / "m.fa"), str(tmp_path / "m.ufi"), str(tmp_path / "m32.ufi")
    g.write_fasta(fa)
    cli = os.path.join(ROOT, "urmap_b200", "bin", "urmap_b200")
    for out, extra in ((ufi, ["-maxix", "100"]), (ufi32, [])):
        r = subprocess.run([cli, "-make_ufi", fa, "-output", out, "-quiet"] + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    ix, ix32 = oracle.Index(ufi), oracle.Index(ufi32)
    try:
        assert ix.max_ix == 100
        reads, names = synth.sim_se(g, 150, 150, 0.02, 0.002, seed=5)
        fq = str(tmp_path / "se.fq")
        synth.write_fastq(fq, reads, names)
        b = oracle.ReadBatch.from_fastq(fq)
        ro, uo, st = oracle.map_se(ix, b, want_stats=True)
        _, _, st32 = oracle.map_se(ix32, b, want_stats=True)
        assert st["row_hops"] > st32["row_hops"]   # lists the default index leaves out are walked here
        re_, ue, cnt = emu_py.emu_map(ix, oracle.RESULT_DTYPE, b.seqs, b.offs, b.n, False)
        assert cnt[1] == 0
        _same(ro, uo, re_, ue)
        r1, r2, names = synth.sim_pe(g, 60, 150, 0.02, 0.002, seed=15)
        f1, f2 = str(tmp_path / "p1.fq"), str(tmp_path / "p2.fq")
        synth.write_fastq(f1, r1, names, b"/1")
        synth.write_fastq(f2, r2, names, b"/2")
        b1, b2 = oracle.ReadBatch.from_fastq(f1), oracle.ReadBatch.from_fastq(f2)
        o1, o2, uo = oracle.map_pe(ix, b1, b2)
        seqs = np.concatenate([b1.seqs, b2.seqs])
        offs = np.concatenate([b1.offs, b2.offs[1:] + b1.offs[-1]]).astype(np.uint32)
        re_, ue, cnt = emu_py.emu_map(ix, oracle.RESULT_DTYPE, seqs, offs, b1.n, True)
        assert cnt[1] == 0
        _same(np.concatenate([o1, o2]), uo, re_, ue)
    finally:
        ix.close()
        ix32.close()


@pytest.mark.parametrize("pe_method", [4, 5])
def test_emulated_first_look(oracle, built_lib, tmp_path, pe_method, monkeypatch):
    """The probe kernel's first look at paired input (the seed-loop exit of Search4/5, search2m4.cpp:79-142, replayed over the
    first eight iterator steps): a good part of clean pairs ends there with the oracle's records; switched off
    (URMB_FLAGS bit 9) nothing ends there and nothing changes."""
    import subprocess
    import emu_py
    from urmap_b200 import synth
    g = synth.make_genome(1_000_000, n_contigs=2, seed=12, repeat_frac=0.10)
    fa, ufi = str(tmp_path / "g.fa"), str(tmp_path / "g.ufi")
    g.write_fasta(fa)
    cli = os.path.join(ROOT, "urmap_b200", "bin", "urmap_b200")
    r = subprocess.run([cli, "-make_ufi", fa, "-output", ufi, "-quiet"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ix = oracle.Index(ufi)
    try:
        r1, r2, names = synth.sim_pe(g, 500, 150, 0.01, 0.001, seed=13)
        f1, f2 = str(tmp_path / "p1.fq"), str(tmp_path / "p2.fq")
        synth.write_fastq(f1, r1, names, b"/1")
        synth.write_fastq(f2, r2, names, b"/2")
        b1, b2 = oracle.ReadBatch.from_fastq(f1), oracle.ReadBatch.from_fastq(f2)
        br = 4 if pe_method == 5 else -1
        o1, o2, uo = oracle.map_pe(ix, b1, b2, pe_method=pe_method, band_radius=br)
        seqs = np.concatenate([b1.seqs, b2.seqs])
        offs = np.concatenate([b1.offs, b2.offs[1:] + b1.offs[-1]]).astype(np.uint32)
        looks = []
        for flags in ("0", "512"):
            monkeypatch.setenv("URMB_FLAGS", flags)
            re_, ue, cnt = emu_py.emu_map(ix, oracle.RESULT_DTYPE, seqs, offs, b1.n, True, pe_method=pe_method, band_radius=br)
            assert cnt[1] == 0
            _same(np.concatenate([o1, o2]), uo, re_, ue)
            looks.append(emu_py.emu_first_look())
        assert looks[1] == 0 and looks[0] > 0.2 * b1.n, looks
    finally:
        ix.close()
