"""Pins the oracle restatement against the UNMODIFIED reference binary (oracle/_ref/urmap) on fresh seeded data.
Skipped where the binary is absent (it is built from /root/reference/src by oracle/Makefile when that exists)."""
import os

import pytest

from urmap_b200 import synth


@pytest.fixture(scope="module")
def ref_env(oracle, tmp_path_factory):
    if not os.path.exists(oracle.REF_BIN):
        pytest.skip("reference binary not built (no /root/reference here)")
    d = tmp_path_factory.mktemp("refenv")
    g = synth.make_genome(1_000_000, n_contigs=2, seed=31, repeat_frac=0.05, n_runs=[(0, 0.3, 500)], tandem=4)
    fa, ufi = str(d / "ref.fa"), str(d / "ref.ufi")
    g.write_fasta(fa)
    oracle.run_reference(["-make_ufi", fa, "-output", ufi])
    return d, g, ufi


@pytest.mark.parametrize("sub,indel,veryfast", [(0.01, 0.001, False), (0.05, 0.01, False), (0.02, 0.002, True)])
def test_se_matches_reference(oracle, ref_env, sub, indel, veryfast):
    d, g, ufi = ref_env
    reads, names = synth.sim_se(g, 4000, 150, sub, indel, seed=int(sub * 1000) + 5)
    fq, sam = str(d / "se.fq"), str(d / "se.sam")
    synth.write_fastq(fq, reads, names)
    oracle.run_reference(["-map", fq, "-ufi", ufi, "-samout", sam, "-threads", "1"] + (["-veryfast"] if veryfast else []))
    ix = oracle.Index(ufi)
    b = oracle.ReadBatch.from_fastq(fq)
    res, runs = oracle.map_se(ix, b, method=7 if veryfast else 6, threads=4)
    mine = oracle.sam_header(ix) + oracle.sam_se(ix, b, res, runs)
    c = synth.compare_sam(sam, mine)
    assert c["identical"] == c["total"] == 4000 and c["header_equal"], c["diffs"][:5]


@pytest.mark.parametrize("sub,indel,veryfast,rl", [(0.01, 0.001, False, 150), (0.05, 0.01, False, 150), (0.02, 0.002, True, 150),
                                                   (0.02, 0.002, False, 250)])
def test_pe_matches_reference(oracle, ref_env, sub, indel, veryfast, rl):
    d, g, ufi = ref_env
    r1, r2, names = synth.sim_pe(g, 3000, rl, sub, indel, seed=int(sub * 1000) + 7)
    f1, f2, sam = str(d / "p1.fq"), str(d / "p2.fq"), str(d / "pe.sam")
    synth.write_fastq(f1, r1, names, b"/1")
    synth.write_fastq(f2, r2, names, b"/2")
    oracle.run_reference(["-map2", f1, "-reverse", f2, "-ufi", ufi, "-samout", sam, "-threads", "1"]
                         + (["-veryfast"] if veryfast else []))
    ix = oracle.Index(ufi)
    b1, b2 = oracle.ReadBatch.from_fastq(f1), oracle.ReadBatch.from_fastq(f2)
    o1, o2, runs = oracle.map_pe(ix, b1, b2, pe_method=5 if veryfast else 4, threads=4)
    mine = oracle.sam_header(ix) + oracle.sam_pe(ix, b1, b2, o1, o2, runs)
    c = synth.compare_sam(sam, mine)
    assert c["identical"] == c["total"] == 6000 and c["header_equal"], c["diffs"][:5]


def test_maxix_above_32_matches_reference(oracle, tmp_path):
    """-maxix 100 (ufindexio.cpp:135-136): the reference indexes and walks lists of up to 100 positions; the restatement's
    rows follow (repeat-rich genome with tandem arrays, single-end and paired)."""
    if not os.path.exists(oracle.REF_BIN):
        pytest.skip("reference binary not built (no /root/reference here)")
    g = synth.make_genome(300_000, n_contigs=2, seed=78, repeat_frac=0.35, n_runs=[(0, 0.3, 300)], tandem=12)
    fa, ufi = str(tmp_path / "m.fa"), str(tmp_path / "m.ufi")
    g.write_fasta(fa)
    oracle.run_reference(["-make_ufi", fa, "-output", ufi, "-maxix", "100"])
    ix = oracle.Index(ufi)
    assert ix.max_ix == 100
    reads, names = synth.sim_se(g, 1500, 150, 0.02, 0.002, seed=5)
    fq, sam = str(tmp_path / "se.fq"), str(tmp_path / "se.sam")
    synth.write_fastq(fq, reads, names)
    oracle.run_reference(["-map", fq, "-ufi", ufi, "-samout", sam, "-threads", "1"])
    b = oracle.ReadBatch.from_fastq(fq)
    res, runs = oracle.map_se(ix, b, threads=4)
    c = synth.compare_sam(sam, oracle.sam_header(ix) + oracle.sam_se(ix, b, res, runs))
    assert c["identical"] == c["total"] == 1500 and c["header_equal"], c["diffs"][:5]
    r1, r2, names = synth.sim_pe(g, 1000, 150, 0.02, 0.002, seed=15)
    f1, f2, sam2 = str(tmp_path / "p1.fq"), str(tmp_path / "p2.fq"), str(tmp_path / "pe.sam")
    synth.write_fastq(f1, r1, names, b"/1")
    synth.write_fastq(f2, r2, names, b"/2")
    oracle.run_reference(["-map2", f1, "-reverse", f2, "-ufi", ufi, "-samout", sam2, "-threads", "1"])
    b1, b2 = oracle.ReadBatch.from_fastq(f1), oracle.ReadBatch.from_fastq(f2)
    o1, o2, runs = oracle.map_pe(ix, b1, b2, threads=4)
    c = synth.compare_sam(sam2, oracle.sam_header(ix) + oracle.sam_pe(ix, b1, b2, o1, o2, runs))
    assert c["identical"] == c["total"] == 2000 and c["header_equal"], c["diffs"][:5]
    ix.close()
