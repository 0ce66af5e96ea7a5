"""The CPU restatement (oracle/) must reproduce the golden SAMs written by the unmodified reference binary
(tests/golden/make_golden.py) in all four modes: -map, -map -veryfast, -map2, -map2 -veryfast."""
import os

import pytest

from urmap_b200 import synth


def _cmp(golden, sam):
    c = synth.compare_sam(golden, sam)
    assert c["header_equal"]
    assert c["only_a"] == 0 and c["only_b"] == 0
    assert c["identical"] == c["total"], c["diffs"][:5]
    return c["total"]


@pytest.mark.parametrize("name,method", [("se.sam", 6), ("se_veryfast.sam", 7)])
def test_se_golden(oracle, golden_oix, golden_dir, name, method):
    b = oracle.ReadBatch.from_fastq(os.path.join(golden_dir, "se.fq"))
    res, runs, st = oracle.map_se(golden_oix, b, method=method, want_stats=True)
    sam = oracle.sam_header(golden_oix) + oracle.sam_se(golden_oix, b, res, runs)
    assert _cmp(os.path.join(golden_dir, name), sam) == 580
    assert st["tb_poison_reads"] == 0  # traceback never reads a cell the same DP call did not write


@pytest.mark.parametrize("name,pe_method", [("pe.sam", 4), ("pe_veryfast.sam", 5)])
def test_pe_golden(oracle, golden_oix, golden_dir, name, pe_method):
    b1 = oracle.ReadBatch.from_fastq(os.path.join(golden_dir, "pe_1.fq"))
    b2 = oracle.ReadBatch.from_fastq(os.path.join(golden_dir, "pe_2.fq"))
    r1, r2, runs, st = oracle.map_pe(golden_oix, b1, b2, pe_method=pe_method, want_stats=True)
    sam = oracle.sam_header(golden_oix) + oracle.sam_pe(golden_oix, b1, b2, r1, r2, runs)
    assert _cmp(os.path.join(golden_dir, name), sam) == 860
    assert st["tb_poison_reads"] == 0
    if pe_method == 4:
        assert st["scan_calls"] > 0  # the fixture exercises mate rescue (ScanPair)


def test_threads_do_not_change_results(oracle, golden_oix, golden_dir):
    b = oracle.ReadBatch.from_fastq(os.path.join(golden_dir, "se.fq"))
    r1, u1 = oracle.map_se(golden_oix, b, threads=1)
    r4, u4 = oracle.map_se(golden_oix, b, threads=4)
    assert (r1 == r4).all() and (u1 == u4).all()


def test_prime_table(oracle):
    # 5 083 340 B FASTA / 0.6 -> 8 856 593 ; human scale -> 5 392 814 809
    assert oracle.get_prime(int(5083340 / 0.6)) == 8856593
    assert oracle.get_prime(int(3151666838 / 0.6)) == 5392814809
    assert oracle.get_prime(100) == 101


def test_cigar_polish(oracle):
    # cigar.cpp:141-199: leading <=2M followed by a >4 gap folds into the next M; same at the tail
    assert oracle.path_to_cigar("M" + "D" * 6 + "M" * 100, 107) == "6I101M"
    assert oracle.path_to_cigar("M" * 100 + "I" * 6 + "MM", 102) == "102M6D"
    assert oracle.path_to_cigar("M" * 3 + "D" * 6 + "M" * 100, 109) == "3M6I100M"
    assert oracle.path_to_cigar("", 150) == "150M"
    assert oracle.path_to_cigar("MMMDMMM", 7) == "3M1I3M"


def test_viterbi_small(oracle):
    # viterbi.cpp:289-302 smoke strings; expected values are what the reference's recurrence gives
    sc, path = oracle.viterbi(b"GGGGATTAC", b"GGGGATTACA", left=False, right=True)
    assert sc == 9.0 and path == "MMMMMMMMMI"
    sc, path = oracle.viterbi(b"GGATTACA", b"GGGGATTACA", left=True, right=False)
    assert sc == 8.0 and path == "IIMMMMMMMM"
    sc, path = oracle.viterbi(b"", b"ACGT", left=False, right=False)
    assert path == "IIII"
