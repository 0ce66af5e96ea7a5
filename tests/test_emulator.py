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
