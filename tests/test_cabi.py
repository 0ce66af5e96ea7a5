"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/urmb.h declares.
No compute calls here (this runs without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, have_gpu


def _declared():
    txt = open(os.path.join(ROOT, "include", "urmb.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(urmb_[a-z_0-9]+)\s*\(", txt)))


def test_exports_match_header(built_lib):
    L = ctypes.CDLL(built_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for nm in names:
        assert hasattr(L, nm), f"liburmb.so does not export {nm}"
    assert sorted(built_lib.EXPORTS) == names


def test_result_layout(built_lib, oracle):
    assert built_lib.RESULT_DTYPE == oracle.RESULT_DTYPE
    assert built_lib.RESULT_DTYPE.itemsize == 20


def test_host_index_parse(built_lib, golden_dir, golden_oix):
    h = built_lib.HostIndex(os.path.join(golden_dir, "ref.ufi"))
    assert h.word_length == 24 and h.max_ix == 32
    assert h.slot_count == golden_oix.slot_count and h.seq_data_size == golden_oix.seq_size
    assert [c[0] for c in h.contigs] == ["ctg1", "ctg2", "ctg3"]
    assert h.contigs[1][2] == h.contigs[0][1] + 32  # PADGAP, ufindex.h:95
    assert (h.seq() == golden_oix.seq()).all()
    assert (h.blob()[:5000] == golden_oix.blob()[:5000]).all()
    h.close()


def test_bad_index_file(built_lib, tmp_path):
    p = tmp_path / "bad.ufi"
    p.write_bytes(b"not a ufi file at all, but long enough to pass the size check........")
    with pytest.raises(built_lib.UrmbError) as e:
        built_lib.HostIndex(str(p))
    assert e.value.code == -2
    with pytest.raises(built_lib.UrmbError):
        built_lib.HostIndex(str(tmp_path / "missing.ufi"))


@pytest.mark.skipif(have_gpu(), reason="checks the no-device failure mode")
def test_no_cpu_fallback(built_lib):
    with pytest.raises(built_lib.UrmbError) as e:
        built_lib.Context(0)
    assert e.value.code == -7 and "no CPU fallback" in str(e.value)
