"""The drop-in command line (urmap_b200/bin/urmap_b200): URMAP's option spellings, UFI files and SAM output."""
import gzip
import os
import shutil
import subprocess

import pytest

from conftest import ROOT, have_gpu
from urmap_b200 import synth

BIN = os.path.join(ROOT, "urmap_b200", "bin", "urmap_b200")


@pytest.fixture(scope="module")
def cli(built_lib):
    assert os.path.exists(BIN)
    return BIN


def run(args, **kw):
    return subprocess.run(args, capture_output=True, text=True, **kw)


def test_make_ufi_is_byte_identical(cli, golden_dir, tmp_path):
    """-make_ufi (CPU builder) reproduces the reference's UFI file byte for byte (ufindex.cpp:83-408)."""
    out = tmp_path / "ref.ufi"
    r = run([cli, "-make_ufi", os.path.join(golden_dir, "ref.fa"), "-output", str(out), "-quiet"])
    assert r.returncode == 0, r.stderr
    assert open(out, "rb").read() == open(os.path.join(golden_dir, "ref.ufi"), "rb").read()


def test_make_ufi_options_match_reference(cli, oracle, golden_dir, tmp_path):
    if not os.path.exists(oracle.REF_BIN):
        pytest.skip("reference binary not available")
    fa = os.path.join(golden_dir, "ref.fa")
    for extra in (["-veryfast"], ["-maxix", "5", "-load_factor", "0.7"], ["-slots", "300007"], ["-wordlength", "20"]):
        a, b = tmp_path / "a.ufi", tmp_path / "b.ufi"
        assert run([cli, "-make_ufi", fa, "-output", str(a), "-quiet"] + extra).returncode == 0
        oracle.run_reference(["-make_ufi", fa, "-output", str(b)] + extra)
        assert open(a, "rb").read() == open(b, "rb").read(), extra


def test_invalid_command_lines(cli):
    r = run([cli, "-bogus", "1"])
    assert r.returncode == 1 and "Invalid command line" in r.stderr
    r = run([cli, "-ufi", "x.ufi"])
    assert r.returncode == 1 and "Invalid command line" in r.stderr
    r = run([cli, "-map2", "a.fq", "-ufi", "x.ufi"])
    assert r.returncode == 1 and "-reverse required" in r.stderr and "---Fatal error---" in r.stderr


@pytest.mark.skipif(have_gpu(), reason="checks the no-device failure mode")
def test_map_without_gpu_fails_loudly(cli, golden_dir, tmp_path):
    r = run([cli, "-map", os.path.join(golden_dir, "se.fq"), "-ufi", os.path.join(golden_dir, "ref.ufi"), "-samout",
             str(tmp_path / "o.sam")])
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("name,args", [
    ("se.sam", ["-map", "se.fq"]),
    ("se_veryfast.sam", ["-map", "se.fq", "-veryfast"]),
    ("pe.sam", ["-map2", "pe_1.fq", "-reverse", "pe_2.fq"]),
    ("pe_veryfast.sam", ["-map2", "pe_1.fq", "-reverse", "pe_2.fq", "-veryfast"]),
])
def test_cli_sam_matches_golden(cli, golden_dir, tmp_path, name, args):
    if not have_gpu():
        pytest.skip("no CUDA device")
    args = [os.path.join(golden_dir, a) if a.endswith(".fq") else a for a in args]
    out = tmp_path / name
    r = run([cli] + args + ["-ufi", os.path.join(golden_dir, "ref.ufi"), "-samout", str(out), "-threads", "4", "-batch", "97"])
    assert r.returncode == 0, r.stderr
    assert "Reads/sec" in r.stderr and "Mapped Q>=10" in r.stderr
    a = open(os.path.join(golden_dir, name), "rb").read().split(b"\n")
    b = open(out, "rb").read().split(b"\n")
    strip = lambda ls: [l for l in ls if not l.startswith(b"@PG")]
    assert strip(a) == strip(b)  # same records, same (input) order, same @SQ header


@pytest.mark.gpu
def test_cli_gz_input_and_gpu_built_index(cli, golden_dir, tmp_path):
    if not have_gpu():
        pytest.skip("no CUDA device")
    gz = tmp_path / "se.fq.gz"
    with open(os.path.join(golden_dir, "se.fq"), "rb") as f, gzip.open(gz, "wb") as g:
        shutil.copyfileobj(f, g)
    ufi = tmp_path / "gpu.ufi"
    r = run([cli, "-make_ufi", os.path.join(golden_dir, "ref.fa"), "-output", str(ufi), "-gpu_build", "-quiet"])
    assert r.returncode == 0, r.stderr
    out = tmp_path / "o.sam"
    r = run([cli, "-map", str(gz), "-ufi", str(ufi), "-samout", str(out), "-quiet"])
    assert r.returncode == 0, r.stderr
    c = synth.compare_sam(os.path.join(golden_dir, "se.sam"), str(out))
    assert c["identical"] == c["total"] == 580 and c["header_equal"]
