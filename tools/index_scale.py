"""Byte identity of the GPU index builder at BASELINE.json's full size: the 3.1 Gb synthetic reference of bench.py is written
as FASTA, the UNMODIFIED reference binary builds its UFI from it (single-threaded: about 12 minutes and 41 GB), the
GPU builder builds the table from the same sequence data, and the two are compared byte for byte (header fields, the
5 * SlotCount-byte blob, the sequence data).  Prints one JSON object.

    python tools/index_scale.py [--genome-len 3100000000]

Runs on a GPU box only; everything lives in /dev/shm and is removed at the end."""
import argparse
import json
import os
import shutil
import struct
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome-len", type=int, default=3_100_000_000)
    ap.add_argument("--workdir", default="/dev/shm/urmb_index_scale")
    ap.add_argument("--ref-timeout", type=int, default=1500)
    a = ap.parse_args()
    import torch
    from oracle import oracle_py as O
    from urmap_b200 import build as BLD
    BLD.build_engine()
    O.build(ref=True)
    device = torch.device("cuda", 0)
    torch.cuda.set_device(device)
    args = argparse.Namespace(genome_len=a.genome_len, pairs_per_step=1, read_len=150, sub=0.01, indel=0.001, single_end=False)
    meta, seq, blob = bench.build_workload(args, 0, 1, device)   # the GPU build (timed inside: meta["build_seconds"])
    os.makedirs(a.workdir, exist_ok=True)
    fa = os.path.join(a.workdir, "ref.fa")
    t0 = time.time()
    with open(fa, "wb") as f:   # 60 columns, the layout index_build.fasta_bytes assumes
        for name, L, off in zip(meta["names"], meta["lens"], meta["offsets"]):
            f.write(b">" + name.encode() + b"\n")
            CH = 60 * (1 << 20)
            for o in range(0, L, CH):
                n = min(CH, L - o)
                s = seq[off + o:off + o + n].cpu().numpy()
                full = n // 60 * 60
                rows = np.empty((full // 60, 61), np.uint8)
                rows[:, :60] = s[:full].reshape(-1, 60)
                rows[:, 60] = 10
                f.write(rows.tobytes())
                if n > full:
                    f.write(s[full:].tobytes() + b"\n")
    bench.log(f"FASTA written in {time.time() - t0:.1f}s ({os.path.getsize(fa)} bytes)")
    out = {"genome_len": a.genome_len, "slot_count": meta["slot_count"], "seq_data_size": meta["seq_data_size"],
           "gpu_build_seconds": meta["build_seconds"], "indexed_positions": meta["indexed"]}
    ufi = os.path.join(a.workdir, "ref.ufi")
    t0 = time.time()
    try:
        p = subprocess.run([O.REF_BIN, "-make_ufi", fa, "-output", ufi], capture_output=True, timeout=a.ref_timeout)
        out["reference_make_ufi_seconds"] = time.time() - t0
        out["reference_returncode"] = p.returncode
    except subprocess.TimeoutExpired:
        out["reference_make_ufi_seconds"] = None
        out["error"] = f"reference -make_ufi did not finish in {a.ref_timeout}s"
        print(json.dumps(out))
        shutil.rmtree(a.workdir, ignore_errors=True)
        return
    # compare: header, blob, sequence data
    raw = np.memmap(ufi, dtype=np.uint8, mode="r")
    hdr = bytes(raw[:24])
    magic, W, maxix, sds = struct.unpack("<IIII", hdr[:16])
    slots = struct.unpack("<Q", hdr[16:24])[0]
    out["reference_header"] = {"word_length": W, "max_ix": maxix, "seq_data_size": sds, "slot_count": slots}
    out["header_equal"] = (W == meta["word_length"] and maxix == meta["max_ix"] and sds == meta["seq_data_size"]
                           and slots == meta["slot_count"])
    m2 = struct.pack("<I", 0x55464932)
    pos = bytes(raw[:1 << 16]).index(m2) + 4
    nblob = 5 * meta["slot_count"]
    CH = 1 << 28
    diff = 0
    long_links = 0
    for o in range(0, nblob, CH):
        n = min(CH, nblob - o)
        mine = blob[o:o + n].cpu().numpy()
        ref = np.asarray(raw[pos + o:pos + o + n])
        diff += int(np.count_nonzero(mine != ref))
    tl = np.asarray(raw[pos:pos + nblob:5])
    long_links = int(np.count_nonzero((tl == 125) | (tl == 253)))
    out["blob_bytes"] = nblob
    out["blob_diff_bytes"] = diff
    out["long_link_records_in_reference"] = long_links
    spos = pos + nblob + 4
    sdiff = 0
    for o in range(0, meta["seq_data_size"], CH):
        n = min(CH, meta["seq_data_size"] - o)
        sdiff += int(np.count_nonzero(seq[o:o + n].cpu().numpy() != np.asarray(raw[spos + o:spos + o + n])))
    out["seq_diff_bytes"] = sdiff
    out["byte_identical"] = bool(out["header_equal"] and diff == 0 and sdiff == 0)
    print(json.dumps(out))
    del raw
    shutil.rmtree(a.workdir, ignore_errors=True)


if __name__ == "__main__":
    main()
