"""BASELINE.json's target size through the drop-in itself: N pairs (default 10 M) of 2x150 bp reads against the 3.1 Gb
synthetic reference, FASTQ files -> `urmap_b200 -map2` -> SAM file, next to the unmodified reference binary on the same
files and host cores, and the two SAM files compared record by record (tools/samdiff.cpp).  Prints one JSON object.

    python tools/cli_scale.py --pairs 10000000 [--no-reference] [--gpus 1]

Runs on a GPU box only (the engine has no CPU path).  Everything lives in /dev/shm and is removed at the end."""
import argparse
import json
import os
import re
import shutil
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=10_000_000)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--gpus-list", default="", help="comma list of GPU counts: strong scaling of the same files through `-gpus G`")
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--tabbedout", action="store_true", help="also write and compare -tabbedout files")
    ap.add_argument("--write-threads", default="", help="comma list of URMB_WRITE_THREADS values to try (SAM file in /dev/shm)")
    ap.add_argument("--batches", default="", help="comma list of -batch values to try with the SAM text sent to /dev/null")
    ap.add_argument("--workdir", default="/dev/shm/urmb_cli_scale")
    a = ap.parse_args()
    # everything lives in RAM (tmpfs): 30 GB UFI + 0.63 GB FASTQ and 1.6 GB of SAM per million pairs, plus the readers'
    # own copies of the index (reference: 30 GB heap; drop-in: page cache mapping)
    need_gb = 30 + 30 + 2.3 * a.pairs / 1e6 + 16
    avail_gb = next(int(l.split()[1]) for l in open("/proc/meminfo") if l.startswith("MemAvailable")) / 1e6
    shm = shutil.disk_usage("/dev/shm").free / 1e9
    bench.log(f"MemAvailable {avail_gb:.0f} GB, /dev/shm free {shm:.0f} GB, need about {need_gb:.0f} GB")
    if avail_gb < need_gb or shm < need_gb - 30:
        print(json.dumps({"error": f"not enough memory for {a.pairs} pairs: {avail_gb:.0f} GB available, {shm:.0f} GB in /dev/shm"}))
        return
    import torch
    from oracle import oracle_py as O
    from urmap_b200 import build as BLD
    from urmap_b200 import gpu_synth
    BLD.build_engine()
    O.build(ref=True)
    device = torch.device("cuda", 0)
    torch.cuda.set_device(device)
    args = argparse.Namespace(genome_len=3_100_000_000, pairs_per_step=1_000_000, read_len=150, sub=0.01, indel=0.001,
                              single_end=False, segdup_frac=0.03, tandem_frac=0.01)
    meta, seq, blob = bench.build_workload(args, 0, 1, device)
    os.makedirs(a.workdir, exist_ok=True)
    ufi = os.path.join(a.workdir, "ref.ufi")
    t0 = time.time()
    bench.write_ufi_file(ufi, meta, seq, blob)
    bench.log(f"UFI file written in {time.time() - t0:.1f}s")
    del blob
    torch.cuda.empty_cache()
    # reads: chunks of 1 M pairs simulated on the GPU, appended to the two FASTQ files
    RL, CH = 150, 1_000_000
    prefix = os.path.join(a.workdir, "reads")
    t0 = time.time()
    for f in (prefix + "_1.fq", prefix + "_2.fq"):
        open(f, "wb").close()
    done = 0
    k = 0
    while done < a.pairs:
        n = min(CH, a.pairs - done)
        r1, r2 = gpu_synth.sim_pe(seq, meta["lens"], meta["offsets"], n, device, RL, 0.01, 0.001, seed=7000 + k)
        h1, h2 = r1.cpu().numpy().reshape(-1), r2.cpu().numpy().reshape(-1)
        tmp = os.path.join(a.workdir, "chunk")
        write_chunk(tmp, h1, h2, done, n, RL)
        for sfx in ("_1.fq", "_2.fq"):
            with open(prefix + sfx, "ab") as out, open(tmp + sfx, "rb") as src:
                shutil.copyfileobj(src, out, 1 << 26)
            os.unlink(tmp + sfx)
        done += n
        k += 1
    bench.log(f"{a.pairs} pairs written as FASTQ in {time.time() - t0:.1f}s")
    del seq
    torch.cuda.empty_cache()
    threads = os.cpu_count()
    exe = os.path.join(ROOT, "urmap_b200", "bin", "urmap_b200")
    samdiff = os.path.join(a.workdir, "samdiff")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", os.path.join(ROOT, "tools", "samdiff.cpp"), "-o", samdiff])
    out = {"pairs": a.pairs, "reads": 2 * a.pairs, "host_threads": threads, "gpus": a.gpus,
           "workload": "3.1 Gb synthetic reference (24 contigs, 10% repeats), 2x150 bp, 1% subs + 0.1% indels"}

    def cli(sam, batch=None, tab=None, gpus=None, **env):
        c = [exe, "-map2", prefix + "_1.fq", "-reverse", prefix + "_2.fq", "-ufi", ufi, "-samout", sam, "-threads",
             str(threads), "-gpus", str(gpus or a.gpus)] + (["-batch", str(batch)] if batch else []) + (["-tabbedout", tab] if tab else [])
        t0 = time.time()
        p = subprocess.run(c, capture_output=True, env=dict(os.environ, URMB_PROFILE="1", **env))
        wall = time.time() - t0
        err = p.stderr.decode(errors="replace")
        if p.returncode != 0:
            raise RuntimeError(err[-500:])
        prof = " ".join(ln for ln in err.splitlines() if ln.startswith("[urmb host]"))
        m = re.search(r"load ([0-9.]+)s.*mapper total ([0-9.]+)s", prof)
        return {"wall_seconds": wall, "seconds_to_load_index": float(m.group(1)), "seconds_in_mapper": float(m.group(2)),
                "reads_per_s_mapper": 2 * a.pairs / float(m.group(2)), "reads_per_s_wall": 2 * a.pairs / wall,
                "host_profile": prof, "summary": [ln.strip() for ln in err.splitlines() if "Mapped Q" in ln or "Unmapped" in ln]}

    out["urmap_b200"] = cli(os.path.join(a.workdir, "urmb.sam"))
    if a.tabbedout:
        out["urmap_b200_with_tabbedout"] = cli(os.path.join(a.workdir, "urmb.sam"), tab=os.path.join(a.workdir, "urmb.tab"))
    out["urmap_b200_to_dev_null"] = cli("/dev/null")
    # strong scaling through the drop-in: the same FASTQ files with -gpus G (index replicated over NVLink, batches dealt
    # round-robin); the stage that bounds each run is visible in its host profile (reader / gpu wait / formatter / writer)
    ngpu = torch.cuda.device_count()
    for G in [int(x) for x in a.gpus_list.split(",") if x]:
        if G > ngpu:
            continue
        out[f"gpus_{G}_to_dev_null"] = cli("/dev/null", gpus=G)
        out[f"gpus_{G}_to_file"] = cli(os.path.join(a.workdir, "urmb2.sam"), gpus=G)
        d = subprocess.run([samdiff, os.path.join(a.workdir, "urmb.sam"), os.path.join(a.workdir, "urmb2.sam")], capture_output=True, text=True)
        out[f"gpus_{G}_same_sam_as_1_gpu"] = json.loads(d.stdout).get("pct") if d.returncode == 0 else None
    for wt in [int(x) for x in a.write_threads.split(",") if x]:
        out[f"to_file_write_threads_{wt}"] = cli(os.path.join(a.workdir, "urmb2.sam"), URMB_WRITE_THREADS=str(wt))
    for bsz in [int(x) for x in a.batches.split(",") if x]:
        out[f"to_dev_null_batch_{bsz}"] = cli("/dev/null", batch=bsz)
        out[f"to_file_batch_{bsz}"] = cli(os.path.join(a.workdir, "urmb2.sam"), batch=bsz)
    if not a.no_reference and os.path.exists(O.REF_BIN):
        c = [O.REF_BIN, "-map2", prefix + "_1.fq", "-reverse", prefix + "_2.fq", "-ufi", ufi, "-samout",
             os.path.join(a.workdir, "ref.sam"), "-threads", str(threads)] + (
                 ["-tabbedout", os.path.join(a.workdir, "ref.tab")] if a.tabbedout else [])
        for attempt in range(3):
            t0 = time.time()
            p = subprocess.run(c, capture_output=True, env=dict(os.environ, OMP_STACKSIZE="64M", OMP_NUM_THREADS=str(threads)))
            wall = time.time() - t0
            if p.returncode == 0:
                break
            bench.log(f"reference exited {p.returncode} (attempt {attempt + 1})")
        err = p.stderr.decode(errors="replace")
        ld = re.search(r"(\d+)\s+Seconds to load index", err)
        mp = re.search(r"(\d+)\s+Seconds in mapper", err)
        out["reference"] = {"returncode": p.returncode, "wall_seconds": wall,
                            "seconds_to_load_index": int(ld.group(1)) if ld else None,
                            "seconds_in_mapper": int(mp.group(1)) if mp else None,
                            "reads_per_s_wall": 2 * a.pairs / wall,
                            "reads_per_s_mapper": (2 * a.pairs / int(mp.group(1))) if mp and int(mp.group(1)) else None}
        if p.returncode == 0:
            d = subprocess.run([samdiff, os.path.join(a.workdir, "ref.sam"), os.path.join(a.workdir, "urmb.sam")],
                               capture_output=True, text=True)
            out["sam_identity_vs_reference"] = json.loads(d.stdout) if d.returncode == 0 else {"error": d.stderr[-300:]}
            if a.tabbedout:
                um = os.path.join(a.workdir, "unmatched_tab.txt")
                d = subprocess.run([samdiff, os.path.join(a.workdir, "ref.tab"), os.path.join(a.workdir, "urmb.tab"), um],
                                   capture_output=True, text=True)
                out["tabbedout_identity_vs_reference"] = json.loads(d.stdout) if d.returncode == 0 else {"error": d.stderr[-300:]}
                if os.path.exists(um) and os.path.getsize(um):   # the lines only one of the two files has (and their SAM records)
                    lines = open(um).read().split("\n")[:40]
                    out["tabbedout_unmatched"] = lines
                    names = {l.split("\t")[1] if l[:2] in ("a\t", "b\t") else l.split("\t")[0] for l in lines if l}
                    recs = []
                    for fn in ("ref.sam", "urmb.sam"):
                        g = subprocess.run(["grep", "-m", "8", "-F", "-f", "/dev/stdin", os.path.join(a.workdir, fn)],
                                           input="\n".join(n for n in names if n), capture_output=True, text=True)
                        recs.append(g.stdout.split("\n")[:8])
                    out["tabbedout_unmatched_sam"] = recs
    print(json.dumps(out))
    shutil.rmtree(a.workdir, ignore_errors=True)


def write_chunk(prefix, r1, r2, base, n, RL):
    """bench.write_fastq_pair with read names continuing at `base`."""
    def one(path, arr, suffix):
        with open(path, "wb") as f:
            lo = base
            while lo < base + n:
                width = len(str(lo))
                hi = min(base + n, 10 ** width)
                m = hi - lo
                idx = np.arange(lo, hi, dtype=np.int64)
                rec = np.empty((m, 2 + width + len(suffix) + 1 + RL + 3 + RL + 1), np.uint8)
                rec[:, 0] = ord("@"); rec[:, 1] = ord("p"); c = 2
                for k in range(width):
                    rec[:, c + k] = (idx // 10 ** (width - 1 - k)) % 10 + ord("0")
                c += width
                rec[:, c:c + len(suffix)] = np.frombuffer(suffix, np.uint8); c += len(suffix)
                rec[:, c] = 10; c += 1
                rec[:, c:c + RL] = arr[(lo - base) * RL:(hi - base) * RL].reshape(m, RL); c += RL
                rec[:, c:c + 3] = np.frombuffer(b"\n+\n", np.uint8); c += 3
                rec[:, c:c + RL] = ord("I"); c += RL
                rec[:, c] = 10
                f.write(rec.tobytes())
                lo = hi
    one(prefix + "_1.fq", r1, b"/1")
    one(prefix + "_2.fq", r2, b"/2")


if __name__ == "__main__":
    main()
