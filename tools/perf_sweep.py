"""Kernel-time sweep over URMB_FLAGS variants on the human-scale workload (run under gpurun)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from urmap_b200 import engine, gpu_synth, index_build

ap = argparse.ArgumentParser()
ap.add_argument("--genome-len", type=int, default=3_100_000_000)
ap.add_argument("--pairs", type=int, default=500_000)
ap.add_argument("--flags", default="3,0,1,2")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--se", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda", 0)
human = args.genome_len >= 100_000_000
seq, names, lens, offsets, sds = gpu_synth.make_seqdata(args.genome_len, dev, n_contigs=24 if human else 3, human_ratios=human)
slots = index_build.slot_count_for(names, lens)
blob = torch.empty(5 * slots + 16, dtype=torch.uint8, device=dev)
torch.cuda.empty_cache()
print("build", index_build.build_index_device(seq.data_ptr(), sds, slots, blob.data_ptr()), flush=True)
B = args.pairs
offs = np.arange(B + 1, dtype=np.uint32) * 150
if args.se:
    r1 = gpu_synth.sim_se(seq, lens, offsets, B, dev).cpu().numpy().reshape(-1)
    r2 = None
else:
    a, b = gpu_synth.sim_pe(seq, lens, offsets, B, dev)
    r1, r2 = a.cpu().numpy().reshape(-1), b.cpu().numpy().reshape(-1)
ref = None
for fl in args.flags.split(","):
    os.environ["URMB_FLAGS"] = fl
    ctx = engine.Context(0)
    ctx.attach_index(24, 32, sds, slots, blob.data_ptr(), seq.data_ptr())
    ctx.upload(0, r1, offs, r2, offs if r2 is not None else None)
    best = None
    for _ in range(args.reps):
        ctx.launch(0)
        tm = ctx.timing(0)
        if best is None or tm["search_ms"] < best["search_ms"]:
            best = tm
    ctx.download(0)
    x1, x2, runs = ctx.wait(0, B, r2 is not None)
    sig = (int(x1["db_pos"].astype(np.uint64).sum()), int(x1["mapq"].sum()), int(x1["score"].sum()))
    if ref is None:
        ref = sig
    n = B * (1 if r2 is None else 2)
    print(f"flags={fl}: probe {best['probe_ms']:.1f} ms search {best['search_ms']:.1f} ms -> {n / (best['probe_ms'] + best['search_ms']) / 1e3:.2f} M reads/s  same_results={sig == ref}", flush=True)
    ctx.close()
