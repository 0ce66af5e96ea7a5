"""Back-to-back step time (context-wide CUDA-event marks) on the human-scale workload for a sweep over one environment
variable read at context creation (run under gpurun).  Usage: python tools/step_sweep.py --var URMB_RESCUE_WARPS --values 148,296,592"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from urmap_b200 import engine, gpu_synth, index_build

ap = argparse.ArgumentParser()
ap.add_argument("--genome-len", type=int, default=3_100_000_000)
ap.add_argument("--pairs", type=int, default=1_000_000)
ap.add_argument("--var", default="URMB_RESCUE_WARPS")
ap.add_argument("--values", default="")
ap.add_argument("--steps", type=int, default=6)
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
batches = []
for k in range(3):
    if args.se:
        batches.append((gpu_synth.sim_se(seq, lens, offsets, B, dev, seed=1000 + k).cpu().numpy().reshape(-1), None))
    else:
        a, b = gpu_synth.sim_pe(seq, lens, offsets, B, dev, seed=1000 + k)
        batches.append((a.cpu().numpy().reshape(-1), b.cpu().numpy().reshape(-1)))
ref = None
for val in (args.values.split(",") if args.values else [""]):
    if val:
        for kv, v in zip(args.var.split("+"), val.split("+")):   # several variables: A+B with values a+b
            os.environ[kv] = v
    ctx = engine.Context(0)
    ctx.attach_index(24, 32, sds, slots, blob.data_ptr(), seq.data_ptr())
    for s in range(3):
        ctx.upload(s, batches[s][0], offs, batches[s][1], offs if batches[s][1] is not None else None)
    for s in range(3):   # warm-up
        ctx.launch(s)
    ctx.mark(0)
    for k in range(args.steps):
        ctx.launch(k % 3)
    ctx.mark(1)
    ms = ctx.mark_elapsed() / args.steps
    tm = ctx.timing((args.steps - 1) % 3)
    ctx.download(0)
    x1, x2, runs = ctx.wait(0, B, not args.se)
    sig = (int(x1["db_pos"].astype(np.uint64).sum()), int(x1["mapq"].sum()), int(x1["score"].sum()))
    if ref is None:
        ref = sig
    n = B * (1 if args.se else 2)
    print(f"{args.var}={val or '-'}: {ms:.2f} ms/step -> {n / ms / 1e3:.2f} M reads/s  same_results={sig == ref} sig={sig}  "
          + " ".join(f"{k}={v:.1f}" for k, v in tm["kernel_ms"].items()), flush=True)
    ctx.close()
