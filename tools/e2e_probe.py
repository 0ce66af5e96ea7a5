"""Where does the end-to-end step lose time against the kernels-only step?  Pipelines of K steps over 6 slots on the bench
workload with the host-side pieces switched on one at a time (run under gpurun):
  launch            kernels only, inputs resident (what `value` times)
  launch+d2h        + result download and wait per step
  h2d+launch        + upload per step, no download
  submit            upload + launch + download + wait (what `e2e` times)"""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from urmap_b200 import engine

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=40)
a = ap.parse_args()
dev = torch.device("cuda", 0)
args = argparse.Namespace(genome_len=3_100_000_000, pairs_per_step=1_000_000, read_len=150, sub=0.01, indel=0.001,
                          single_end=False, segdup_frac=0.03, tandem_frac=0.01)
meta, seq, blob = bench.build_workload(args, 0, 1, dev)
batches = bench.make_batches(meta, seq, dev, 3, 1_000_000, True, 150, 0.01, 0.001, seed0=1000)
ctx = engine.Context(0)
ctx.attach_index(meta["word_length"], meta["max_ix"], meta["seq_data_size"], meta["slot_count"], blob.data_ptr(), seq.data_ptr())
NS, K = 6, a.steps
B = len(batches[0][4]) - 1
def up(k):
    _, _, a1, a2, offs = batches[k % 3]
    ctx.upload(k % NS, a1, offs, a2, offs)
for mode in ("launch", "launch+d2h", "h2d+launch", "submit", "launch"):
    for s in range(NS):
        up(s)
    for s in range(NS):
        ctx.launch(s); ctx.download(s); ctx.wait(s, B, True)
    for s in range(NS):
        up(s)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(K):
        s = k % NS
        if mode == "launch":
            ctx.launch(s)
        elif mode == "launch+d2h":
            if k >= NS:
                ctx.wait(s, B, True)
            ctx.launch(s); ctx.download(s)
        elif mode == "h2d+launch":
            if k >= NS:
                ctx.download(s); ctx.wait(s, B, True)   # the slot must have drained before it is restaged
                up(k)
            ctx.launch(s)
        else:
            if k >= NS:
                ctx.wait(s, B, True)
            _, _, a1, a2, offs = batches[k % 3]
            ctx.submit(s, a1, offs, a2, offs)
    for s in range(NS):
        try:
            ctx.download(s); ctx.wait(s, B, True)
        except Exception:
            pass
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{mode:12s}: {1e3 * dt / K:.2f} ms/step over {K} steps", flush=True)
