"""N>1 host logic on CPU: world_size-2 gloo run of the sharding / index-replication plumbing (urmap_b200/dist.py)."""
import os
import subprocess
import sys
import textwrap

from conftest import ROOT


def test_shard_range_balanced():
    from urmap_b200.dist import shard_range
    for n in (0, 1, 7, 10, 1000001):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        import torch
        from urmap_b200 import dist as D
        rank, local_rank, world = D.init(backend="gloo")
        assert world == 2
        meta = seq = blob = None
        if rank == 0:
            seq = torch.arange(5000, dtype=torch.int64).to(torch.uint8)
            blob = (torch.arange(12345, dtype=torch.int64) * 7).to(torch.uint8)
            meta = {{"slot_count": 2469, "seq_data_size": 5000, "seq_alloc": 5000, "blob_alloc": 12345}}
        meta, seq, blob = D.broadcast_index(meta, seq, blob, "cpu")
        assert meta["slot_count"] == 2469
        assert int(seq.sum()) == int(torch.arange(5000, dtype=torch.int64).to(torch.uint8).sum())
        assert int(blob[100]) == (100 * 7) % 256
        lo, hi = D.shard_range(1001, rank, world)
        tot = D.sum_over_ranks(float(hi - lo), "cpu")
        mx = D.max_over_ranks(float(rank + 1), "cpu")
        assert tot == 1001.0 and mx == 2.0
        D.barrier()
        print("rank", rank, "ok")
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", str(script)], capture_output=True, text=True, env=env,
                       timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2
