"""One-process-per-GPU plumbing (torch.distributed): index replication and read sharding.

The mapping path shards by reads with NO data-path collective: the only collective is the
one-off broadcast of the UFI index (blob + sequence data) from rank 0 into every GPU's HBM over NVLink."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend=None):
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_range(n_units: int, rank: int, world: int):
    """Contiguous, balanced split of n_units reads/pairs; pairs never straddle ranks."""
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_bytes(t: torch.Tensor, src=0, chunk=1 << 30):
    """Broadcast a (possibly tens-of-GB) uint8 tensor in place, chunked so that no single collective is huge."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return t
    flat = t.view(-1)
    for o in range(0, flat.numel(), chunk):
        dist.broadcast(flat[o:o + chunk], src=src)
    return t


def broadcast_index(meta, seq: torch.Tensor | None, blob: torch.Tensor | None, device, src=0):
    """meta: dict with word_length, max_ix, seq_data_size, slot_count, names, lens, offsets (only needed on src).
    Returns (meta, seq, blob) on every rank; non-src ranks allocate the device buffers."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return meta, seq, blob
    box = [meta if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    meta = box[0]
    if dist.get_rank() != src:
        seq = torch.empty(meta["seq_alloc"], dtype=torch.uint8, device=device)
        blob = torch.empty(meta["blob_alloc"], dtype=torch.uint8, device=device)
    broadcast_bytes(seq, src)
    broadcast_bytes(blob, src)
    return meta, seq, blob


def max_over_ranks(x: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def finalize():
    """Tear the process group down at the end of a run (NCCL warns about leaked resources otherwise)."""
    if dist.is_initialized():
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:
            pass
