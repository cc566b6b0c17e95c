"""Multi-GPU host logic: the hot path shards by pattern.  The index is replicated on every
GPU, each rank searches a contiguous range of the batch, and results return to the host of rank 0 in the caller's
pattern order.  No collective on the data path; torch.distributed (gloo on CPU, nccl on GPUs) only carries the
final host-side gather when a single process wants the whole result."""
from __future__ import annotations

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of n patterns: ranks differ by at most one pattern."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return n * rank // world, n * (rank + 1) // world


def shard_patterns(patterns, rank: int, world: int):
    """Slice of a batch owned by `rank`: a 2-D uint8 array (fixed length) or a list of byte strings."""
    a, b = shard_range(len(patterns), rank, world)
    return patterns[a:b]


def merge_counts(parts) -> np.ndarray:
    """Concatenate per-rank count arrays in rank order."""
    parts = [np.asarray(p) for p in parts]
    return np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint32)


def merge_csr(parts):
    """Concatenate per-rank CSR results [(offs u64[n_r+1], positions P[total_r]), ...] into one CSR result:
    every rank's offsets are rebased by the number of positions of the ranks before it."""
    offs_out = [np.zeros(1, dtype=np.uint64)]
    pos_out = []
    base = np.uint64(0)
    for offs, pos in parts:
        offs = np.asarray(offs, dtype=np.uint64)
        offs_out.append(offs[1:] + base)
        pos_out.append(np.asarray(pos))
        base = base + offs[-1]
    positions = np.concatenate(pos_out) if pos_out else np.zeros(0, dtype=np.uint32)
    return np.concatenate(offs_out), positions


def gather_to_rank0(local: np.ndarray, group=None):
    """Variable-length gather of a 1-D numpy array to rank 0 over torch.distributed (any backend).
    Returns the list of per-rank arrays on rank 0, None elsewhere."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    raw = np.ascontiguousarray(local).view(np.uint8).reshape(-1)
    size = torch.tensor([raw.size], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, size, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(sizes) if sizes else 0
    buf = torch.zeros(max(cap, 1), dtype=torch.uint8, device=dev)
    if raw.size:
        buf[:raw.size] = torch.from_numpy(raw.copy()).to(dev)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)  # all_gather keeps the code backend-agnostic (gloo has no GPU gather)
    if rank != 0:
        return None
    return [b[:s].cpu().numpy().view(local.dtype) for b, s in zip(bufs, sizes)]
