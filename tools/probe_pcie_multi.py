"""Aggregate host<->device copy bandwidth with k of the N ranks active (run under torchrun): the platform ceiling of
the end-to-end (host-buffer) path.  Each active rank copies 2 GB pinned -> device and 1.2 GB device -> pinned
concurrently, 5 times; rank 0 prints the aggregate GB/s for k = 1, 2, 4, ..., N."""
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
hp = torch.empty(2 * 10**9, dtype=torch.uint8).pin_memory()
dp = torch.empty(2 * 10**9, dtype=torch.uint8, device="cuda")
hq = torch.empty(12 * 10**8, dtype=torch.uint8).pin_memory()
dq = torch.empty(12 * 10**8, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


k = 1
while k <= world:
    for mode in ("h2d", "d2h", "both"):
        barrier()
        t0 = time.time()
        if rank < k:
            for _ in range(5):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s1):
                        dp.copy_(hp, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s2):
                        hq.copy_(dq, non_blocking=True)
        barrier()
        dt = time.time() - t0
        if rank == 0:
            gb = k * 5 * ((2.0 if mode in ("h2d", "both") else 0) + (1.2 if mode in ("d2h", "both") else 0))
            print(f"active ranks {k}: {mode:5s} {gb / dt:7.1f} GB/s aggregate ({dt * 1e3 / 5:.1f} ms per round)", flush=True)
    k *= 2
if world > 1:
    dist.destroy_process_group()
