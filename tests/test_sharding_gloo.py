"""Multi-GPU host logic on CPU: world_size-2 gloo run of the pattern-sharded path.
Each rank owns a contiguous pattern range; the per-rank compute is played by the CPU oracle (a test stand-in
for the device), rank 0 gathers and merges, and the merged result must equal the unsharded one."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sview_fmindex_b200 import sharding


def test_shard_ranges_cover_and_balance():
    for n in (0, 1, 2, 7, 100, 1001):
        for world in (1, 2, 3, 8):
            rs = [sharding.shard_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_merge_csr_rebases_offsets():
    a = (np.array([0, 2, 2, 5], dtype=np.uint64), np.array([9, 8, 1, 2, 3], dtype=np.uint32))
    b = (np.array([0, 0, 1], dtype=np.uint64), np.array([7], dtype=np.uint32))
    offs, pos = sharding.merge_csr([a, b])
    assert list(offs) == [0, 2, 2, 5, 5, 6] and list(pos) == [9, 8, 1, 2, 3, 7]
    offs, pos = sharding.merge_csr([])
    assert list(offs) == [0] and pos.size == 0


def _worker(rank, world, port, blob_path, pats_path, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    t = po.IndexType(32, 3, 64, True)
    blob_raw = np.load(blob_path)
    blob = po.aligned_empty(blob_raw.size)
    blob[:] = blob_raw
    ora = po.OracleFmIndex.load(blob, t)          # "index replicated on every rank"
    pats = np.load(pats_path)
    mine = sharding.shard_patterns(pats, rank, world)
    counts, offs, pos, _ = ora.locate_batch(mine, threads=1)
    g_counts = sharding.gather_to_rank0(counts.astype(np.uint32))
    g_offs = sharding.gather_to_rank0(offs)
    g_pos = sharding.gather_to_rank0(pos)
    dist.barrier()
    if rank == 0:
        m_counts = sharding.merge_counts(g_counts)
        m_offs, m_pos = sharding.merge_csr(list(zip(g_offs, g_pos)))
        np.savez(out_path, counts=m_counts, offs=m_offs, pos=m_pos)
    dist.destroy_process_group()


def test_pattern_sharded_run_world2_gloo(tmp_path, oracle):
    po = oracle
    rng = np.random.default_rng(8)
    n = 60_000
    text = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)]
    table, sc = po.encoding_table([b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"])
    t = po.IndexType(32, 3, 64, True)
    blob = po.build_blob(t, text, sc, table, 3, 2)
    starts = rng.integers(0, n - 7, size=2001)        # odd count: uneven shards
    pats = text[starts[:, None] + np.arange(7)[None, :]].copy()
    pats[::9, 1] = ord("N")
    np.save(tmp_path / "blob.npy", np.asarray(blob))
    np.save(tmp_path / "pats.npy", pats)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(2, port, str(tmp_path / "blob.npy"), str(tmp_path / "pats.npy"), out), nprocs=2, join=True)
    got = np.load(out)
    ora = po.OracleFmIndex.load(blob, t)
    counts, offs, pos, _ = ora.locate_batch(pats, threads=2)
    assert np.array_equal(got["counts"].astype(np.uint64), counts)
    assert np.array_equal(got["offs"], offs)
    assert np.array_equal(got["pos"], pos)
