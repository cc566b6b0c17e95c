#!/usr/bin/env python
"""bench.py -- count+locate patterns/s on the reference's benchmark configurations (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W [--config cfg1]       # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ... [--config]   # reference algorithm on the host cores

Default workload = the configuration BASELINE.json's metric is quoted on (cfg1: 1 Gbp random nucleotide text, seed 42, u32
positions, Block3<u64>, SA sampling ratio 2, kLTS 3, 10^8 x 20 bp patterns cut from the text per GPU and step, count +
locate).  The other BASELINE configurations are selected with --config (see CONFIGS).  The index is replicated on every GPU,
pattern batches are sharded over the GPUs, no collective on the data path.

A step = one batch through the hot path.  `value` times the device-resident pipeline (svfm_locate_batch_device /
svfm_count_batch_device, inputs already in HBM) with CUDA events on the session's stream; `e2e` times the host-buffer C-ABI
calls on pinned HOST buffers, copies inside the timed region.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "count+locate patterns/s"
UNIT = "patterns/s"
DNA5 = [b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"]   # bench/src/build/mod.rs:30 (reference bench CLI)
DNA4 = [b"Aa", b"Cc", b"Gg", b"Tt"]           # bench/src/build/sview_memory.rs:22-24,51 (T doubles as the wildcard)
AMINO = [bytes([c]) for c in b"ACDEFGHIKLMNPQRSTVWY"]

# BASELINE.json configs.  batch = patterns per GPU per step of the device-resident arm;
# e2e_batch = patterns per GPU per call of the host-buffer arm (cfg5: the per-GPU share of 10^9 patterns over 8 GPUs, streamed).
CONFIGS = {
    "cfg0": dict(baseline="configs[0]: README bench, 100,000 x 20 bp, in-memory count+locate", n=10**9, alphabet=b"ACGT", rare=0,
                 symbols=DNA5, wildcard=False, pos_bits=32, planes=3, vec_bits=64, k=3, r=2, plen=20, batch=100_000, mode="locate",
                 pack_bits=2),
    "cfg1": dict(baseline="configs[0]/[1] index, 10^8 x 20 bp per GPU and step, count+locate (the headline metric)", n=10**9,
                 alphabet=b"ACGT", rare=0, symbols=DNA5, wildcard=False, pos_bits=32, planes=3, vec_bits=64, k=3, r=2, plen=20,
                 batch=10**8, mode="locate", pack_bits=2),
    "cfg2": dict(baseline="configs[1]: same index, 10^8 x 20 bp count-only, pattern-sharded", n=10**9, alphabet=b"ACGT", rare=0,
                 symbols=DNA5, wildcard=False, pos_bits=32, planes=3, vec_bits=64, k=3, r=2, plen=20, batch=10**8, mode="count",
                 pack_bits=2),
    "cfg3_r2": dict(baseline="configs[2]: 1 Gbp Block2<u64> ACGT-only index, 150 bp reads, locate, SA ratio 2", n=10**9,
                    alphabet=b"ACGT", rare=0, symbols=DNA4, wildcard=False, pos_bits=32, planes=2, vec_bits=64, k=3, r=2, plen=150,
                    batch=10**7, mode="locate", pack_bits=2),
    "cfg3_r16": dict(baseline="configs[2]: 1 Gbp Block2<u64> ACGT-only index, 150 bp reads, locate, SA ratio 16", n=10**9,
                     alphabet=b"ACGT", rare=0, symbols=DNA4, wildcard=False, pos_bits=32, planes=2, vec_bits=64, k=3, r=16, plen=150,
                     batch=10**7, mode="locate", pack_bits=2),
    "cfg4_b5": dict(baseline="configs[3]: 500 Maa protein text, 20 symbols + wildcard, Block5<u64>, 12-mers, kLTS 3", n=5 * 10**8,
                    alphabet=b"ACDEFGHIKLMNPQRSTVWY", rare=1000, symbols=AMINO, wildcard=True, pos_bits=32, planes=5, vec_bits=64,
                    k=3, r=2, plen=12, batch=10**8, mode="locate", pack_bits=5),
    "cfg4_b6": dict(baseline="configs[3]: same on the widest Block<u64>, Block6<u64>", n=5 * 10**8,
                    alphabet=b"ACDEFGHIKLMNPQRSTVWY", rare=1000, symbols=AMINO, wildcard=True, pos_bits=32, planes=6, vec_bits=64,
                    k=3, r=2, plen=12, batch=10**8, mode="locate", pack_bits=5),
    "cfg5": dict(baseline="configs[4]: 3.1 Gbp, u64 positions, Block3<u64>, 10^9 x 32 bp count+locate across 8 GPUs (1.25 x 10^8 "
                          "patterns per GPU, streamed from host memory in the end-to-end arm)", n=31 * 10**8, alphabet=b"ACGT", rare=0,
                 symbols=DNA5, wildcard=False, pos_bits=64, planes=3, vec_bits=64, k=3, r=2, plen=32, batch=5 * 10**7, mode="locate",
                 pack_bits=2, e2e_batch=125_000_000),
}
PHASES = ["presort", "search", "scan", "locate", "segsort", "sortback"]   # SVFM_PHASE_* of include/svfm.h


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg1", choices=sorted(CONFIGS))
    ap.add_argument("--text-len", type=float, default=0, help="override the configuration's text length")
    ap.add_argument("--batch", type=float, default=0, help="override patterns per GPU per step")
    ap.add_argument("--pattern-len", type=int, default=0, help="override the pattern length")
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 5)")
    ap.add_argument("--ref-build", default="auto", choices=["auto", "device", "oracle"],
                    help="reference arm: who builds the blob (device = this repo's GPU builder in a child process, oracle = the "
                         "CPU restatement of the reference builder; auto = device when a GPU is visible)")
    return ap.parse_args()


def resolve_config(args):
    cfg = dict(CONFIGS[args.config])
    cfg["name"] = args.config
    if args.text_len:
        cfg["n"] = int(args.text_len)
    if args.batch:
        cfg["batch"] = int(args.batch)
    if args.pattern_len:
        cfg["plen"] = args.pattern_len
    cfg.setdefault("e2e_batch", cfg["batch"])
    return cfg


def chk(L, rc, what=""):
    if rc:
        raise RuntimeError(f"{what}: svfm rc={rc} {(L.svfm_last_error() or b'').decode(errors='replace')}")


# ------------------------------------------------------------------------------------------------
# work per pattern that occurs, the figures `roofline.achieved` is computed from
# ------------------------------------------------------------------------------------------------
def survey_8d(cfg, occ):
    L_, P_, k, r = cfg["plen"], cfg["pos_bits"] // 8, cfg["k"], cfg["r"]
    nb = cfg["planes"] * cfg["vec_bits"] // 8                       # bytes of one occ block
    Q = 2 * (L_ - k)                                                # rank queries
    W = occ * (r - 1)                                               # expected LF-walk steps
    s_b = {16: 1.0, 24: 1.5, 32: 1.0, 40: 2.0, 48: 2.25}.get(nb, nb / 32.0 + 0.5)
    a_count = L_ + 2 * P_ + Q * (P_ + nb) + P_
    a_locate = W * (nb + P_) + occ * 2 * P_
    sectors_count = Q * (1 + s_b)
    sectors_locate = W * (1 + s_b) + occ
    return {"Q": Q, "W": W, "block_bytes": nb, "a_count": a_count, "a_locate": a_locate,
            "sectors_count": sectors_count, "sectors_locate": sectors_locate}


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi in the background, timestamps taken on arrival)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 7:
                self.samples.append((time.time(), parts))

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, windows):
        sm, smax, reasons = [], 0.0, set()
        for t, p in self.samples:
            if not any(a <= t <= b for a, b in windows):
                continue
            try:
                sm.append(float(p[0]))
                smax = max(smax, float(p[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:  # region shorter than the sampling period: fall back to every sample
            for t, p in self.samples:
                try:
                    sm.append(float(p[0]))
                    smax = max(smax, float(p[1]))
                except ValueError:
                    pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# index construction on the device (bytes identical to the oracle builder: tests/test_gpu_builder.py)
# ------------------------------------------------------------------------------------------------
def make_encoder(fm, cfg):
    return fm.EncodingTable.from_symbols_with_wildcard(cfg["symbols"]) if cfg["wildcard"] else fm.EncodingTable.from_symbols(cfg["symbols"])


def build_index_on_device(L, torch, fm, cfg, seed, device):
    n = cfg["n"]
    t0 = time.time()
    d_text = torch.empty(n, dtype=torch.uint8, device=f"cuda:{device}")
    alpha = np.frombuffer(cfg["alphabet"], dtype=np.uint8)
    chk(L, L.svfm_bench_synth_text(d_text.data_ptr(), n, seed, alpha.ctypes.data, len(cfg["alphabet"]), cfg["rare"], ord("X"), None),
        "synth_text")
    enc = make_encoder(fm, cfg)
    it = fm.IndexType(cfg["pos_bits"], cfg["planes"], cfg["vec_bits"], True)
    b = fm.FmIndexBuilder(n, enc.symbol_count(), enc, it)
    b.kmer_size, b.sampling_ratio = cfg["k"], cfg["r"]
    size = b.blob_size()
    d_blob = torch.empty(size, dtype=torch.uint8, device=f"cuda:{device}")
    t1 = time.time()
    b.build_device(d_text.data_ptr(), d_blob.data_ptr(), size, device)
    torch.cuda.synchronize()
    t2 = time.time()
    return d_text, d_blob, size, it, enc, {"synth_text_s": round(t1 - t0, 3), "build_s": round(t2 - t1, 3)}


def bind_to_gpu_numa_node(gpu_index: int):
    """Pin this process (and the library's worker threads, which inherit the mask) to the CPUs NVML reports as local
    to the GPU, so that the pinned host buffers of the end-to-end path are allocated on the GPU's NUMA node and several
    ranks do not share one socket's memory controllers.  Best effort; returns the CPU count or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def load_traffic(config_name: str):
    """ncu DRAM bytes per step and kernel for this configuration (profiles/roofline_traffic.json, regenerated from an ncu
    launch list by tools/roofline_traffic.py), or None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        return d.get("configs", {}).get(config_name)
    except Exception:
        return None


def host_digest(offs: np.ndarray, pos: np.ndarray) -> int:
    """sum_i sum_p (p+1)*(2i+1) mod 2^64 over a CSR locate result (the digest svfm_bench_verify_locate and the oracle compute)."""
    n = len(offs) - 1
    total = np.uint64(0)
    step = 1 << 24
    with np.errstate(over="ignore"):
        for a in range(0, n, step):
            b = min(n, a + step)
            o = offs[a:b + 1].astype(np.int64)
            cnt = np.diff(o)
            idx = np.repeat(np.arange(a, b, dtype=np.uint64), cnt)
            p = pos[o[0]:o[-1]].astype(np.uint64)
            total += ((p + np.uint64(1)) * (np.uint64(2) * idx + np.uint64(1))).sum(dtype=np.uint64)
    return int(total)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import sview_fmindex_b200 as fm
    from sview_fmindex_b200 import _ffi

    cfg = resolve_config(args)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    L = _ffi.lib()
    n, B, plen = cfg["n"], cfg["batch"], cfg["plen"]
    K, W = args.steps, max(args.warmup, 0)
    mode = cfg["mode"]
    P_ = cfg["pos_bits"] // 8
    np_pos = np.uint32 if P_ == 4 else np.uint64
    t_pos = torch.int32 if P_ == 4 else torch.int64
    want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- index: built on this GPU; rank 0 also takes it through the reference-facing svfm_load (host blob) ----
    d_text, d_blob, blob_size, it, enc, build_info = build_index_on_device(L, torch, fm, cfg, args.seed, local_rank)
    host_blob = None
    if rank == 0 and (want_cpu or blob_size < (6 << 30)):
        t0 = time.time()
        host_blob = fm.aligned_empty(blob_size)
        torch.from_numpy(host_blob).copy_(d_blob)
        t1 = time.time()
        ix = fm.FmIndex.load(host_blob, it, device=local_rank)   # FmIndex::load(&blob): header checks + upload
        build_info["blob_d2h_s"] = round(t1 - t0, 3)
        build_info["load_from_host_s"] = round(time.time() - t1, 3)
        if not want_cpu:
            host_blob = None
    else:
        t1 = time.time()
        ix = fm.FmIndex.load_device(d_blob.data_ptr(), blob_size, it, device=local_rank)
        build_info["load_from_device_s"] = round(time.time() - t1, 3)
    del d_blob
    torch.cuda.empty_cache()
    info = ix.info()
    assert info.text_len == n
    index_memory = ix.memory()

    sess = C.c_void_p()
    chk(L, L.svfm_session_create(ix.handle, C.byref(sess)), "session_create")
    chk(L, L.svfm_session_set_timing(sess, 1))
    stream = torch.cuda.ExternalStream(L.svfm_session_stream(sess), device=f"cuda:{local_rank}")

    # ---- synthetic pattern batches, resident in HBM before the timed region ------------------------------
    free_b, _ = torch.cuda.mem_get_info()
    per_batch = B * plen + B * 8
    n_distinct = int(max(1, min(W + K, 8, (free_b * 0.30) // max(per_batch, 1))))
    batches = []
    for s in range(n_distinct):
        d_p = torch.empty(B * plen, dtype=torch.uint8, device="cuda")
        d_s = torch.empty(B, dtype=torch.int64, device="cuda")
        chk(L, L.svfm_bench_synth_patterns(d_text.data_ptr(), n, d_p.data_ptr(), d_s.data_ptr(), B, plen,
                                           args.seed + 1000 * (rank + 1) + s, None), "synth_patterns")
        batches.append((d_p, d_s))
    torch.cuda.synchronize()
    d_offs = torch.empty(B + 1, dtype=torch.int64, device="cuda")
    d_counts = torch.empty(B, dtype=t_pos, device="cuda")

    def locate_device(i, m=B):
        d_p, _ = batches[i % n_distinct]
        dpos, total = C.c_void_p(), C.c_uint64()
        chk(L, L.svfm_locate_batch_device(sess, d_p.data_ptr(), None, m, plen, 0, d_offs.data_ptr(), C.byref(dpos),
                                          C.byref(total)), "locate_batch_device")
        return dpos, total.value

    def count_device(i, m=B):
        chk(L, L.svfm_count_batch_device(sess, batches[i % n_distinct][0].data_ptr(), None, m, plen, 0, d_counts.data_ptr()),
            "count_batch_device")
        return None, 0

    step_device = locate_device if mode == "locate" else count_device

    # ---- gather roofline microbenchmark (same box, same run) ----------------------------
    ws_bytes = int((info.blob_len - info.off_rank_checkpoints) // 32 * 32)  # checkpoints + blocks
    gbuf = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    sps, gms = C.c_double(), C.c_double()
    chk(L, L.svfm_bench_gather32(gbuf.data_ptr(), ws_bytes, 1 << 28, 3, 7, C.byref(sps), C.byref(gms), None), "gather32")
    del gbuf
    G = sps.value

    sampler = ClockSampler(local_rank)
    windows = []
    # ---- warm-up + timed region: device-resident pipeline ------------------------------------------------------
    for w in range(W):
        step_device(w)
    chk(L, L.svfm_session_sync(sess))
    ms = (C.c_double * 8)()
    ln = (C.c_uint64 * 8)()
    chk(L, L.svfm_session_get_timing(sess, ms, ln, 1))
    barrier()
    launches0 = L.svfm_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tw0 = time.time()
    e0.record(stream)
    last = None
    for i in range(K):
        last = step_device(W + i)
    e1.record(stream)
    chk(L, L.svfm_session_sync(sess))
    torch.cuda.synchronize()
    tw1 = time.time()
    windows.append((tw0, tw1))
    dev_ms = e0.elapsed_time(e1)
    launches = L.svfm_launch_count() - launches0
    chk(L, L.svfm_session_get_timing(sess, ms, ln, 1))
    phase_ms = [ms[i] / K for i in range(6)]
    phase_launch = [int(ln[i]) // K for i in range(6)]
    barrier()
    dev_ms_max = max_over_ranks(dev_ms)
    value = world * B * K / (dev_ms_max * 1e-3)

    # ---- size-independent check of the last timed step (outside the timed region) ------------------------------
    d_p, d_s = batches[(W + K - 1) % n_distinct]
    viol = (C.c_uint64 * 3)()
    dig = C.c_uint64()
    if mode == "locate":
        chk(L, L.svfm_bench_verify_locate(d_text.data_ptr(), n, d_p.data_ptr(), plen, B, d_s.data_ptr(), d_offs.data_ptr(),
                                          last[0], cfg["pos_bits"], enc.table.ctypes.data, viol, C.byref(dig), None), "verify")
        occurrences = int(last[1])
        verified = {"patterns": B, "occurrences": occurrences, "bad_positions": int(viol[0]),
                    "missing_source_position": int(viol[1]), "empty_lists": int(viol[2])}
        if viol[0] or viol[1] or viol[2]:
            raise SystemExit(f"bench.py: result check FAILED {verified}")
    else:
        # count-only: the counts must equal the list lengths of a locate of the same batch
        csum, cdig = C.c_uint64(), C.c_uint64()
        chk(L, L.svfm_bench_count_digest(d_counts.data_ptr(), cfg["pos_bits"], B, C.byref(csum), C.byref(cdig), None))
        dpos, tot = locate_device(W + K - 1)
        chk(L, L.svfm_session_sync(sess))
        lens = (d_offs[1:] - d_offs[:-1])
        same = bool(torch.equal(lens.to(d_counts.dtype), d_counts))
        occurrences = int(tot)
        verified = {"patterns": B, "sum_of_counts": int(csum.value), "occurrences_located": occurrences,
                    "counts_equal_locate_list_lengths": same}
        if not same or int(csum.value) != occurrences:
            raise SystemExit(f"bench.py: result check FAILED {verified}")
    occ = occurrences / B

    # count-only throughput next to a locate run (configs[1] of BASELINE.json), reported as an extra
    count_only = None
    if mode == "locate":
        for w in range(2):
            count_device(w)
        chk(L, L.svfm_session_sync(sess))
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        kc = max(3, min(K, 10))
        for i in range(kc):
            count_device(i)
        c1.record(stream)
        chk(L, L.svfm_session_sync(sess))
        count_only = world * B * kc / (max_over_ranks(c0.elapsed_time(c1)) * 1e-3)
        chk(L, L.svfm_session_get_timing(sess, ms, ln, 1))

    # ---- end to end: the C-ABI host-buffer calls on pinned host memory -------------------------------------------
    e2e = e2e_ascii = None
    if not args.no_e2e:
        Ke = args.e2e_steps or min(K, 5)
        Be = int(cfg["e2e_batch"])
        bits = cfg["pack_bits"]
        bpp = (plen * bits + 7) // 8
        n_host = 2 if Be * plen * 2 < (24 << 30) else 1
        # host patterns: batch s of the e2e arm = Be patterns drawn like the device batches
        d_pe = torch.empty(Be * plen, dtype=torch.uint8, device="cuda")
        d_se = torch.empty(Be, dtype=torch.int64, device="cuda")
        h_ascii, h_packed = [], []
        for s in range(n_host):
            chk(L, L.svfm_bench_synth_patterns(d_text.data_ptr(), n, d_pe.data_ptr(), d_se.data_ptr(), Be, plen,
                                               args.seed + 77_000 * (rank + 1) + s, None), "synth_patterns(e2e)")
            torch.cuda.synchronize()
            p = L.svfm_host_alloc(Be * plen)
            q = L.svfm_host_alloc(Be * bpp)
            if not p or not q:
                raise SystemExit("bench.py: pinned allocation failed")
            arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(Be * plen,))
            torch.from_numpy(arr).copy_(d_pe)
            chk(L, L.svfm_pack_patterns(p, Be, plen, enc.table.ctypes.data, bits, q), "pack_patterns")  # the caller's packing: not timed
            h_ascii.append(p)
            h_packed.append(q)
        cap = int(Be * max(1.0, occ) * 1.25) + 1024
        h_offs_p = L.svfm_host_alloc((Be + 1) * 8)
        h_pos_p = L.svfm_host_alloc(cap * P_)
        h_cnt_p = L.svfm_host_alloc(Be * P_)
        if not h_offs_p or not h_pos_p or not h_cnt_p:
            raise SystemExit("bench.py: pinned allocation failed")
        total = C.c_uint64()

        def call_packed(i):
            if mode == "locate":
                chk(L, L.svfm_locate_batch_packed(ix.handle, h_packed[i % n_host], Be, plen, bits, _ffi.SVFM_OFFS32, h_offs_p, h_pos_p, cap,
                                                  C.byref(total)), "locate_batch_packed")
            else:
                chk(L, L.svfm_count_batch_packed(ix.handle, h_packed[i % n_host], Be, plen, bits, 0, h_cnt_p), "count_batch_packed")

        def call_ascii(i):
            if mode == "locate":
                chk(L, L.svfm_locate_batch(ix.handle, h_ascii[i % n_host], None, Be, plen, 0, h_offs_p, h_pos_p, cap, C.byref(total)),
                    "locate_batch")
            else:
                chk(L, L.svfm_count_batch(ix.handle, h_ascii[i % n_host], None, Be, plen, 0, h_cnt_p), "count_batch")

        def timed(call):
            for w in range(min(W, 2) or 1):
                call(w)
            barrier()
            t0 = time.time()
            for i in range(Ke):
                call(i)
            torch.cuda.synchronize()
            t1 = time.time()
            windows.append((t0, t1))
            return max_over_ranks(t1 - t0)

        def check_host_result(offs_dtype):
            """the host result of the last call must be the device result of the same batch: offsets AND a digest of the
            positions (sum (p+1)(2i+1) mod 2^64), which the device computes over its own copy"""
            i_last = (Ke - 1) % n_host
            chk(L, L.svfm_bench_synth_patterns(d_text.data_ptr(), n, d_pe.data_ptr(), d_se.data_ptr(), Be, plen,
                                               args.seed + 77_000 * (rank + 1) + i_last, None))
            torch.cuda.synchronize()   # the generator runs on the default stream, the session's stream does not wait for it
            if mode == "count":
                d_c = torch.empty(Be, dtype=t_pos, device="cuda")
                chk(L, L.svfm_count_batch_device(sess, d_pe.data_ptr(), None, Be, plen, 0, d_c.data_ptr()))
                chk(L, L.svfm_session_sync(sess))
                h_c = np.ctypeslib.as_array(C.cast(h_cnt_p, C.POINTER(C.c_uint32 if P_ == 4 else C.c_uint64)), shape=(Be,))
                return {"counts_equal": bool(np.array_equal(h_c, d_c.cpu().numpy().view(np_pos)))}
            d_o = torch.empty(Be + 1, dtype=torch.int64, device="cuda")
            dpos, tot = C.c_void_p(), C.c_uint64()
            chk(L, L.svfm_locate_batch_device(sess, d_pe.data_ptr(), None, Be, plen, 0, d_o.data_ptr(), C.byref(dpos), C.byref(tot)))
            chk(L, L.svfm_session_sync(sess))
            dd = C.c_uint64()
            v3 = (C.c_uint64 * 3)()
            chk(L, L.svfm_bench_verify_locate(d_text.data_ptr(), n, d_pe.data_ptr(), plen, Be, d_se.data_ptr(), d_o.data_ptr(), dpos,
                                              cfg["pos_bits"], enc.table.ctypes.data, v3, C.byref(dd), None))
            h_o = np.ctypeslib.as_array(C.cast(h_offs_p, C.POINTER(C.c_uint32 if offs_dtype == np.uint32 else C.c_uint64)), shape=(Be + 1,))
            h_p = np.ctypeslib.as_array(C.cast(h_pos_p, C.POINTER(C.c_uint32 if P_ == 4 else C.c_uint64)), shape=(int(total.value),))
            same_offs = bool(total.value == tot.value and np.array_equal(h_o.astype(np.uint64), d_o.cpu().numpy().astype(np.uint64)))
            hd = host_digest(h_o, h_p)
            return {"offsets_equal": same_offs, "positions_digest_host": hd, "positions_digest_device": int(dd.value),
                    "positions_digest_equal": hd == int(dd.value), "device_check_violations": [int(x) for x in v3]}

        def ok(c):
            return all(v for k, v in c.items() if k.endswith("_equal")) and not any(c.get("device_check_violations", []))

        wall = timed(call_packed)
        d2h = int((Be + 1) * 4 + total.value * P_) if mode == "locate" else Be * P_
        chk_p = check_host_result(np.uint32)
        e2e = {"value": world * Be * Ke / wall, "unit": UNIT, "h2d_bytes_per_step": Be * bpp, "d2h_bytes_per_step": d2h, "steps": Ke,
               "patterns_per_gpu_per_call": Be,
               "entry_point": ("svfm_locate_batch_packed(bits=%d, SVFM_OFFS32)" % bits) if mode == "locate" else "svfm_count_batch_packed(bits=%d)" % bits,
               "input": f"{bits}-bit packed symbol indices ({bpp} B per pattern, packed by the caller with svfm_pack_patterns before the "
                        f"timed region), u32 CSR offsets + positions out",
               "timer": "host wall clock around the C-ABI call (pinned host buffers, every copy inside), max over ranks",
               "matches_device_result": chk_p}
        if not ok(chk_p):
            raise SystemExit(f"bench.py: packed e2e result differs from the device-resident result {chk_p}")
        wall = timed(call_ascii)
        d2h = int((Be + 1) * 8 + total.value * P_) if mode == "locate" else Be * P_
        chk_a = check_host_result(np.uint64)
        e2e_ascii = {"value": world * Be * Ke / wall, "unit": UNIT, "h2d_bytes_per_step": Be * plen, "d2h_bytes_per_step": d2h, "steps": Ke,
                     "patterns_per_gpu_per_call": Be,
                     "entry_point": "svfm_locate_batch" if mode == "locate" else "svfm_count_batch",
                     "input": f"{plen} B ASCII per pattern (what the Rust wrapper passes), u64 CSR offsets + positions out",
                     "matches_device_result": chk_a}
        if not ok(chk_a):
            raise SystemExit(f"bench.py: e2e result differs from the device-resident result {chk_a}")
        for p in h_ascii + h_packed + [h_offs_p, h_pos_p, h_cnt_p]:
            L.svfm_host_free(p)
        del d_pe, d_se
    sampler.stop()
    clocks = sampler.summary(windows)

    # ---- CPU baseline: the oracle (reference algorithm) on the host cores, bounded sample, rank 0 at N=1 -------
    cpu_baseline = None
    if want_cpu:
        os.sched_setaffinity(0, all_cpus)  # the CPU baseline gets every host core again
        from oracle import pyoracle as po
        ora = po.OracleFmIndex.load(host_blob, po.IndexType(cfg["pos_bits"], cfg["planes"], cfg["vec_bits"], True))
        cores = os.cpu_count() or 1
        pats0 = batches[0][0]
        m0 = min(B, 100_000)
        probe = pats0[:m0 * plen].cpu().numpy().reshape(-1, plen)
        t0 = time.time()
        ora.locate_batch(probe, threads=cores, want_positions=False)
        rate = len(probe) / max(time.time() - t0, 1e-6)
        M = int(min(B, max(m0, rate * 12)))
        sample = pats0[:M * plen].cpu().numpy().reshape(-1, plen)
        t0 = time.time()
        ocounts, _, _, ock = ora.locate_batch(sample, threads=cores, want_positions=False)
        dt = time.time() - t0
        # parity of the same sample on the GPU (the oracle is the checker here, never the product path)
        chk(L, L.svfm_count_batch_device(sess, pats0.data_ptr(), None, M, plen, 0, d_counts.data_ptr()))
        chk(L, L.svfm_session_sync(sess))
        gcounts = d_counts[:M].cpu().numpy().view(np_pos).astype(np.uint64)
        dpos, total = C.c_void_p(), C.c_uint64()
        chk(L, L.svfm_locate_batch_device(sess, pats0.data_ptr(), None, M, plen, 0, d_offs.data_ptr(), C.byref(dpos), C.byref(total)))
        chk(L, L.svfm_session_sync(sess))
        gd = C.c_uint64()
        chk(L, L.svfm_bench_verify_locate(d_text.data_ptr(), n, pats0.data_ptr(), plen, M, None, d_offs.data_ptr(), dpos, cfg["pos_bits"],
                                          enc.table.ctypes.data, viol, C.byref(gd), None))
        parity = bool(np.array_equal(gcounts, ocounts) and gd.value == ock)
        cpu_baseline = {"value": M / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"first {M} patterns of step 0 ({plen} symbols), count+locate, one reference call per pattern "
                                  f"(oracle/fm_oracle.c), {cores} pthreads, {dt:.1f} s",
                        "gpu_bit_exact_on_sample": parity}
        if not parity:
            raise SystemExit("bench.py: GPU result differs from the CPU oracle on the baseline sample")

    # ---- roofline ---------------------------------------------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    s8 = survey_8d(cfg, occ)
    alg_pp = s8["a_count"] + (s8["a_locate"] if mode == "locate" else 0.0)   # SURVEY 8d algorithmic bytes per pattern
    sectors_pp = s8["sectors_count"] + (s8["sectors_locate"] if mode == "locate" else 0.0)
    dom = int(np.argmax(phase_ms))
    traffic = load_traffic(cfg["name"])
    per_phase = {}
    step_bytes = 0.0
    if traffic and traffic.get("patterns_per_step") == B:
        for ph, name in enumerate(PHASES):
            tb = sum(k["dram_bytes_per_step"] for k in traffic["kernels"] if k["phase"] == name)
            step_bytes += tb
            if phase_ms[ph] > 0 and tb:
                per_phase[name] = {"ms": round(phase_ms[ph], 4), "dram_gb_per_step": round(tb / 1e9, 3),
                                   "dram_gbs": round(tb / 1e9 / (phase_ms[ph] * 1e-3), 1),
                                   "dram_frac_of_peak": round(tb / 1e9 / (phase_ms[ph] * 1e-3) / peak, 4),
                                   "kernels": sorted({k["kernel"] for k in traffic["kernels"] if k["phase"] == name})}
    dom_name = PHASES[dom]
    dom_traffic = per_phase.get(dom_name, {}).get("dram_gb_per_step")
    dom_launches = max(phase_launch[dom], 1)
    # the phase's share of the SURVEY 8d bytes: the search phases carry the count part, locate + sortback the locate part
    a_dom = {"presort": s8["a_count"], "search": s8["a_count"], "locate": s8["a_locate"], "sortback": s8["a_locate"]}.get(dom_name, alg_pp)
    achieved = a_dom * B / (phase_ms[dom] * 1e-3) / 1e9 if phase_ms[dom] > 0 else 0.0
    roofline = {
        "bound": "hbm", "kernel": (per_phase.get(dom_name, {}).get("kernels") or [dom_name])[0] if per_phase else dom_name,
        "phase": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": dom_traffic * 1e9 / dom_launches if dom_traffic else None,
        "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
        "algorithmic_bytes_per_pattern": a_dom, "patterns_per_launch": B, "launches_per_batch": dom_launches,
        "kernel_ms_per_launch": phase_ms[dom] / dom_launches,
        "frac_is": "algorithmic_speedup_vs_naive: SURVEY 8d charges every rank query a private checkpoint word + block from DRAM and "
                   "counts all L-k steps; the engine resolves the last 12-14 symbols with one extended-table lookup and keeps the batch "
                   "in SA order so that patterns share index sectors, so this number says how much of the naive traffic was designed "
                   "away, NOT how busy HBM is -- for that read dram_frac_of_peak below",
        "dram_frac_of_peak": per_phase.get(dom_name, {}).get("dram_frac_of_peak"),
        "per_phase": per_phase or None,
        "whole_step": {"algorithmic_bytes_per_pattern": alg_pp, "algorithmic_gbs": alg_pp * B / (dev_ms_max / K * 1e-3) / 1e9,
                       "algorithmic_frac_of_peak": alg_pp * B / (dev_ms_max / K * 1e-3) / 1e9 / peak,
                       "dram_bytes_per_pattern_ncu": round(step_bytes / B, 1) if step_bytes else None,
                       "dram_frac_of_peak": round(step_bytes / 1e9 / (dev_ms_max / K * 1e-3) / peak, 4) if step_bytes else None,
                       "traffic_source": traffic.get("source") if traffic else None},
    }
    gather = {"sectors_per_s": G, "tb_per_s": G * 32 / 1e12, "working_set_bytes": ws_bytes, "sectors_per_pattern": sectors_pp,
              "bound_patterns_per_s": G / sectors_pp, "frac_of_gather_roofline": (value / world) / (G / sectors_pp)}

    if rank == 0:
        out = {
            "metric": METRIC if mode == "locate" else "count patterns/s", "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32" if P_ == 4 else "u64", "data": "synthetic",
            "config": {"workload": f"{cfg['name']} = BASELINE {cfg['baseline']}: {n} symbols uniform over {cfg['alphabet'].decode()} "
                                   f"(seed {args.seed}), u{cfg['pos_bits']} positions, Block{cfg['planes']}<u{cfg['vec_bits']}>, "
                                   f"S={enc.symbol_count()}, SA ratio {cfg['r']}, kLTS {cfg['k']}; {B} x {plen}-symbol patterns cut from "
                                   f"the text per GPU per step, {'count+locate (CSR offsets + positions)' if mode == 'locate' else 'count only'}",
                       "name": cfg["name"], "text_len": n, "patterns_per_gpu_per_step": B, "pattern_len": plen,
                       "blob_bytes": int(info.blob_len),
                       "sharding": f"index replicated, patterns sharded over {world} GPU(s), no collective",
                       "l2_policy": f"inputs larger than L2 ({B * plen / 1e9:.2f} GB of patterns + {info.blob_len / 1e9:.1f} GB index per step), "
                                    f"no flush" if B * plen > (256 << 20) else "batch smaller than L2: latency-bound configuration, no flush",
                       "distinct_batches": n_distinct, "cpus_bound_to_gpu_numa_node": numa},
            "clocks": clocks, "e2e": e2e, "e2e_ascii": e2e_ascii, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "gather_roofline": gather, "count_only_patterns_per_s": count_only,
            "phase_ms_per_step": dict(zip(PHASES, [round(x, 4) for x in phase_ms])),
            "phase_launches_per_step": dict(zip(PHASES, phase_launch)),
            "index_memory_bytes": index_memory,
            "verified": verified, "index_build": build_info,
        }
        print(json.dumps(out), flush=True)
    L.svfm_session_destroy(sess)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU algorithm (oracle port; the Rust crate cannot be built here)
# ------------------------------------------------------------------------------------------------
_CHILD_BUILD = r"""
import sys, json, numpy as np, torch
sys.path.insert(0, sys.argv[1])
import bench, sview_fmindex_b200 as fm
from sview_fmindex_b200 import _ffi
cfg = json.loads(sys.argv[2]); cfg["alphabet"] = cfg["alphabet"].encode(); cfg["symbols"] = [s.encode() for s in cfg["symbols"]]
L = _ffi.lib(); dev = int(sys.argv[4]); torch.cuda.set_device(dev)
d_text, d_blob, size, it, enc, binfo = bench.build_index_on_device(L, torch, fm, cfg, int(sys.argv[3]), dev)
np.save(sys.argv[5] + ".text.npy", d_text.cpu().numpy()); np.save(sys.argv[5] + ".blob.npy", d_blob.cpu().numpy())
"""


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    from sview_fmindex_b200 import synth   # numpy only: does not load libsvfm.so
    cfg = resolve_config(args)
    n, plen = cfg["n"], cfg["plen"]
    K, W = args.steps, max(args.warmup, 0)
    cores = os.cpu_count() or 1
    otype = po.IndexType(cfg["pos_bits"], cfg["planes"], cfg["vec_bits"], True)
    table, sc = po.encoding_table(cfg["symbols"], cfg["wildcard"])
    how = args.ref_build
    if how == "auto":
        try:
            how = "device" if subprocess.run(["nvidia-smi", "-L"], capture_output=True, timeout=20).returncode == 0 else "oracle"
        except Exception:
            how = "oracle"
    if how == "device":
        # the blob comes from this repo's GPU builder, run in a CHILD process: this process -- the one that is timed -- never
        # loads libsvfm.so.  (Bytes identical to the oracle builder: tests/test_gpu_builder.py.)
        import shutil
        need = 2 * (n + int(5.5 * n))
        tmp_dir = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > need else "/tmp"
        tmp = os.path.join(tmp_dir, f"svfm_ref_{os.getpid()}")
        jcfg = dict(cfg, alphabet=cfg["alphabet"].decode(), symbols=[s.decode() for s in cfg["symbols"]])
        subprocess.check_call([sys.executable, "-c", _CHILD_BUILD, ROOT, json.dumps(jcfg), str(args.seed),
                               os.environ.get("LOCAL_RANK", "0"), tmp])
        text = np.load(tmp + ".text.npy")
        raw = np.load(tmp + ".blob.npy")
        blob = po.aligned_empty(raw.size)
        blob[:] = raw
        del raw
        os.remove(tmp + ".text.npy")
        os.remove(tmp + ".blob.npy")
        built_by = "svfm_build_device in a child process (bytes identical to the oracle builder); the search runs on the CPU only"
    else:
        text = synth.synth_text(n, args.seed, cfg["alphabet"], cfg["rare"], ord("X"))
        blob = po.build_blob(otype, text, sc, table, cfg["k"], cfg["r"])
        built_by = "oracle builder on the host (CPU restatement of the reference's FmIndexBuilder)"
    ora = po.OracleFmIndex.load(blob, otype)
    want_pos = cfg["mode"] == "locate"
    # every position is computed (LF walk + sampled-SA read) and folded into a checksum instead of being stored: one search
    # per pattern like the reference's `locate`, minus the Vec pushes -- the faster, i.e. the more conservative, CPU arm
    probe, _ = synth.synth_patterns(text, 50_000, plen, args.seed + 999)
    t0 = time.time()
    ora.locate_batch(probe, threads=cores, want_positions=False)
    rate = len(probe) / max(time.time() - t0, 1e-6)
    budget_s = 150.0 / max(K + W, 1)
    M = int(max(50_000, rate * min(3.0, budget_s)))
    times = []
    occ = 0
    for s in range(W + K):
        pats, _ = synth.synth_patterns(text, M, plen, args.seed + 1000 + s)
        t0 = time.time()
        counts, _, _, _ = ora.locate_batch(pats, threads=cores, want_positions=False)
        dt = time.time() - t0
        if s >= W:
            times.append(dt)
            occ += int(counts.sum())
    total_t = sum(times)
    value = M * K / total_t
    out = {
        "impl": "reference", "metric": METRIC if want_pos else "count patterns/s", "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": K, "warmup": W, "ms_per_step": total_t / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32" if cfg["pos_bits"] == 32 else "u64", "data": "synthetic",
        "config": {"workload": f"{cfg['name']} = BASELINE {cfg['baseline']}: {n} symbols uniform over {cfg['alphabet'].decode()} (seed "
                               f"{args.seed}), u{cfg['pos_bits']} positions, Block{cfg['planes']}<u{cfg['vec_bits']}>, S={sc}, SA ratio "
                               f"{cfg['r']}, kLTS {cfg['k']}; {'count+locate (every position computed and checksummed)' if want_pos else 'count'} of {plen}-symbol "
                               f"patterns cut from the text; each step = bounded sample of {M} patterns",
                   "name": cfg["name"], "text_len": n, "pattern_len": plen, "patterns_per_step": M, "index_built_by": built_by},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{M} patterns per step, {K} steps, one reference call per pattern, {cores} pthreads; the Rust "
                                   f"crate cannot be compiled in this image (no cargo/rustc), so the C restatement oracle/fm_oracle.c runs"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "occurrences": occ,
    }
    print(json.dumps(out), flush=True)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
