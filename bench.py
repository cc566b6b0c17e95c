#!/usr/bin/env python
"""bench.py -- count+locate patterns/s on the reference's headline configuration (BASELINE.json):
1 Gbp random nucleotide text (seed 42), u32 positions, Block3<u64>, SA sampling ratio 2, kLTS 3,
20 bp patterns cut from the text, pattern batches sharded over the GPUs (index replicated, no collective).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # reference algorithm on the host cores

A step = one `locate` batch (count_i = out_offs[i+1]-out_offs[i] and the positions of every occurrence) over
`--batch` patterns per GPU.  `value` times the device-resident pipeline (svfm_locate_batch_device, inputs already
in HBM) with CUDA events on the session stream; `e2e` times the reference-facing C-ABI call
(svfm_locate_batch) on pinned HOST buffers, copies included.  Prints ONE JSON line on rank 0.
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

SYMBOLS = [b"Aa", b"Cc", b"Gg", b"Tt", b"Nn"]  # bench/src/build/mod.rs:30 (reference bench CLI)
METRIC = "count+locate patterns/s"
UNIT = "patterns/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--text-len", type=float, default=1e9)
    ap.add_argument("--batch", type=float, default=1e8, help="patterns per GPU per step")
    ap.add_argument("--pattern-len", type=int, default=20)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = same as --steps")
    return ap.parse_args()


def chk(L, rc, what=""):
    if rc:
        raise RuntimeError(f"{what}: svfm rc={rc} {(L.svfm_last_error() or b'').decode(errors='replace')}")


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
# index construction shared by both arms (device build; bytes identical to the oracle builder,
# tests/test_gpu_builder.py)
# ------------------------------------------------------------------------------------------------
def build_index_on_device(L, torch, fm, n, seed, device):
    from sview_fmindex_b200 import synth
    t0 = time.time()
    d_text = torch.empty(n, dtype=torch.uint8, device=f"cuda:{device}")
    alpha = np.frombuffer(synth.NUCLEOTIDES, dtype=np.uint8)
    chk(L, L.svfm_bench_synth_text(d_text.data_ptr(), n, seed, alpha.ctypes.data, 4, 0, 0, None), "synth_text")
    enc = fm.EncodingTable.from_symbols(SYMBOLS)
    it = fm.IndexType(32, 3, 64, True)
    b = fm.FmIndexBuilder(n, enc.symbol_count(), enc, it)
    b.kmer_size, b.sampling_ratio = 3, 2
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


def roofline_traffic(kernel: str, batch: int):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture of this same workload (profiles/roofline_traffic.json), or None."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        e = d.get(kernel, {}).get(str(batch))
        return float(e) if e is not None else None
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import sview_fmindex_b200 as fm
    from sview_fmindex_b200 import _ffi

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
    n = int(args.text_len)
    B = int(args.batch)
    plen = args.pattern_len
    K, W = args.steps, max(args.warmup, 0)

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

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- index: built on this GPU, then loaded through the reference-facing svfm_load on rank 0 -------------
    d_text, d_blob, blob_size, it, enc, build_info = build_index_on_device(L, torch, fm, n, args.seed, local_rank)
    host_blob = None
    if rank == 0:
        t0 = time.time()
        host_blob = fm.aligned_empty(blob_size)
        chk(L, 0 if torch.from_numpy(host_blob).copy_(d_blob) is not None else 1)
        t1 = time.time()
        ix = fm.FmIndex.load(host_blob, it, device=local_rank)   # FmIndex::load(&blob): header checks + upload
        build_info["blob_d2h_s"] = round(t1 - t0, 3)
        build_info["load_from_host_s"] = round(time.time() - t1, 3)
    else:
        ix = fm.FmIndex.load_device(d_blob.data_ptr(), blob_size, it, device=local_rank)
    del d_blob
    torch.cuda.empty_cache()
    info = ix.info()
    assert info.text_len == n

    sess = C.c_void_p()
    chk(L, L.svfm_session_create(ix.handle, C.byref(sess)), "session_create")
    chk(L, L.svfm_session_set_timing(sess, 1))
    stream = torch.cuda.ExternalStream(L.svfm_session_stream(sess), device=f"cuda:{local_rank}")

    # ---- synthetic pattern batches, resident in HBM before the timed region ------------------------------
    free_b, _ = torch.cuda.mem_get_info()
    per_batch = B * plen + B * 8
    n_distinct = int(max(1, min(W + K, 8, (free_b * 0.35) // max(per_batch, 1))))
    batches = []
    for s in range(n_distinct):
        d_p = torch.empty(B * plen, dtype=torch.uint8, device="cuda")
        d_s = torch.empty(B, dtype=torch.int64, device="cuda")
        chk(L, L.svfm_bench_synth_patterns(d_text.data_ptr(), n, d_p.data_ptr(), d_s.data_ptr(), B, plen,
                                           args.seed + 1000 * (rank + 1) + s, None), "synth_patterns")
        batches.append((d_p, d_s))
    torch.cuda.synchronize()
    d_offs = torch.empty(B + 1, dtype=torch.int64, device="cuda")
    d_counts = torch.empty(B, dtype=torch.int32, device="cuda")

    def step_device(i):
        d_p, _ = batches[i % n_distinct]
        dpos, total = C.c_void_p(), C.c_uint64()
        chk(L, L.svfm_locate_batch_device(sess, d_p.data_ptr(), None, B, plen, 0, d_offs.data_ptr(), C.byref(dpos),
                                          C.byref(total)), "locate_batch_device")
        return dpos, total.value

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
    chk(L, L.svfm_bench_verify_locate(d_text.data_ptr(), n, d_p.data_ptr(), plen, B, d_s.data_ptr(), d_offs.data_ptr(),
                                      last[0], 32, enc.table.ctypes.data, viol, C.byref(dig), None), "verify")
    verified = {"patterns": B, "occurrences": int(last[1]), "bad_positions": int(viol[0]), "missing_source_position": int(viol[1]),
                "empty_lists": int(viol[2])}
    if viol[0] or viol[1] or viol[2]:
        raise SystemExit(f"bench.py: result check FAILED {verified}")

    # count-only throughput (configs[1] of BASELINE.json), reported as an extra
    for w in range(2):
        chk(L, L.svfm_count_batch_device(sess, batches[w % n_distinct][0].data_ptr(), None, B, plen, 0, d_counts.data_ptr()))
    chk(L, L.svfm_session_sync(sess))
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(stream)
    kc = max(3, min(K, 10))
    for i in range(kc):
        chk(L, L.svfm_count_batch_device(sess, batches[i % n_distinct][0].data_ptr(), None, B, plen, 0, d_counts.data_ptr()))
    c1.record(stream)
    chk(L, L.svfm_session_sync(sess))
    count_only = world * B * kc / (max_over_ranks(c0.elapsed_time(c1)) * 1e-3)
    chk(L, L.svfm_session_get_timing(sess, ms, ln, 1))

    # ---- end to end: reference-facing C-ABI call on pinned host buffers ----------------------------------------
    e2e = None
    if not args.no_e2e:
        Ke = args.e2e_steps or K
        n_host = min(n_distinct, 2)
        cap = B + B // 4 + 1024
        h_pats = []
        for s in range(n_host):
            p = L.svfm_host_alloc(B * plen)
            if not p:
                raise SystemExit("bench.py: pinned allocation failed")
            arr = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(B * plen,))
            torch.from_numpy(arr).copy_(batches[s][0])
            h_pats.append((p, arr))
        h_offs_p = L.svfm_host_alloc((B + 1) * 8)
        h_pos_p = L.svfm_host_alloc(cap * 4)
        h_offs = np.ctypeslib.as_array(C.cast(h_offs_p, C.POINTER(C.c_uint64)), shape=(B + 1,))
        h_pos = np.ctypeslib.as_array(C.cast(h_pos_p, C.POINTER(C.c_uint32)), shape=(cap,))
        total = C.c_uint64()

        def step_host(i):
            chk(L, L.svfm_locate_batch(ix.handle, h_pats[i % n_host][0], None, B, plen, 0, h_offs_p, h_pos_p, cap,
                                       C.byref(total)), "locate_batch")

        for w in range(min(W, 2) or 1):
            step_host(w)
        barrier()
        t0 = time.time()
        for i in range(Ke):
            step_host(i)
        torch.cuda.synchronize()
        t1 = time.time()
        windows.append((t0, t1))
        wall = max_over_ranks(t1 - t0)
        # the host result must be the device result of the same batch
        i_last = (Ke - 1) % n_host
        step_device(i_last)
        chk(L, L.svfm_session_sync(sess))
        same = bool(np.array_equal(h_offs, d_offs.cpu().numpy().astype(np.uint64)))
        e2e = {"value": world * B * Ke / wall, "unit": UNIT, "h2d_bytes_per_step": B * plen,
               "d2h_bytes_per_step": int((B + 1) * 8 + total.value * 4), "steps": Ke,
               "timer": "host wall clock around svfm_locate_batch (pinned host buffers, copies inside), max over ranks",
               "matches_device_result": same}
        if not same:
            raise SystemExit("bench.py: e2e result differs from the device-resident result")
        for p, _ in h_pats:
            L.svfm_host_free(p)
        host_result = (h_offs, h_pos, int(total.value))
    sampler.stop()
    clocks = sampler.summary(windows)

    # ---- CPU baseline: the oracle (reference algorithm) on the host cores, bounded sample, rank 0 at N=1 -------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)  # the CPU baseline gets every host core again
        from oracle import pyoracle as po
        ora = po.OracleFmIndex.load(host_blob, po.IndexType(32, 3, 64, True))
        cores = os.cpu_count() or 1
        pats0 = batches[0][0]
        probe = pats0[:200_000 * plen].cpu().numpy().reshape(-1, plen)
        t0 = time.time()
        ora.locate_batch(probe, threads=cores, want_positions=False)
        rate = len(probe) / max(time.time() - t0, 1e-6)
        M = int(min(B, max(200_000, rate * 12)))
        sample = pats0[:M * plen].cpu().numpy().reshape(-1, plen)
        t0 = time.time()
        ocounts, _, _, ock = ora.locate_batch(sample, threads=cores, want_positions=False)
        dt = time.time() - t0
        # parity of the same sample on the GPU (the oracle is the checker here, never the product path)
        chk(L, L.svfm_count_batch_device(sess, pats0.data_ptr(), None, M, plen, 0, d_counts.data_ptr()))
        chk(L, L.svfm_session_sync(sess))
        gcounts = d_counts[:M].cpu().numpy().astype(np.uint64)
        dpos, total = C.c_void_p(), C.c_uint64()
        chk(L, L.svfm_locate_batch_device(sess, pats0.data_ptr(), None, M, plen, 0, d_offs.data_ptr(), C.byref(dpos), C.byref(total)))
        chk(L, L.svfm_session_sync(sess))
        gd = C.c_uint64()
        chk(L, L.svfm_bench_verify_locate(d_text.data_ptr(), n, pats0.data_ptr(), plen, M, None, d_offs.data_ptr(), dpos, 32,
                                          enc.table.ctypes.data, viol, C.byref(gd), None))
        parity = bool(np.array_equal(gcounts, ocounts) and gd.value == ock)
        cpu_baseline = {"value": M / dt, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"first {M} patterns of step 0 ({plen} bp), count+locate, one reference call per pattern "
                                  f"(oracle/fm_oracle.c), {cores} pthreads, {dt:.1f} s",
                        "gpu_bit_exact_on_sample": parity}
        if not parity:
            raise SystemExit("bench.py: GPU result differs from the CPU oracle on the baseline sample")

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------
    names = ["presort(pack_sweep+radix sort by table index)", "search(sweep_round_kernel)", "scan", "locate_warp_kernel", "segsort",
             "sortback(radix sort by pattern index + CSR offsets)"]
    kernel_of = {0: "pack_sweep_kernel+cub onesweep", 1: "sweep_round_kernel", 3: "locate_warp_kernel", 5: "cub onesweep"}
    dom = int(np.argmax(phase_ms))
    P_, Nb = 4, 24
    Q = 2 * (plen - 3)
    occ = verified["occurrences"] / B
    alg_bytes = {1: plen + 2 * P_ + Q * (P_ + Nb) + P_,           # SURVEY.md section 8d, count part: 984 B at L=20
                 3: occ * (1 * (Nb + P_) + 2 * P_),                # locate part: W=occ*(r-1) LF steps + SA read + output
                 0: 2 * 4 * 2 * 4 + plen,                          # pattern bytes + (table index, item) through 3 radix passes
                 5: 2 * 8 * 4}                                     # (pattern index, position) records through 4 radix passes
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    a_bytes = alg_bytes.get(dom, 0.0) * B
    n_launch = max(phase_launch[dom], 1)
    achieved = a_bytes / (phase_ms[dom] * 1e-3) / 1e9 if phase_ms[dom] > 0 else 0.0
    traffic = roofline_traffic(kernel_of.get(dom, names[dom]), B)
    roofline = {"bound": "hbm", "kernel": kernel_of.get(dom, names[dom]), "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                "algorithmic_bytes_per_pattern": alg_bytes.get(dom), "patterns_per_launch": B,
                "launches_per_batch": n_launch, "kernel_ms_per_launch": phase_ms[dom] / n_launch,
                "algorithmic_bytes_per_launch": a_bytes / n_launch,
                "note": "algorithmic bytes = SURVEY.md 8d per-pattern figure (every rank query counted as a private checkpoint word + "
                        "block, 17 steps from the blob's k=3 table) x patterns, spread evenly over the launches of one batch; "
                        "the engine resolves the last 12-14 symbols with one extended-table lookup and keeps the batch in SA order so that "
                        "patterns share index sectors, hence frac > 1; `traffic` (ncu dram bytes per launch) / kernel_ms_per_launch "
                        "is the HBM throughput the kernel really sustains"}
    if traffic:
        roofline["dram_gbs_sustained"] = traffic / 1e9 / (phase_ms[dom] / n_launch * 1e-3)
        roofline["dram_frac_of_peak"] = roofline["dram_gbs_sustained"] / peak
    sectors_per_pattern = (Q + occ * 1) * 2.5 + occ   # (Q+W)*(1+1.5) + occ
    gather = {"sectors_per_s": G, "tb_per_s": G * 32 / 1e12, "working_set_bytes": ws_bytes,
              "sectors_per_pattern": sectors_per_pattern, "bound_patterns_per_s": G / sectors_per_pattern,
              "frac_of_gather_roofline": (value / world) / (G / sectors_per_pattern)}

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[0]/[1] index: {n} bp uniform ACGT (seed {args.seed}), u32 positions, "
                                   f"Block3<u64>, symbols Aa,Cc,Gg,Tt,Nn (S=5), SA ratio 2, kLTS 3; {B} x {plen} bp patterns "
                                   f"cut from the text per GPU per step, count+locate (CSR offsets + positions)",
                       "text_len": n, "patterns_per_gpu_per_step": B, "pattern_len": plen, "blob_bytes": int(info.blob_len),
                       "sharding": f"index replicated, patterns sharded over {world} GPU(s), no collective",
                       "l2_policy": "inputs larger than L2 (2 GB of patterns + 2.7 GB index per step), no flush",
                       "distinct_batches": n_distinct, "cpus_bound_to_gpu_numa_node": numa},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "gather_roofline": gather, "count_only_patterns_per_s": count_only,
            "phase_ms_per_step": dict(zip(names, [round(x, 4) for x in phase_ms])),
            "phase_launches_per_step": dict(zip(names, phase_launch)),
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
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import pyoracle as po
    from sview_fmindex_b200 import synth
    n = int(args.text_len)
    plen = args.pattern_len
    K, W = args.steps, max(args.warmup, 0)
    cores = os.cpu_count() or 1
    built_by = None
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        import torch

        import sview_fmindex_b200 as fm
        from sview_fmindex_b200 import _ffi
        L = _ffi.lib()
        dev = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(dev)
        d_text, d_blob, size, it, enc, binfo = build_index_on_device(L, torch, fm, n, args.seed, dev)
        blob = po.aligned_empty(size)
        torch.from_numpy(blob).copy_(d_blob)
        text = d_text.cpu().numpy()
        del d_text, d_blob
        torch.cuda.empty_cache()
        built_by = "svfm_build_device (bytes identical to the oracle builder: tests/test_gpu_builder.py); search runs on the CPU only"
    else:
        text = synth.synth_text(n, args.seed, synth.NUCLEOTIDES)
        table, sc = po.encoding_table(SYMBOLS)
        blob = po.build_blob(po.IndexType(32, 3, 64, True), text, sc, table, 3, 2)
        built_by = "oracle builder on the host (no GPU visible)"
    ora = po.OracleFmIndex.load(blob, po.IndexType(32, 3, 64, True))
    probe, _ = synth.synth_patterns(text, 100_000, plen, args.seed + 999)
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
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
        "ms_per_step": total_t / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"{n} bp uniform ACGT (seed {args.seed}), u32 positions, Block3<u64>, S=5, SA ratio 2, kLTS 3; "
                               f"count+locate of {plen} bp patterns cut from the text; each step = bounded sample of {M} patterns",
                   "text_len": n, "pattern_len": plen, "patterns_per_step": M, "index_built_by": built_by},
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
