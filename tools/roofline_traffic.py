#!/usr/bin/env python
"""Regenerate profiles/roofline_traffic.json from ncu launch lists.

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        -s <launches before the first step> -c <N> --csv --log-file gpurun_out/launches_<cfg>.csv \
        python bench.py --config <cfg> --steps 2 --warmup 1 --no-cpu-baseline --no-e2e
    python tools/roofline_traffic.py cfg1=profiles/r2_cfg1_launches.csv cfg3_r2=profiles/r2_cfg3_r2_launches.csv ...

For every configuration the tool cuts ONE steady-state step out of the list (from the second occurrence of the step's first
kernel to the next one), maps every kernel to its SVFM_PHASE_* phase and sums dram__bytes_read + dram__bytes_write per
kernel.  bench.py divides these bytes by the phase times it measures live with CUDA events (never under ncu) to report the
HBM throughput each phase really sustains.  tests/test_host_abi.py fails when a kernel named here no longer exists in the
sources, i.e. when the file has gone stale."""
from __future__ import annotations

import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIRST = re.compile(r"pack_sweep_kernel|pack_keys_kernel|unpack_patterns_kernel|search_kernel")
BENCH_ONLY = re.compile(r"verify_locate|synth_text|synth_patterns|gather32|count_digest|flush_l2|bwt_kernel|blocks_kernel|ck_chunk|ck_scan|"
                        r"ck_apply|sample_sa|find_pidx|tie_writeback|ext_level1|ext_expand|ilv_build|kmer_hist|DeviceSelect|"
                        r"DeviceCompact|at::|elementwise|vectorized")
UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3,
         "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}


def phase_of(name: str) -> str:
    if re.search(r"pack_sweep_kernel|pack_keys_kernel|unpack_patterns_kernel", name):
        return "presort"
    if "DeviceRadixSort" in name:
        return "presort" if ("SweepPay" in name or "policy_hub<unsigned long" in name) else "sortback"
    if re.search(r"sweep_round_kernel|search_kernel", name):
        return "search"
    if re.search(r"sb_scan_kernel|DeviceScan", name):
        return "scan"
    if re.search(r"locate_warp_kernel|locate_rows_kernel|locate_direct_kernel", name):
        return "locate"
    if re.search(r"sb_place_kernel|scatter_counts_kernel|run_starts_kernel|narrow_offs_kernel|add_base_kernel", name):
        return "sortback"
    return "other"


def short(name: str) -> str:
    m = re.match(r"(?:void )?((?:svfm|cub)::(?:<unnamed>::)?[A-Za-z0-9_]+)", name)
    return m.group(1).replace("<unnamed>::", "") if m else name.split("(")[0][:60]


def read_launches(path: str):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hi]
    ix = {n: i for i, n in enumerate(h)}
    launches, order = {}, []
    for r in rows[hi + 1:]:
        if len(r) < len(h):
            continue
        k = int(r[ix["ID"]])
        if k not in launches:
            launches[k] = {"name": r[ix["Kernel Name"]]}
            order.append(k)
        v = float(r[ix["Metric Value"]].replace(",", "")) * UNITS.get(r[ix["Metric Unit"]], 1.0)
        launches[k][r[ix["Metric Name"]]] = v
    return [launches[k] for k in order]


def one_step(launches):
    firsts = [i for i, l in enumerate(launches) if FIRST.search(l["name"]) and not BENCH_ONLY.search(l["name"])]
    # a step starts at a FIRST kernel that follows a non-FIRST kernel (unpack + pack are both FIRST kernels of the same step)
    starts = [i for i in firsts if i == 0 or not FIRST.search(launches[i - 1]["name"])]
    if len(starts) < 3:
        raise SystemExit(f"need at least three steps in the launch list, found {len(starts)}")
    a, b = starts[1], starts[2]
    seg = []
    for l in launches[a:b]:
        if BENCH_ONLY.search(l["name"]):
            break
        seg.append(l)
    return seg


def summarise(path: str, patterns_per_step: int):
    seg = one_step(read_launches(path))
    agg = {}
    for l in seg:
        key = (short(l["name"]), phase_of(l["name"]))
        e = agg.setdefault(key, {"launches_per_step": 0, "dram_bytes_per_step": 0.0, "ms_per_step_under_ncu": 0.0})
        e["launches_per_step"] += 1
        e["dram_bytes_per_step"] += l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
        e["ms_per_step_under_ncu"] += l.get("gpu__time_duration.sum", 0.0)
    kernels = [{"kernel": k, "phase": ph, "launches_per_step": v["launches_per_step"],
                "dram_bytes_per_step": round(v["dram_bytes_per_step"]), "ms_per_step_under_ncu": round(v["ms_per_step_under_ncu"], 4)}
               for (k, ph), v in agg.items()]
    return {"source": os.path.relpath(path, ROOT), "patterns_per_step": patterns_per_step, "kernels": kernels,
            "dram_bytes_per_step": round(sum(k["dram_bytes_per_step"] for k in kernels)),
            "ms_per_step_under_ncu": round(sum(k["ms_per_step_under_ncu"] for k in kernels), 4)}


def main():
    sys.path.insert(0, ROOT)
    import bench
    out_path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        doc = json.load(open(out_path))
        if "configs" not in doc:
            doc = {"configs": {}}
    except Exception:
        doc = {"configs": {}}
    doc["_comment"] = ("dram__bytes_read.sum + dram__bytes_write.sum per kernel over ONE steady-state step of `bench.py --config <cfg>`, "
                       "cut out of an ncu launch list by tools/roofline_traffic.py (the command is in its docstring); regenerate after "
                       "any kernel change")
    for arg in sys.argv[1:]:
        name, path = arg.split("=", 1)
        doc["configs"][name] = summarise(os.path.abspath(path), bench.CONFIGS[name]["batch"])
        c = doc["configs"][name]
        print(f"{name}: {len(c['kernels'])} kernels, {c['dram_bytes_per_step'] / 1e9:.2f} GB and {c['ms_per_step_under_ncu']:.2f} ms per step under ncu")
    json.dump(doc, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
