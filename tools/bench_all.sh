#!/bin/bash
# Developer helper (GPU box): one bench line per BASELINE configuration -> gpurun_out/bench_<cfg>.json
#   tools/bench_all.sh [cfg ...]     BENCH_ARGS overrides "--steps 5 --warmup 3"
mkdir -p gpurun_out
ARGS=${BENCH_ARGS:---steps 5 --warmup 3}
CFGS=${@:-cfg1 cfg0 cfg2 cfg3_r2 cfg3_r16 cfg4_b5 cfg4_b6 cfg5}
for c in $CFGS; do
  timeout 600 python bench.py --config $c $ARGS > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err || { echo "$c FAILED rc=$?"; tail -4 gpurun_out/bench_$c.err; continue; }
  python - $c <<PY
import json,sys
d=json.load(open("gpurun_out/bench_%s.json"%sys.argv[1]))
e=d.get("e2e") or {}; ea=d.get("e2e_ascii") or {}; cb=d.get("cpu_baseline") or {}
r=d["roofline"]
print("%-9s %8.3f ms/step %7.3f G/s | e2e %6.3f ascii %6.3f | cpu %7.2f M/s | dom %s alg-frac %.2f dram-frac %s | %s"%(sys.argv[1], d["ms_per_step"], d["value"]/1e9, e.get("value",0)/1e9, ea.get("value",0)/1e9, cb.get("value",0)/1e6, r["phase"], r["frac"], r.get("dram_frac_of_peak"), " ".join("%s=%.2f"%(k,v) for k,v in d["phase_ms_per_step"].items())))
PY
done
