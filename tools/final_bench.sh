#!/bin/bash
# Developer helper (GPU box): the full bench line (device arm, end-to-end arm, CPU baseline) of every listed configuration.
mkdir -p gpurun_out
for c in "$@"; do
  STEPS="--steps 5 --warmup 3"; [ $c == cfg1 ] && STEPS="--steps 10 --warmup 3"
  timeout 900 python bench.py --config $c $STEPS > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err || { echo "$c bench FAILED"; tail -3 gpurun_out/r2_bench_$c.err; continue; }
  python - $c <<PY
import json,sys
d=json.load(open("gpurun_out/r2_bench_%s.json"%sys.argv[1]))
e=d.get("e2e") or {}; ea=d.get("e2e_ascii") or {}; cb=d.get("cpu_baseline") or {}
print("%-9s %8.3f ms/step %7.3f G/s | e2e %6.3f ascii %6.3f | cpu %7.2f M/s x%d | %s"%(sys.argv[1], d["ms_per_step"], d["value"]/1e9, e.get("value",0)/1e9, ea.get("value",0)/1e9, cb.get("value",0)/1e6, cb.get("cores",0), " ".join("%s=%.2f"%(k,v) for k,v in d["phase_ms_per_step"].items())))
PY
done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_cfg1_reference_arm.json 2> gpurun_out/r2_bench_ref.err; tail -c 400 gpurun_out/r2_bench_cfg1_reference_arm.json
