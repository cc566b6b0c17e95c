#!/bin/bash
# Developer helper (GPU box): for every BASELINE configuration one full bench line (profiles/r2_bench_<cfg>.json material) and
# one ncu launch list with DRAM bytes (input of tools/roofline_traffic.py).   tools/profile_all.sh [cfg ...]
mkdir -p gpurun_out
CFGS=${@:-cfg1 cfg0 cfg2 cfg3_r2 cfg3_r16 cfg4_b5 cfg4_b6 cfg5}
for c in $CFGS; do
  STEPS="--steps 5 --warmup 3"; [ $c == cfg1 ] && STEPS="--steps 10 --warmup 3"
  timeout 900 python bench.py --config $c $STEPS > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err || { echo "$c bench FAILED"; tail -3 gpurun_out/r2_bench_$c.err; continue; }
  python - $c <<PY
import json,sys
d=json.load(open("gpurun_out/r2_bench_%s.json"%sys.argv[1]))
e=d.get("e2e") or {}; ea=d.get("e2e_ascii") or {}; cb=d.get("cpu_baseline") or {}
print("%-9s %8.3f ms/step %7.3f G/s | e2e %6.3f ascii %6.3f | cpu %7.2f M/s x%d | %s"%(sys.argv[1], d["ms_per_step"], d["value"]/1e9, e.get("value",0)/1e9, ea.get("value",0)/1e9, cb.get("value",0)/1e6, cb.get("cores",0), " ".join("%s=%.2f"%(k,v) for k,v in d["phase_ms_per_step"].items())))
PY
  timeout 600 python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain_$c.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_${c}_launches.csv \
      python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$c.log 2>&1 || echo "$c ncu FAILED"
done
