#!/bin/bash
# Developer helper (GPU box): device-resident bench of the default library and of variants.
#   tools/ab.sh [name:"ENV=1 OTHER=2" ...]   a variant library is selected with SVFM_LIB_PATH=tools/dev_libs/libsvfm_X.so
#   -> gpurun_out/ab_<name>.json, one summary line each.  BENCH_ARGS overrides the bench arguments.
mkdir -p gpurun_out
ARGS=${BENCH_ARGS:---steps 5 --warmup 3 --no-cpu-baseline --no-e2e}
run() {
  local name=$1 envs=$2
  env $envs timeout 300 python bench.py $ARGS > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err || { echo "$name FAILED"; tail -3 gpurun_out/ab_$name.err; return; }
  python - "$name" <<PY
import json,sys
d=json.load(open("gpurun_out/ab_%s.json"%sys.argv[1]))
ph=d["phase_ms_per_step"]
e=d.get("e2e") or {}
ea=d.get("e2e_ascii") or {}
print("%-12s %.3f ms/step  %.2f G/s  count-only %.2f G/s  e2e %.2f (ascii %.2f) | "%(sys.argv[1], d["ms_per_step"], d["value"]/1e9, (d.get("count_only_patterns_per_s") or 0)/1e9, e.get("value",0)/1e9, ea.get("value",0)/1e9) + "  ".join("%s=%.2f"%(k.split("(")[0],v) for k,v in ph.items()))
PY
}
if [ $# -eq 0 ]; then run default ""; fi
for v in "$@"; do run "${v%%:*}" "${v#*:}"; done
