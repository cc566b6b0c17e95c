#!/bin/bash
# Developer helper (GPU box): device-resident bench of several configurations under two environment settings.
#   tools/cfg_ab.sh "ENV=1" cfg3_r2 cfg5 ...
E=$1; shift
for c in "$@"; do
  BENCH_ARGS="--config $c --steps 3 --warmup 2 --no-cpu-baseline --no-e2e" tools/ab.sh "${c}_def:" "${c}_alt:$E" | sed 's/e2e 0.00 (ascii 0.00) //'
done
