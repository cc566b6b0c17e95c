#!/bin/bash
# Developer helper (GPU box): end-to-end arm of bench.py under a few chunk / worker / sweep-threshold settings.
export BENCH_ARGS="--steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 5"
M=25165824
for c in 4194304 6291456 8388608; do for w in 3 4 6; do
tools/ab.sh "e_c${c}_w$w:SVFM_SWEEP_MIN=$M SVFM_CHUNK=$c SVFM_WORKERS=$w" | sed 's/| presort.*//'
done; done
