for B in 4166400 8333568 16666880 33333760; do
  BENCH_ARGS="--batch $B --steps 10 --warmup 3 --no-cpu-baseline --no-e2e" tools/ab.sh "sw_$B:SVFM_SWEEP_MIN=0" "ge_$B:SVFM_SWEEP_MIN=18446744073709551615"
done
BENCH_ARGS="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e" tools/ab.sh "ge_1e8:SVFM_SWEEP_MIN=18446744073709551615"
