#!/bin/bash
# Developer helper (GPU box): ncu launch lists (time + DRAM bytes per launch) for the listed configurations, one ncu --set full
# capture of the cfg1 sweep pipeline, and the device-assertion build over tests/sanitize_case.py.
mkdir -p gpurun_out
for c in "$@"; do
  timeout 600 python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/plain_$c.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_${c}_launches.csv \
      python bench.py --config $c --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_$c.log 2>&1 || echo "$c ncu FAILED"
  echo "$c: $(grep -c svfm gpurun_out/r2_${c}_launches.csv) svfm rows"
done
ncu --set full --clock-control none --import-source on -k regex:"locate_direct_kernel|sweep_round_kernel|pack_sweep_kernel|sb_place_kernel" -s 4 -c 5 \
    -o gpurun_out/r2c_bench_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_r2c.log 2>&1 || echo "ncu full FAILED"
SVFM_LIB_PATH=tools/dev_libs/libsvfm_checks.so python tests/sanitize_case.py > gpurun_out/sanitize_checks.log 2>&1; echo "checks build rc=$?"; tail -2 gpurun_out/sanitize_checks.log
