#!/bin/bash
# cfg 3 diet (third-order fp64 helpers, convex polygon half-planes): new tests, the kernels' parity tests, instruction counters
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r02c_pytest_focus.txt 2>&1
tail -15 gpurun_out/r02c_pytest_focus.txt
M=smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
for what in "cfg3 dense" "cfg3 grid" "cfg2 grid" "cfg2 image" "cfg2 dense" "cfg1 grid" "cfg5 grid"; do
  set -- $what
  ncu --metrics $M --clock-control none -k regex:optk_jit_kernel -s 2 -c 1 --csv --log-file gpurun_out/ncu_c_$1_$2.csv python tools/profile_config.py $1 $2 > /dev/null 2>&1
  tail -4 gpurun_out/ncu_c_$1_$2.csv | cut -d, -f5,12-
done
python tools/measure_configs.py > gpurun_out/r02c_configs.json 2> gpurun_out/r02c_configs.err; tail -3 gpurun_out/r02c_configs.err
