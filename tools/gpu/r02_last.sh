#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_last.txt
tail -3 gpurun_out/pytest_gpu_last.txt
M=smsp__inst_executed.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
for what in "cfg2 image" "cfg2 grid" "cfg5 grid" "cfg3 dense" "cfg1 grid" "cfg2 dense"; do
  set -- $what
  timeout 60 ncu --metrics $M --clock-control none -k regex:optk_jit_kernel -s 2 -c 1 --csv --log-file gpurun_out/ncu_l_$1_$2.csv python tools/profile_config.py $1 $2 > /dev/null 2>&1
  echo $what $(grep inst_executed.sum gpurun_out/ncu_l_$1_$2.csv | tail -1 | awk -F'","' '{print $NF}') $(grep gpu__time gpurun_out/ncu_l_$1_$2.csv | tail -1 | awk -F'","' '{print $NF}')
done
