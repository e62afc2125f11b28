#!/bin/bash
# after the relative frames: ncu --set full of the cfg 5 / cfg 3 fused kernels, instruction counters of all fused kernels
mkdir -p gpurun_out
for what in "cfg5 grid" "cfg3 grid"; do
  set -- $what
  ncu --set full --clock-control none --import-source on -k regex:optk_jit_kernel -s 2 -c 1 -f -o gpurun_out/r02b_prof_$1_$2 \
    python tools/profile_config.py $1 $2 > gpurun_out/r02b_prof_$1_$2.log 2>&1
done
M=smsp__inst_executed.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_red.sum,l1tex__t_requests_pipe_lsu_mem_global_op_red.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
for what in "cfg2 grid" "cfg2 image" "cfg3 grid" "cfg3 dense" "cfg1 grid" "cfg5 grid" "cfg1 dense" "cfg2 dense"; do
  set -- $what
  ncu --metrics $M --clock-control none -k regex:optk_jit_kernel -s 2 -c 1 --csv --log-file gpurun_out/ncu_$1_$2.csv python tools/profile_config.py $1 $2 > /dev/null 2>&1
done
ls gpurun_out/ncu_*.csv gpurun_out/r02b_*.ncu-rep
