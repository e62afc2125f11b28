#!/bin/bash
# Round 2, first GPU call: probe for the real reference, the GPU suite, the strong-scaling image simulations at N = 1,
# executed-instruction counts of the fused kernels.
mkdir -p gpurun_out
(
  python -c "import named_arrays"; python -c "import astropy"; python -c "import optika"
  timeout 20 python -m pip download --no-deps -d /tmp/pd named-arrays astropy 2>&1 | tail -3
  timeout 10 curl -sI https://pypi.org | head -1
  ls /opt/wheelhouse | grep -i -E "astropy|named"
  nproc; free -g | head -2; nvidia-smi -L; df -h /dev/shm | tail -1
) > gpurun_out/probe.txt 2>&1
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
tail -3 gpurun_out/pytest_gpu.txt
python bench.py --only-strong --strong cfg3,cfg5 > gpurun_out/strong_n1.json 2> gpurun_out/strong_n1.err
OPTK_TRACE_SPREAD=0 python bench.py --only-strong --strong cfg5 > gpurun_out/strong_n1_nospread.json 2> gpurun_out/strong_n1_nospread.err
M=smsp__inst_executed.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_red.sum,l1tex__t_requests_pipe_lsu_mem_global_op_red.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
for what in "cfg2 grid" "cfg2 image" "cfg3 grid" "cfg3 dense" "cfg1 grid"; do
  set -- $what
  ncu --metrics $M --clock-control none -k regex:optk_jit_kernel -s 2 -c 1 --csv --log-file gpurun_out/ncu_$1_$2.csv python tools/profile_config.py $1 $2 > /dev/null 2>&1
done
python tools/measure_grid.py > gpurun_out/grid.json 2> gpurun_out/grid.err
head -c 3000 gpurun_out/strong_n1.json
tail -5 gpurun_out/strong_n1.err
grep -h "inst_executed.sum\|time_duration" gpurun_out/ncu_*.csv | tail -12
