#!/bin/bash
# Round 2, final 1-GPU call: whole GPU suite, bench line, reference arm, launch list, smoke, and the measurement scripts
# behind DESIGN.md's tables (configs, grids, strong-scaling objects at N = 1, the user-level image() call, counters).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
tail -3 gpurun_out/pytest_gpu.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --num-wavelength 8 --no-cpu-baseline --strong none > gpurun_out/bench_under_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -3 gpurun_out/smoke.txt
head -c 1800 gpurun_out/bench_n1.json; echo; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_reference.json | head -c 600; echo
python tools/measure_configs.py > gpurun_out/r02c_configs.json 2> gpurun_out/r02c_configs.err
python tools/measure_grid.py > gpurun_out/grid4.json 2>/dev/null
python tools/measure_image_call.py > gpurun_out/image_call.json 2> gpurun_out/image_call.err
python -c "
import json
for r in json.load(open('gpurun_out/grid4.json'))['results']:
    print(r['config'][:58], r['jitter'], round(r['ms_grid_fused_trace_bin'],2), '%.3g' % r['intercepts_per_s_grid_fused'], round(r['fp64_tflops_algorithmic'],1))
d=json.load(open('gpurun_out/image_call.json'))
for k,v in d.items(): print(k, round(v['seconds'],2))
"
M=smsp__inst_executed.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_red.sum,l1tex__t_requests_pipe_lsu_mem_global_op_red.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
for what in "cfg3 dense" "cfg3 grid" "cfg2 grid" "cfg2 image" "cfg2 dense" "cfg1 grid" "cfg1 dense" "cfg5 grid"; do
  set -- $what
  ncu --metrics $M --clock-control none -k regex:optk_jit_kernel -s 2 -c 1 --csv --log-file gpurun_out/ncu_c_$1_$2.csv python tools/profile_config.py $1 $2 > /dev/null 2>&1
done
for what in "cfg5 grid" "cfg2 grid" "cfg2 image" "cfg3 grid"; do
  set -- $what
  ncu --set full --clock-control none --import-source on -k regex:optk_jit_kernel -s 2 -c 1 -f -o gpurun_out/r02e_prof_$1_$2 \
    python tools/profile_config.py $1 $2 > gpurun_out/r02e_prof_$1_$2.log 2>&1
done
ls gpurun_out/r02e_*.ncu-rep
