#!/bin/bash
# after the cfg 3 diet: strong-scaling objects at N = 1, on-device grid timings, ncu --set full of the fused kernels
mkdir -p gpurun_out
python bench.py --only-strong --strong cfg3,cfg5 > gpurun_out/strong_n1_d.json 2>/dev/null
python tools/gpu/summarize_strong.py gpurun_out/strong_n1_d.json
python tools/measure_grid.py > gpurun_out/grid4.json 2>/dev/null
python -c "
import json
for r in json.load(open('gpurun_out/grid4.json'))['results']:
    print(r['config'][:58], r['jitter'], round(r['ms_grid_fused_trace_bin'],2), '%.3g' % r['intercepts_per_s_grid_fused'], round(r['fp64_tflops_algorithmic'],1))
"
for what in "cfg5 grid" "cfg2 grid" "cfg3 dense"; do
  set -- $what
  ncu --set full --clock-control none --import-source on -k regex:optk_jit_kernel -s 2 -c 1 -f -o gpurun_out/r02d_prof_$1_$2 \
    python tools/profile_config.py $1 $2 > gpurun_out/r02d_prof_$1_$2.log 2>&1
done
ls -la gpurun_out/r02d_*.ncu-rep
