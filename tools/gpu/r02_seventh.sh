#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
tail -6 gpurun_out/pytest_gpu.txt | cut -c 1-250
python bench.py --only-strong --strong cfg3,cfg5 > gpurun_out/strong_relative.json 2>/dev/null
OPTK_TRACE_RELATIVE=0 OPTK_JIT_CACHE= python bench.py --only-strong --strong cfg3,cfg5 > gpurun_out/strong_norelative.json 2>/dev/null
python tools/gpu/summarize_strong.py gpurun_out/strong_relative.json gpurun_out/strong_norelative.json
python tools/measure_grid.py > gpurun_out/grid3.json 2>/dev/null
python -c "
import json
for r in json.load(open('gpurun_out/grid3.json'))['results']:
    print(r['config'][:58], r['jitter'], round(r['ms_grid_fused_trace_bin'],2), '%.3g' % r['intercepts_per_s_grid_fused'], round(r['fp64_tflops_algorithmic'],1))
"
