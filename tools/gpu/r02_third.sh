#!/bin/bash
# Round 2, third 1-GPU call: the whole GPU suite (coating tables, models), coated-system timings, N = 1 strong lines.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
tail -4 gpurun_out/pytest_gpu.txt
python tools/measure_coatings.py > gpurun_out/coatings.json 2> gpurun_out/coatings.err; tail -3 gpurun_out/coatings.err; head -c 1800 gpurun_out/coatings.json
python bench.py --only-strong --strong cfg3,cfg5 > gpurun_out/strong_n1b.json 2> gpurun_out/strong_n1b.err; tail -c 900 gpurun_out/strong_n1b.json
OPTK_BIN_DIRECT=0 python tools/measure_grid.py > gpurun_out/grid_nodirect.json 2> gpurun_out/grid_nodirect.err
python tools/measure_grid.py > gpurun_out/grid2.json 2> gpurun_out/grid2.err
