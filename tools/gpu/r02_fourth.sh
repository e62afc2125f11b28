#!/bin/bash
# coating tables: tests and timings
mkdir -p gpurun_out
python -m pytest tests/test_gpu_coatings.py tests/test_gpu_models.py tests/test_gpu_trace.py -m gpu -q 2>&1 | tail -25 > gpurun_out/pytest_coatings.txt
tail -6 gpurun_out/pytest_coatings.txt
python tools/measure_coatings.py > gpurun_out/coatings.json 2> gpurun_out/coatings.err; tail -3 gpurun_out/coatings.err; head -c 2500 gpurun_out/coatings.json
