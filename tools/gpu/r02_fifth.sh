#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_grid.py tests/test_gpu_detector.py tests/test_gpu_jit.py tests/test_gpu_coatings.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_fifth.txt
tail -30 gpurun_out/pytest_fifth.txt | cut -c 1-250
