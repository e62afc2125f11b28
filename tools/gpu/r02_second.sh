#!/bin/bash
# Round 2, second GPU call (1 GPU): the whole GPU suite, the bench line, the reference arm, the launch list.
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.txt
tail -3 gpurun_out/pytest_gpu.txt
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --num-wavelength 8 --no-cpu-baseline --strong none > gpurun_out/bench_under_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -3 gpurun_out/smoke.txt
head -c 2500 gpurun_out/bench_n1.json; echo; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_reference.json | head -c 800
