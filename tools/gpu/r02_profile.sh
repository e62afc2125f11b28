#!/bin/bash
# ncu --set full captures of the fused kernels (one launch each), round 2
mkdir -p gpurun_out
for what in "cfg2 grid" "cfg2 image" "cfg5 grid" "cfg3 grid"; do
  set -- $what
  ncu --set full --clock-control none --import-source on -k regex:optk_jit_kernel -s 2 -c 1 -f -o gpurun_out/r02_prof_$1_$2 \
    python tools/profile_config.py $1 $2 > gpurun_out/r02_prof_$1_$2.log 2>&1
  ls -la gpurun_out/r02_prof_$1_$2.ncu-rep
done
