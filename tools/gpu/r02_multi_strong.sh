#!/bin/bash
# strong-scaling simulations only, at N ranks: tools/gpu/r02_multi_strong.sh N
N=$1
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$RUN bench.py --gpus $N --only-strong --strong cfg3,cfg5 --strong-steps 3 > gpurun_out/strong_n$N.json 2> gpurun_out/strong_n$N.err
python tools/gpu/summarize_strong.py gpurun_out/strong_n$N.json; tail -2 gpurun_out/strong_n$N.err | cut -c 1-300
