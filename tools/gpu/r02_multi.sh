#!/bin/bash
# Round 2, multi-GPU call: tools/gpu/r02_multi.sh N -- parity of the sharded + reduced image, the strong-scaling
# simulations and the bench line at N ranks.
N=$1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_gpus_n$N.txt; nvidia-smi topo -m >> gpurun_out/multi_gpus_n$N.txt 2>&1
python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -8 > gpurun_out/pytest_multi_n$N.txt
cat gpurun_out/pytest_multi_n$N.txt
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$RUN bench.py --gpus $N --only-strong --strong cfg3,cfg5 --strong-steps 3 > gpurun_out/strong_n$N.json 2> gpurun_out/strong_n$N.err
head -c 3000 gpurun_out/strong_n$N.json; tail -3 gpurun_out/strong_n$N.err
if [ "$2" = "ab" ]; then
  OPTK_REDUCE_TRANSPORT=nccl $RUN bench.py --gpus $N --only-strong --strong cfg5 --strong-steps 3 > gpurun_out/strong_n${N}_nccl.json 2> gpurun_out/strong_n${N}_nccl.err
  OPTK_REDUCE_TRANSPORT=nccl TORCH_NCCL_HIGH_PRIORITY=1 $RUN bench.py --gpus $N --only-strong --strong cfg5 --strong-steps 3 > gpurun_out/strong_n${N}_nccl_hp.json 2> gpurun_out/strong_n${N}_nccl_hp.err
  exit 0
fi
$RUN bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
head -c 1500 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
