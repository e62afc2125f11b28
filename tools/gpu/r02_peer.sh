#!/bin/bash
# peer-memory reduction at N ranks: parity test, strong-scaling simulations with the peer and the NCCL transport
N=$1
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -12 > gpurun_out/pytest_multi_peer_n$N.txt
cat gpurun_out/pytest_multi_peer_n$N.txt
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$RUN bench.py --gpus $N --only-strong --strong cfg3,cfg5 --strong-steps 3 > gpurun_out/strong_peer_n$N.json 2> gpurun_out/strong_peer_n$N.err
tail -c 2500 gpurun_out/strong_peer_n$N.json; tail -5 gpurun_out/strong_peer_n$N.err
OPTK_REDUCE_TRANSPORT=nccl TORCH_NCCL_HIGH_PRIORITY=1 $RUN bench.py --gpus $N --only-strong --strong cfg5 --strong-steps 3 > gpurun_out/strong_nccl_hp_n$N.json 2> gpurun_out/strong_nccl_hp_n$N.err
tail -c 1200 gpurun_out/strong_nccl_hp_n$N.json
