#!/bin/bash
# reduction transports at N ranks: parity test, strong-scaling simulations with NCCL (high-priority group, the default)
# and with the peer-memory transport
N=$1
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -12 > gpurun_out/pytest_multi_n$N.txt
cat gpurun_out/pytest_multi_n$N.txt
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$RUN bench.py --gpus $N --only-strong --strong cfg3,cfg5 --strong-steps 3 > gpurun_out/strong_nccl_n$N.json 2> gpurun_out/strong_nccl_n$N.err
tail -5 gpurun_out/strong_nccl_n$N.err | cut -c 1-400
OPTK_REDUCE_TRANSPORT=peer $RUN bench.py --gpus $N --only-strong --strong cfg3,cfg5 --strong-steps 3 > gpurun_out/strong_peer_n$N.json 2> gpurun_out/strong_peer_n$N.err
tail -5 gpurun_out/strong_peer_n$N.err | cut -c 1-400
OPTK_REDUCE_TRANSPORT=peer python -m pytest tests/test_gpu_multi.py -q -x -k "2" 2>&1 | tail -3
python tools/gpu/summarize_strong.py gpurun_out/strong_nccl_n$N.json gpurun_out/strong_peer_n$N.json
