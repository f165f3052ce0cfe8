#!/bin/bash
# multi-GPU call: N GPUs of one box.  2-GPU parity tests (multi-process and single-process), then the bench line at N.
cd "$(dirname "$0")/.."
N=${1:-2}
T=${2:-r2_multi}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${T}_n${N}_gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/${T}_n${N}_gpus.txt 2>&1
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_frame_multi.py -m gpu -q ) > gpurun_out/${T}_n${N}_tests.log 2>&1
tail -5 gpurun_out/${T}_n${N}_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${N} --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus ${N} --steps 20 --warmup 3 \
    > gpurun_out/${T}_n${N}_bench.json 2> gpurun_out/${T}_n${N}_bench.err
tail -c 1500 gpurun_out/${T}_n${N}_bench.err
head -c 3000 gpurun_out/${T}_n${N}_bench.json
