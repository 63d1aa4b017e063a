#!/bin/bash
# Multi-GPU measurement recipe (run on the GPU box under gpurun --gpus G):
#   tools/run_multi_gpu.sh G TAG
# 1. the NCCL parity test (rows + pull halo, rows + all-gather, column slabs vs the single-GPU class)
# 2. the driver's own bench command: serving replicas + the sharded N=10M and N=1M lattices in one JSON line
G=${1:-2}; TAG=${2:-mg}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_sharded.py -x -q -k nccl 2>&1 | tail -3
timeout 1500 $TR --master-port 29512 bench.py --gpus $G --steps 5 --warmup 3 \
  > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -c 1500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
