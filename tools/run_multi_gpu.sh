#!/bin/bash
# Multi-GPU measurement recipe (run on the GPU box under gpurun --gpus G):
#   tools/run_multi_gpu.sh G TAG [N D K CHAIN] [SERVING=1]
G=${1:-2}; TAG=${2:-mg}; N=${3:-1000000}; D=${4:-768}; K=${5:-16}; CH=${6:-0}; SERVING=${7:-1}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
mkdir -p gpurun_out
for MODE in rows columns; do
  timeout 300 $TR --master-port 29511 tools/sharded_check.py --N 20000 --D 384 --k 10 --mode $MODE \
    > gpurun_out/${TAG}_check_${MODE}.json 2> gpurun_out/${TAG}_check_${MODE}.err
  tail -1 gpurun_out/${TAG}_check_${MODE}.json | cut -c1-600
done
# rows with the fused P2P halo + columns (one build), then rows with the NCCL all-gather halo
timeout 1500 $TR --master-port 29512 bench.py --gpus $G --workload large --N $N --D $D --k $K --chain-len $CH \
  --partition both --steps 3 --warmup 1 > gpurun_out/${TAG}_large.json 2> gpurun_out/${TAG}_large.err
tail -1 gpurun_out/${TAG}_large.json | cut -c1-3500; tail -3 gpurun_out/${TAG}_large.err
if [ "$SERVING" = "1" ]; then
  timeout 600 $TR --master-port 29513 bench.py --gpus $G --steps 3 --warmup 3 --no-large \
    > gpurun_out/${TAG}_serving.json 2> gpurun_out/${TAG}_serving.err
  tail -1 gpurun_out/${TAG}_serving.json | cut -c1-1500; tail -3 gpurun_out/${TAG}_serving.err
fi
