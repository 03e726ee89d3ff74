#!/bin/bash
# Multi-GPU visit: timeline of the data-parallel step (rank 0) + bench.py under torchrun on N GPUs.
TAG=${1:-run}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/timeline.py --config 3 > gpurun_out/${TAG}_timeline_dp${N}.txt 2> gpurun_out/${TAG}_timeline_dp${N}.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 100 --warmup 10 --no-predict > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench_n${N}.err
head -c 400 gpurun_out/${TAG}_bench_n${N}.json; tail -5 gpurun_out/${TAG}_bench_n${N}.err
