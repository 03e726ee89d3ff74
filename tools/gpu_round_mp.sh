#!/bin/bash
# Multi-GPU visit: bench.py under torchrun on N GPUs (config 3), plus the reference arm.
TAG=${1:-run}; N=${2:-2}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench_n${N}.err
head -c 400 gpurun_out/${TAG}_bench_n${N}.json; tail -5 gpurun_out/${TAG}_bench_n${N}.err
