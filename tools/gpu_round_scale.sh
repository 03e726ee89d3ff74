#!/bin/bash
# Scaling record on N GPUs of one box: bench.py under torchrun (config 3), timeline of rank 0.
TAG=${1:-run}; N=${2:-8}
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench_n${N}.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/timeline.py --config 3 > gpurun_out/${TAG}_timeline_dp${N}.txt 2> gpurun_out/${TAG}_timeline_dp${N}.err
head -c 300 gpurun_out/${TAG}_bench_n${N}.json; tail -3 gpurun_out/${TAG}_bench_n${N}.err
