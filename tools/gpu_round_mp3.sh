#!/bin/bash
# Multi-GPU visit: data-parallel kernel tests, then timeline (rank 0) + bench.py under torchrun on N GPUs.
TAG=${1:-run}; N=${2:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_kernels.py tests/test_parallel.py -m gpu -q -p no:cacheprovider -k "dp or parallel or peer or rank" > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/timeline.py --config 3 > gpurun_out/${TAG}_timeline_dp${N}.txt 2> gpurun_out/${TAG}_timeline_dp${N}.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 100 --warmup 10 --no-predict > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench_n${N}.err
tail -4 gpurun_out/${TAG}_tests.log; head -c 300 gpurun_out/${TAG}_bench_n${N}.json; tail -5 gpurun_out/${TAG}_bench_n${N}.err
