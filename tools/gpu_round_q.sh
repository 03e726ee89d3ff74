#!/bin/bash
TAG=${1:-q}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --config 5 --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
tail -3 gpurun_out/${TAG}_tests.log; head -c 300 gpurun_out/${TAG}_bench_c5.json
