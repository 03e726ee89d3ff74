#!/bin/bash
TAG=${1:-run}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-gpu-eager > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench_n1.err
timeout 600 python tools/timeline.py > gpurun_out/${TAG}_timeline.txt 2>&1
timeout 600 python bench.py --config 3 --steps 100 --warmup 10 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
tail -5 gpurun_out/${TAG}_tests.log; head -c 300 gpurun_out/${TAG}_bench_n1.json; tail -3 gpurun_out/${TAG}_bench_n1.err
