#!/bin/bash
# full GPU tests + config 2 bench + config 5 bench (late / chain weight-gradient fork) + timeline
TAG=${1:-z}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_bench_c2.json 2> gpurun_out/${TAG}_bench_c2.err
timeout 600 python bench.py --config 5 --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
SCNB_WGRAD_FORK=chain timeout 600 python bench.py --config 5 --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_bench_c5_chain.json 2> gpurun_out/${TAG}_bench_c5_chain.err
timeout 300 python tools/timeline.py --config 5 > gpurun_out/${TAG}_timeline_c5.txt 2>&1
SCNB_WGRAD_FORK=chain timeout 300 python tools/timeline.py --config 5 > gpurun_out/${TAG}_timeline_c5_chain.txt 2>&1
tail -6 gpurun_out/${TAG}_tests.log; head -c 300 gpurun_out/${TAG}_bench_c2.json; echo;  head -c 300 gpurun_out/${TAG}_bench_c5.json; echo; head -c 300 gpurun_out/${TAG}_bench_c5_chain.json; tail -3 gpurun_out/${TAG}_bench_c5.err
