#!/bin/bash
TAG=${1:-run}
mkdir -p gpurun_out
timeout 600 python tools/timeline.py > gpurun_out/${TAG}_timeline.txt 2>&1
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
SCNB_PREFETCH=0 timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_bench_n1_nopf.json 2> gpurun_out/${TAG}_bench_n1_nopf.err
head -c 250 gpurun_out/${TAG}_bench_n1.json; echo; head -c 250 gpurun_out/${TAG}_bench_n1_nopf.json
