#!/bin/bash
TAG=${1:-run}
mkdir -p gpurun_out
for AT in start fl; do
SCNB_PREFETCH_AT=$AT timeout 600 python tools/timeline.py > gpurun_out/${TAG}_timeline_$AT.txt 2>&1
SCNB_PREFETCH_AT=$AT timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_bench_$AT.json 2> gpurun_out/${TAG}_bench_$AT.err
head -c 250 gpurun_out/${TAG}_bench_$AT.json; echo
done
