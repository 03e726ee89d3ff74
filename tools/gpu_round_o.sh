#!/bin/bash
TAG=${1:-run}
mkdir -p gpurun_out
for M in cache nocache; do
SCNB_PREFETCH_GATHER=$M timeout 300 python tools/timeline.py > gpurun_out/${TAG}_timeline_$M.txt 2>&1
for i in 1 2; do
SCNB_PREFETCH_GATHER=$M timeout 600 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_bench_$M.json 2> gpurun_out/${TAG}_bench_$M.err
echo "$M"; head -c 200 gpurun_out/${TAG}_bench_$M.json | cut -c 60-200
done
grep "k_gather_rows\|k_hs_fwd\|k_csr_dense" gpurun_out/${TAG}_timeline_$M.txt
done
