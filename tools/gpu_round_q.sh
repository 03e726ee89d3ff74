#!/bin/bash
TAG=${1:-q}
mkdir -p gpurun_out
for v in side chain side chain; do
SCNB_WGRAD_LAST=$v timeout 600 python bench.py --config 5 --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e 2>/dev/null | head -c 330 >> gpurun_out/${TAG}_last.txt; echo " last=$v" >> gpurun_out/${TAG}_last.txt
done
grep -o '"ms_per_step": [0-9.]*\| last=[a-z]*' gpurun_out/${TAG}_last.txt
