#!/bin/bash
TAG=${1:-run}
mkdir -p gpurun_out
timeout 300 python tools/hs_phases.py > gpurun_out/${TAG}_hs_phases.txt 2>&1
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-gpu-eager --no-predict > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench_n1.err
cat gpurun_out/${TAG}_hs_phases.txt | tail -6; head -c 300 gpurun_out/${TAG}_bench_n1.json
