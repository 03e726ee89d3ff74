#!/bin/bash
TAG=${1:-run}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
timeout 300 python tools/hs_phases.py > gpurun_out/${TAG}_hs_phases.txt 2>&1
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-gpu-eager --no-predict > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench_n1.err
timeout 600 python tools/timeline.py > gpurun_out/${TAG}_timeline.txt 2>&1
tail -4 gpurun_out/${TAG}_tests.log; tail -4 gpurun_out/${TAG}_hs_phases.txt; head -c 300 gpurun_out/${TAG}_bench_n1.json
