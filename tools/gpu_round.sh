#!/bin/bash
# One GPU-box visit: tests, smoke, bench (ours + reference arm).  Everything lands in gpurun_out/ with a tag.
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
tail -5 gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_smoke.log; head -c 600 gpurun_out/${TAG}_bench_n1.json; tail -3 gpurun_out/${TAG}_bench_n1.err
