#!/bin/bash
# GPU-box visit at the head of the tree: tests, smoke, bench, timeline, ncu launch list.
TAG=${1:-w}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 ) > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
( time timeout 900 python bench.py --steps 100 --warmup 10 ) > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench_n1.err
timeout 300 python tools/timeline.py > gpurun_out/${TAG}_timeline.txt 2>&1
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_ncu_launch.log 2>&1
tail -5 gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_smoke.log; head -c 600 gpurun_out/${TAG}_bench_n1.json; tail -3 gpurun_out/${TAG}_bench_n1.err
