#!/bin/bash
TAG=${1:-run}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multitask.py tests/test_reference_suite.py tests/test_engine_misc.py -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench_n1.err
timeout 600 python tools/timeline.py > gpurun_out/${TAG}_timeline.txt 2>&1
tail -5 gpurun_out/${TAG}_tests.log; head -c 400 gpurun_out/${TAG}_bench_n1.json; tail -3 gpurun_out/${TAG}_bench_n1.err
