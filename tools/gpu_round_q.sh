#!/bin/bash
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench_default.err
tail -2 gpurun_out/${TAG}_smoke.log | cut -c1-200; head -c 400 gpurun_out/${TAG}_bench_default.json; tail -2 gpurun_out/${TAG}_bench_default.err
