#!/bin/bash
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_kernels.py tests/test_full_shape_parity.py -m gpu -q -p no:cacheprovider -k "bn_act or config5 or planes or tcgen05" > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --config 5 --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
timeout 300 python tools/timeline.py --config 5 > gpurun_out/${TAG}_timeline_c5.txt 2>&1
tail -4 gpurun_out/${TAG}_tests.log; head -c 300 gpurun_out/${TAG}_bench_c5.json
