#!/bin/bash
# Planes form of the tcgen05 hidden GEMMs: kernel tests, config-5 parity, GEMM timings, config-5 bench with and without it.
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_kernels.py tests/test_full_shape_parity.py -m gpu -q -p no:cacheprovider -k "tcgen05 or config5" > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
timeout 300 python tools/bench_gemm.py > gpurun_out/${TAG}_gemm.txt 2>&1
timeout 300 python tools/bench_gemm.py 16384 1024 256 >> gpurun_out/${TAG}_gemm.txt 2>&1
timeout 600 python bench.py --config 5 --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-predict > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
SCNB_GEMM_TC_PLANES=0 timeout 600 python bench.py --config 5 --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_bench_c5_noplanes.json 2> gpurun_out/${TAG}_bench_c5_noplanes.err
tail -6 gpurun_out/${TAG}_tests.log; cat gpurun_out/${TAG}_gemm.txt; head -c 400 gpurun_out/${TAG}_bench_c5.json; echo; head -c 400 gpurun_out/${TAG}_bench_c5_noplanes.json; tail -3 gpurun_out/${TAG}_bench_c5.err
