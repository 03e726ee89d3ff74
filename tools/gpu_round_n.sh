#!/bin/bash
TAG=${1:-run}
mkdir -p gpurun_out
SCNB_FWD_ROWS=128 timeout 600 python -m pytest tests/test_cuda_kernels.py tests/test_full_shape_parity.py -m gpu -q -p no:cacheprovider -x -k "tc or config2 or first" > gpurun_out/${TAG}_tests128.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests128.log; tail -3 gpurun_out/${TAG}_tests128.log
for R in 256 128; do
SCNB_FWD_ROWS=$R timeout 300 python tools/timeline.py > gpurun_out/${TAG}_timeline_rows$R.txt 2>&1
SCNB_FWD_ROWS=$R timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_bench_rows$R.json 2> gpurun_out/${TAG}_bench_rows$R.err
echo "rows $R"; head -c 200 gpurun_out/${TAG}_bench_rows$R.json | cut -c 60-200; grep "k_csr_dense_fwd_tc" gpurun_out/${TAG}_timeline_rows$R.txt
done
