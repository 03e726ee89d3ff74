#!/bin/bash
TAG=${1:-run}
mkdir -p gpurun_out
SCNB_PDL=1 timeout 900 python -m pytest tests/test_full_shape_parity.py tests/test_cuda_parity.py -m gpu -q -p no:cacheprovider -x > gpurun_out/${TAG}_tests_pdl1.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests_pdl1.log
for pdl in 0 1; do
SCNB_PDL=$pdl timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_bench_pdl${pdl}.json 2> gpurun_out/${TAG}_bench_pdl${pdl}.err
SCNB_PDL=$pdl timeout 600 python tools/timeline.py > gpurun_out/${TAG}_timeline_pdl${pdl}.txt 2>&1
done
tail -3 gpurun_out/${TAG}_tests_pdl1.log; head -c 250 gpurun_out/${TAG}_bench_pdl0.json; echo; head -c 250 gpurun_out/${TAG}_bench_pdl1.json
