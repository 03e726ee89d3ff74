#!/bin/bash
# config-5 timeline, bn tests, split-points-late experiment on config 2
TAG=${1:-y}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_kernels.py tests/test_full_shape_parity.py -m gpu -q -p no:cacheprovider -k "bn_act or config5" > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --config 5 --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
timeout 300 python tools/timeline.py --config 5 > gpurun_out/${TAG}_timeline_c5.txt 2>&1
for v in 0 1 2 0 1 2; do
SCNB_SPLITS_LATE=$v timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e 2>/dev/null | head -c 330 >> gpurun_out/${TAG}_late.txt; echo " late=$v" >> gpurun_out/${TAG}_late.txt
done
SCNB_SPLITS_LATE=2 timeout 300 python tools/timeline.py > gpurun_out/${TAG}_timeline_late2.txt 2>&1
tail -4 gpurun_out/${TAG}_tests.log; head -c 300 gpurun_out/${TAG}_bench_c5.json; echo; grep -o '"ms_per_step": [0-9.]*\| late=[012]' gpurun_out/${TAG}_late.txt
