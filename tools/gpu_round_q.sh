#!/bin/bash
TAG=${1:-q}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_engine_misc.py -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
SCNB_PREFETCH_AT=heads timeout 300 python -m pytest tests/test_engine_misc.py -m gpu -q -p no:cacheprovider -k prefetched >> gpurun_out/${TAG}_tests.log 2>&1
for v in start heads start heads; do
SCNB_PREFETCH_AT=$v timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e 2>/dev/null | head -c 330 >> gpurun_out/${TAG}_pf.txt; echo " at=$v" >> gpurun_out/${TAG}_pf.txt
done
SCNB_PREFETCH_AT=heads timeout 300 python tools/timeline.py > gpurun_out/${TAG}_timeline_pfheads.txt 2>&1
grep passed gpurun_out/${TAG}_tests.log; grep -o '"ms_per_step": [0-9.]*\| at=[a-z]*' gpurun_out/${TAG}_pf.txt
