#!/bin/bash
# fused heads (head_ops.cu): tests, bench with the heads fused / split, timeline
TAG=${1:-run}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -30 gpurun_out/${TAG}_tests.log
for S in fused split; do
SCNB_HEADS=$S timeout 900 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-gpu-eager --no-predict > gpurun_out/${TAG}_bench_$S.json 2> gpurun_out/${TAG}_bench_$S.err
tail -3 gpurun_out/${TAG}_bench_$S.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_$S.json")); print("heads=$S", d["value"], d["ms_per_step"], d["e2e"]["feeds"], d["losses_last_step"])
PY
done
timeout 300 python tools/timeline.py > gpurun_out/${TAG}_timeline.txt 2>&1
grep -v "^/opt\|_warn" gpurun_out/${TAG}_timeline.txt | head -50
