#!/bin/bash
TAG=${1:-run}
mkdir -p gpurun_out
timeout 300 python tools/head_phases.py > gpurun_out/${TAG}_head_phases.txt 2>&1
cat gpurun_out/${TAG}_head_phases.txt
timeout 1500 python -m pytest tests/test_engine_misc.py tests/test_full_shape_parity.py tests/test_cuda_parity.py tests/test_parallel.py -m gpu -q -p no:cacheprovider -x > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -15 gpurun_out/${TAG}_tests.log
for S in fused split; do
SCNB_BWD_IN=$S timeout 900 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_bench_$S.json 2> gpurun_out/${TAG}_bench_$S.err
tail -3 gpurun_out/${TAG}_bench_$S.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_$S.json")); print("bwd_in=$S", d["value"], d["ms_per_step"], d["losses_last_step"])
PY
done
timeout 300 python tools/timeline.py > gpurun_out/${TAG}_timeline.txt 2>&1
grep -v "^/opt\|_warn" gpurun_out/${TAG}_timeline.txt | sed -n 14,40p
