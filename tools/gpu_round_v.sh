#!/bin/bash
TAG=${1:-run}
mkdir -p gpurun_out
timeout 300 python tools/head_phases.py > gpurun_out/${TAG}_head_phases.txt 2>&1
grep "launch 3" gpurun_out/${TAG}_head_phases.txt
timeout 300 python tools/hs_phases.py > gpurun_out/${TAG}_hs_phases.txt 2>&1
tail -1 gpurun_out/${TAG}_hs_phases.txt
timeout 1500 python -m pytest tests/test_engine_misc.py tests/test_full_shape_parity.py tests/test_cuda_parity.py tests/test_cuda_kernels.py -m gpu -q -p no:cacheprovider -x > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
for i in 1 2; do
timeout 900 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_bench$i.json 2> gpurun_out/${TAG}_bench$i.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench$i.json")); print(d["value"], d["ms_per_step"], d["losses_last_step"])
PY
done
timeout 300 python tools/timeline.py > gpurun_out/${TAG}_timeline.txt 2>&1
grep -v "^/opt\|_warn" gpurun_out/${TAG}_timeline.txt | sed -n 1,40p
