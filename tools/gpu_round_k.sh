#!/bin/bash
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_surface.py tests/test_engine_misc.py -m gpu -q -p no:cacheprovider -x > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
FEED=mapped timeout 300 python tools/timeline_e2e.py > gpurun_out/${TAG}_timeline_e2e_mapped.txt 2>&1
timeout 900 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-gpu-eager --no-predict > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench_n1.err
tail -4 gpurun_out/${TAG}_tests.log; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_n1.json")); print(d["value"], d["ms_per_step"], d["e2e"]["feeds"])
except Exception as e: print("bench parse failed", e)
PY
tail -5 gpurun_out/${TAG}_bench_n1.err
