#!/bin/bash
TAG=${1:-run}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -x > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
timeout 300 python tools/timeline.py > gpurun_out/${TAG}_timeline.txt 2>&1
FEED=mapped timeout 300 python tools/timeline_e2e.py > gpurun_out/${TAG}_timeline_e2e_mapped.txt 2>&1
for Z in 1 0; do
SCNB_ZERO_AT_END=$Z timeout 900 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-gpu-eager --no-predict > gpurun_out/${TAG}_bench_z$Z.json 2> gpurun_out/${TAG}_bench_z$Z.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_z$Z.json")); print("zero_at_end=$Z", d["value"], d["ms_per_step"], d["e2e"]["feeds"])
PY
done
