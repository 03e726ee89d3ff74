#!/bin/bash
# forward split table: kernel tests, full-shape parity, phase stamps of the first layer, bench with / without the table
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_kernels.py tests/test_full_shape_parity.py tests/test_engine_misc.py -m gpu -q -p no:cacheprovider -x > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
for S in 1 0; do
SCNB_FWD_SPLITS=$S SCNB_FWDTC_DBG=1 timeout 300 python tools/timeline.py --steps 3 2>&1 | grep fwdtc | tail -4 > gpurun_out/${TAG}_fwd_stamps_s$S.txt
cat gpurun_out/${TAG}_fwd_stamps_s$S.txt
SCNB_FWD_SPLITS=$S timeout 900 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-gpu-eager --no-predict > gpurun_out/${TAG}_bench_s$S.json 2> gpurun_out/${TAG}_bench_s$S.err
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_s$S.json")); print("fwd_splits=$S", d["value"], d["ms_per_step"], d["e2e"]["feeds"])
PY
done
timeout 300 python tools/timeline.py > gpurun_out/${TAG}_timeline.txt 2>&1
