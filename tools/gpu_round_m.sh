#!/bin/bash
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_surface.py -m gpu -q -p no:cacheprovider -x -k host > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
for C in 12 16 24 32; do
SCNB_FETCH_CTAS=$C FEED=mapped timeout 300 python tools/timeline_e2e.py > gpurun_out/${TAG}_tl_mapped_c$C.txt 2>&1
SCNB_FETCH_CTAS=$C FEED=mapped python - <<PY
import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import bench
class A: steps=300; warmup=10
dev=torch.device("cuda",0)
ms,h2d,d2h,n=bench.time_e2e(A, dev, 0, 1, feed="mapped")
print("fetch CTAs", os.environ["SCNB_FETCH_CTAS"], "e2e mapped", round(512/ms/1e3,3), "M cells/s", round(ms*1e3,1), "us/step")
PY
done
