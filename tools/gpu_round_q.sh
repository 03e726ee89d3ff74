#!/bin/bash
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_kernels.py tests/test_cuda_parity.py tests/test_full_shape_parity.py tests/test_gene_remap.py tests/test_surface.py -m gpu -q -p no:cacheprovider -k "fused_bn_eval or predict or tensor_core" > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-e2e > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
SCNB_PREDICT_FUSE_BN=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-e2e > gpurun_out/${TAG}_bench_unfused.json 2> gpurun_out/${TAG}_bench_unfused.err
tail -5 gpurun_out/${TAG}_tests.log
python - <<PY
import json
for f in ("gpurun_out/${TAG}_bench.json","gpurun_out/${TAG}_bench_unfused.json"):
    d=json.loads(open(f).read().strip().splitlines()[-1]); p=d["predict"]; print(f, p["value"], p["kernel_ms_per_chunk"])
PY
