#!/bin/bash
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_kernels.py tests/test_full_shape_parity.py -m gpu -q -p no:cacheprovider -k "config5 or planes or tcgen05" > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
timeout 300 python tools/bench_gemm.py > gpurun_out/${TAG}_gemm.txt 2>&1
timeout 600 python bench.py --config 5 --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
tail -4 gpurun_out/${TAG}_tests.log; tail -4 gpurun_out/${TAG}_gemm.txt; head -c 300 gpurun_out/${TAG}_bench_c5.json
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_c5.json").read().strip().splitlines()[-1]); r=d["roofline"]["kernel_ms_per_step"]
print({k:v for k,v in r.items() if "TN" in k})
PY
