#!/bin/bash
TAG=${1:-run}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_full_shape_parity.py tests/test_cuda_parity.py tests/test_surface.py tests/test_engine_misc.py -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-gpu-eager --no-predict > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench_n1.err
timeout 600 python tools/timeline.py > gpurun_out/${TAG}_timeline.txt 2>&1
timeout 300 python tools/profile_step.py --steps 2 > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_hs_fwd|k_hs_bwd|k_teacher_head|k_classif_mixmatch|k_dan_tail|k_hs_act|k_hs_mix" -c 14 -o gpurun_out/${TAG}_prof python tools/profile_step.py --steps 2 > gpurun_out/${TAG}_ncu.log 2>&1
tail -5 gpurun_out/${TAG}_tests.log; head -c 300 gpurun_out/${TAG}_bench_n1.json; tail -3 gpurun_out/${TAG}_bench_n1.err; tail -3 gpurun_out/${TAG}_ncu.log
