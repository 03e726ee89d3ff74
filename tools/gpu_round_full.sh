#!/bin/bash
# Full GPU-box visit for the record: tests, smoke, bench (ours + reference arm), timeline, ncu launch list of the bench command
# and one --set full capture of the step's top kernels.  Everything lands in gpurun_out/ with a tag.
TAG=${1:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${TAG}_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
timeout 600 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
echo "bench rc=$?" >> gpurun_out/${TAG}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
timeout 600 python bench.py --config 3 --steps 100 --warmup 10 --no-cpu-baseline --no-gpu-eager --no-predict > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
timeout 600 python bench.py --config 5 --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-predict > gpurun_out/${TAG}_bench_c5.json 2> gpurun_out/${TAG}_bench_c5.err
timeout 600 python bench.py --config 6 --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_bench_c6.json 2> gpurun_out/${TAG}_bench_c6.err
timeout 300 python tools/timeline.py > gpurun_out/${TAG}_timeline.txt 2>&1
timeout 300 python tools/timeline.py --config 5 > gpurun_out/${TAG}_timeline_c5.txt 2>&1
timeout 300 python tools/bench_gemm.py > gpurun_out/${TAG}_gemm.txt 2>&1
FEED=mapped timeout 300 python tools/timeline_e2e.py > gpurun_out/${TAG}_timeline_e2e_mapped.txt 2>&1
# ncu: only after the same command has exited 0 without it
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-predict --no-e2e > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 300 python tools/profile_step.py --steps 2 > gpurun_out/${TAG}_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_w1_bwd_optim_tc|k_csr_dense_fwd_tc|k_hs_fwd|k_hs_bwd|k_hs_dan_head|k_hs_classif_head|k_gather_rows" -c 12 -o gpurun_out/${TAG}_prof python tools/profile_step.py --steps 2 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -5 gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_smoke.log; head -c 600 gpurun_out/${TAG}_bench_n1.json; tail -3 gpurun_out/${TAG}_bench_n1.err; tail -2 gpurun_out/${TAG}_ncu_full.log
