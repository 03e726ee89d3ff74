#!/bin/bash
# ncu --set full capture of the config-5 kernels (after the same command has exited 0 without ncu)
TAG=${1:-n5}
mkdir -p gpurun_out
timeout 300 python tools/profile_step.py --config 5 --steps 1 > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_gemm_planes|k_bn_act_fwd_v8|k_bn_act_bwd_v8|k_csr_dense_fwd_tc|k_w1_bwd_optim_tc|k_colsum" -c 40 -o gpurun_out/${TAG}_prof python tools/profile_step.py --config 5 --steps 1 > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_plain.log; tail -3 gpurun_out/${TAG}_ncu_full.log
