cd $GRAFT_REPO_ROOT
for bn in 0 16 64 128; do echo "== SCNB_GEMM_TC_BN=$bn"; SCNB_GEMM_TC_BN=$bn timeout 120 python tools/bench_gemm_small.py 2>&1 | tail -6; done > gpurun_out/r2_gemm_small.log 2>&1
cat gpurun_out/r2_gemm_small.log
