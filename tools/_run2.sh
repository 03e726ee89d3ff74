cd $GRAFT_REPO_ROOT
(timeout 600 python -m pytest tests/test_gene_remap.py -x -q -m gpu 2>&1 | tail -5) > gpurun_out/r2_remap_test.log 2>&1
(timeout 300 python tools/exp_w1tc.py 0 5 6 7 2>&1 | tail -20) > gpurun_out/r2_exp_w1tc_b.log 2>&1
(SCNB_W1TC_DBG=1 SCNB_W1TC_EXP=5 timeout 120 python tools/dbg_w1tc.py 2>&1 | grep w1tc | tail -4) > gpurun_out/r2_dbg5.log 2>&1
(SCNB_W1TC_DBG=1 SCNB_W1TC_EXP=6 timeout 120 python tools/dbg_w1tc.py 2>&1 | grep w1tc | tail -4) > gpurun_out/r2_dbg6.log 2>&1
cat gpurun_out/r2_remap_test.log gpurun_out/r2_exp_w1tc_b.log gpurun_out/r2_dbg5.log gpurun_out/r2_dbg6.log
