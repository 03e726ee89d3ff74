cd $GRAFT_REPO_ROOT
(timeout 300 python tools/exp_w1tc.py 0 6 8 9 2>&1 | tail -20) > gpurun_out/r2_exp_w1tc_c.log 2>&1
(SCNB_W1TC_DBG=1 SCNB_W1TC_EXP=8 timeout 120 python tools/dbg_w1tc.py 2>&1 | grep w1tc | tail -4) > gpurun_out/r2_dbg8.log 2>&1
(SCNB_W1TC_DBG=1 SCNB_W1TC_EXP=9 timeout 120 python tools/dbg_w1tc.py 2>&1 | grep w1tc | tail -4) > gpurun_out/r2_dbg9.log 2>&1
cat gpurun_out/r2_exp_w1tc_c.log gpurun_out/r2_dbg8.log gpurun_out/r2_dbg9.log
