set -x
cd $GRAFT_REPO_ROOT
(timeout 600 python -m pytest tests/test_gene_remap.py -x -q -m gpu 2>&1 | tail -15) > gpurun_out/r2_remap_test.log 2>&1
(timeout 300 python tools/exp_w1tc.py 2>&1 | tail -20) > gpurun_out/r2_exp_w1tc.log 2>&1
(SCNB_W1TC_DBG=1 SCNB_W1TC_EXP=0 timeout 120 python tools/dbg_w1tc.py 2>&1 | tail -12) > gpurun_out/r2_dbg0.log 2>&1
(SCNB_W1TC_DBG=1 SCNB_W1TC_EXP=1 timeout 120 python tools/dbg_w1tc.py 2>&1 | tail -8) > gpurun_out/r2_dbg1.log 2>&1
(SCNB_W1TC_DBG=1 SCNB_W1TC_EXP=2 timeout 120 python tools/dbg_w1tc.py 2>&1 | tail -8) > gpurun_out/r2_dbg2.log 2>&1
cat gpurun_out/r2_remap_test.log gpurun_out/r2_exp_w1tc.log
