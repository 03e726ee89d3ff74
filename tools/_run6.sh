cd $GRAFT_REPO_ROOT
(timeout 300 python tools/exp_w1tc.py 0 10 9 2>&1 | tail -12) > gpurun_out/r2_exp_w1tc_d.log 2>&1
(SCNB_W1TC_DBG=1 timeout 120 python tools/dbg_w1tc.py 2>&1 | grep w1tc | tail -4) > gpurun_out/r2_dbg9b.log 2>&1
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3) > gpurun_out/r2_tests2.log 2>&1
(timeout 600 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err)
cat gpurun_out/r2_exp_w1tc_d.log gpurun_out/r2_dbg9b.log gpurun_out/r2_tests2.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_c.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['predict'].get('gene_remap'))
PY
