cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6) > gpurun_out/r2_tests.log 2>&1
(timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err)
cat gpurun_out/r2_tests.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n1.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['predict']['value'], d['predict'].get('gene_remap'))
PY
tail -3 gpurun_out/r2_bench_n1.err
