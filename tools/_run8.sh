cd $GRAFT_REPO_ROOT
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3) > gpurun_out/r2_tests3.log 2>&1
(timeout 900 python bench.py > gpurun_out/r2_final_n1.json 2> gpurun_out/r2_final_n1.err)
(timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err)
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_w1_bwd_optim_tc -c 1 -o gpurun_out/r2_prof_w1tc_final -f python tools/profile_step.py --steps 1 > gpurun_out/r2_ncu_full2.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_v13.csv python tools/profile_step.py --steps 1 > gpurun_out/r2_ncu_ll2.log 2>&1
cat gpurun_out/r2_tests3.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_final_n1.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['predict']['value'], d['cpu_baseline']['value'], d['predict'].get('gene_remap',{}).get('ms_per_chunk'))
print(open('gpurun_out/r2_final_ref.json').read()[:600])
PY
