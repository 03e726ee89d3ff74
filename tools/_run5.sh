cd $GRAFT_REPO_ROOT
(timeout 300 python -m pytest tests/test_gene_remap.py -x -q -m gpu 2>&1 | tail -3) > gpurun_out/r2_remap_test2.log 2>&1
(timeout 600 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err)
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_b.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline']['frac'], d['predict'].get('gene_remap'))
PY
# launch list of one eager step (cold-cache, serialised: shares only)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_v12.csv python tools/profile_step.py --steps 1 > gpurun_out/r2_ncu_ll.log 2>&1
# full capture of the dW1+optimizer kernel (ring + sink form) and of the remap fill kernel
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_w1_bwd_optim_tc -c 1 -o gpurun_out/r2_prof_w1tc_ring -f python tools/profile_step.py --steps 1 > gpurun_out/r2_ncu_full.log 2>&1
cat > /tmp/remap_drv.py <<'PY'
import sys, torch
sys.path.insert(0, '.')
import bench
from scnym_b200.synth import synth_device_csr
from scnym_b200.dataprep import remap_columns
c = bench.CFG
tile, _ = synth_device_csr(1 << 16, c["G"], c["density"], c["C"], seed=2, device="cuda")
g = torch.Generator().manual_seed(7)
cm = torch.randperm(c["G"], generator=g).to(torch.int32); cm[torch.rand(c["G"], generator=g) < 0.1] = -1
cm = cm.cuda()
for _ in range(2):
    remap_columns(tile, cm, c["G"])
torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_remap -c 8 -o gpurun_out/r2_prof_remap -f python /tmp/remap_drv.py > gpurun_out/r2_ncu_remap.log 2>&1
cat gpurun_out/r2_remap_test2.log; tail -2 gpurun_out/r2_ncu_ll.log gpurun_out/r2_ncu_full.log gpurun_out/r2_ncu_remap.log
