#!/bin/bash
# exchange-kernel experiments on N GPUs: full kernel, without peer stores, without peer loads, neither; config-3 arena
TAG=${1:-run}; N=${2:-2}
mkdir -p gpurun_out
for P in 0 1 2 3; do
  echo "== SCNB_DP_PROBE=$P" >> gpurun_out/${TAG}_dpx.log
  SCNB_DP_PROBE=$P timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/bench_dp_exchange.py --total 7900000 2>/dev/null | grep "^rank" >> gpurun_out/${TAG}_dpx.log
done
echo "== one rank (local traffic only)" >> gpurun_out/${TAG}_dpx.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29522 tools/bench_dp_exchange.py --total 3950000 2>/dev/null | grep "^rank" >> gpurun_out/${TAG}_dpx.log
cat gpurun_out/${TAG}_dpx.log
