#!/bin/bash
TAG=${1:-run}
L=gpurun_out/${TAG}_dp_local.log
for T in 1000000 2000000 3950000 7900000; do echo "== total $T" >> $L; python tools/exp_dp_local.py --total $T 2>&1 | grep warm >> $L; done
for P in 4 8 12; do echo "== probe $P (4 = no stores, 8 = no optimizer math)" >> $L; SCNB_DP_PROBE=$P python tools/exp_dp_local.py 2>&1 | grep "dp_.*warm" >> $L; done
cat $L
