#!/bin/bash
TAG=${1:-run}
mkdir -p gpurun_out
for C in 8 16 32 64; do
SCNB_FETCH_CTAS=$C FEED=mapped timeout 300 python tools/timeline_e2e.py > gpurun_out/${TAG}_tl_mapped_c$C.txt 2>&1
echo "== $C CTAs"; head -4 gpurun_out/${TAG}_tl_mapped_c$C.txt | grep steps; grep "k_fetch_rows\|k_hs_fwd" gpurun_out/${TAG}_tl_mapped_c$C.txt | head -6
done
