#!/bin/bash
# final verification run of the session: the whole GPU suite (incl. the U-Net family tests, first time on a GPU) + smoke
mkdir -p gpurun_out
S=$(date +%s)
timeout 420 python -m pytest tests -m gpu -q -s -rf 2>&1 > gpurun_out/r2f_all_raw.log
grep -E "^\[zoo (eval|train)\] (unet|unetp|unetpp|unet2|unet_rexnet13):|passed|failed|^FAILED|^E  " gpurun_out/r2f_all_raw.log | cut -c1-600 > gpurun_out/r2f_all.log
tail -30 gpurun_out/r2f_all.log
echo "== suite done at $(( $(date +%s) - S )) s"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
