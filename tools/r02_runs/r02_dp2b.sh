#!/bin/bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534"
for bps in 4 2 1; do echo "== blocks/SM $bps"; GACCUM_DP_BLOCKS_PER_SM=$bps timeout 300 $T tools/dp_timeline.py 2>/dev/null | grep -v "^\*\|OMP_NUM\|^$"; done
