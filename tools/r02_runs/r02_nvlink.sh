#!/bin/bash
# NVLink data counters around K apply steps (tools/nvlink_bytes.py); usage: r02_nvlink.sh W
W=${1:-2}
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29517 tools/nvlink_bytes.py bert_small 100 > gpurun_out/nvlink_${W}gpu.log 2>&1
tail -12 gpurun_out/nvlink_${W}gpu.log
