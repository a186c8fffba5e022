#!/bin/bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29537"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
timeout 300 $T tools/dp_timeline.py 2>/dev/null | grep -v "^\*\|OMP_NUM\|^$\|NCCL version\|isolated\|last block" | head -22
timeout 300 $T tools/dp_timeline.py bert_large 2>/dev/null | grep -v "^\*\|OMP_NUM\|^$\|NCCL version\|isolated\|last block" | head -22
timeout 600 $T bench.py --gpus 2 --steps 200 --warmup 5 --e2e-steps 0 --model-steps 0 --cpu-budget 0 2>/dev/null | grep '^{' | tail -1 > gpurun_out/r02_dp2g_fused.json
python -c "
import json; d=json.load(open('gpurun_out/r02_dp2g_fused.json')); print('fused W=2', round(d['value']), 'apply-step us', round(d['roofline']['avg_launch_us'],1), 'acc us', round(d['roofline_accumulate']['avg_launch_us'],1), 'parity', d['parity']['max_rel_err'], d['parity']['replicas_identical'])"
timeout 600 $T bench.py --gpus 2 --steps 64 --warmup 3 --e2e-steps 0 --model-steps 0 --cpu-budget 0 --workload bert_large --parity-steps 0 2>/dev/null | grep '^{' | tail -1 > gpurun_out/r02_dp2g_large.json
python -c "
import json; d=json.load(open('gpurun_out/r02_dp2g_large.json')); print('fused W=2 bert_large', round(d['value']), 'apply-step us', round(d['roofline']['avg_launch_us'],1), 'acc us', round(d['roofline_accumulate']['avg_launch_us'],1))"
