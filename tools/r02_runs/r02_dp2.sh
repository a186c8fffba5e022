#!/bin/bash
# 2-GPU: parity of the push-based dp_apply_kernel (tests), then short benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 "$@" 2>gpurun_out/r02_dp2_err.log | grep '^{' | tail -1; }
run --steps 200 --warmup 5 --e2e-steps 16 --model-steps 0 --cpu-budget 0 > gpurun_out/r02_dp2_fused.json
python -c "
import json; d=json.load(open('gpurun_out/r02_dp2_fused.json')); print('fused W=2', round(d['value']), 'apply-step us', round(d['roofline']['avg_launch_us'],1), 'acc us', round(d['roofline_accumulate']['avg_launch_us'],1), 'parity', d.get('parity'), 'e2e', d.get('e2e',{}).get('value'), d['gpu_launches'], d['config']['apply_launches'])"
run --steps 200 --warmup 5 --e2e-steps 0 --model-steps 0 --cpu-budget 0 --dp allreduce > gpurun_out/r02_dp2_allreduce.json
python -c "
import json; d=json.load(open('gpurun_out/r02_dp2_allreduce.json')); print('allreduce W=2', round(d['value']), 'apply-step us', round(d['roofline']['avg_launch_us'],1), 'parity', d.get('parity'))"
tail -5 gpurun_out/r02_dp2_err.log
