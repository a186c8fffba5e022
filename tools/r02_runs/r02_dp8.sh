#!/bin/bash
# 8-GPU: parity tests of the fused kernel at W=8, timeline, driver-style bench lines (BERT-Small, BERT-Large)
mkdir -p gpurun_out
W=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29540"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "fused_dp_kernel or host_buffer" 2>&1 | tail -3
timeout 300 $T tools/dp_timeline.py 2>/dev/null | grep -v "^\*\|OMP_NUM\|^$\|NCCL version\|isolated" | head -14
timeout 600 $T bench.py --gpus $W --steps 200 --warmup 5 --cpu-budget 0 2>gpurun_out/r02_dp${W}_err.log | grep '^{' | tail -1 > gpurun_out/r02_dp${W}_bert_small.json
python -c "
import json; d=json.load(open('gpurun_out/r02_dp${W}_bert_small.json')); print('fused W=$W', round(d['value']), 'apply-step us', round(d['roofline']['avg_launch_us'],1), 'acc us', round(d['roofline_accumulate']['avg_launch_us'],1), 'parity', d['parity']['max_rel_err'], d['parity']['replicas_identical'], 'e2e', round(d['e2e']['value']), 'with_model', d.get('with_model',{}).get('value'))"
timeout 600 $T bench.py --gpus $W --steps 128 --warmup 3 --cpu-budget 0 --workload bert_large --e2e-steps 0 --model-steps 0 2>>gpurun_out/r02_dp${W}_err.log | grep '^{' | tail -1 > gpurun_out/r02_dp${W}_bert_large.json
python -c "
import json; d=json.load(open('gpurun_out/r02_dp${W}_bert_large.json')); print('fused W=$W bert_large', round(d['value']), 'apply-step us', round(d['roofline']['avg_launch_us'],1), 'acc us', round(d['roofline_accumulate']['avg_launch_us'],1), 'parity', d['parity']['max_rel_err'], d['parity']['replicas_identical'])"
timeout 600 $T bench.py --gpus $W --steps 200 --warmup 5 --cpu-budget 0 --dp allreduce --e2e-steps 0 --model-steps 0 2>>gpurun_out/r02_dp${W}_err.log | grep '^{' | tail -1 > gpurun_out/r02_dp${W}_allreduce.json
python -c "
import json; d=json.load(open('gpurun_out/r02_dp${W}_allreduce.json')); print('nccl allreduce W=$W', round(d['value']), 'apply-step us', round(d['roofline']['avg_launch_us'],1))"
tail -3 gpurun_out/r02_dp${W}_err.log
