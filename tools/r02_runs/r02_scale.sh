#!/bin/bash
# driver-style scaling line at W GPUs (BERT-Small default bench) + BERT-Large (BASELINE config 5) with per-rank diagnostics
mkdir -p gpurun_out
W=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $T bench.py --gpus $W --steps 200 --warmup 5 --cpu-budget 0 2>gpurun_out/r02_scale${W}_err.log | grep '^{' | tail -1 > gpurun_out/r02_scale${W}_bert_small.json
python -c "
import json; d=json.load(open('gpurun_out/r02_scale${W}_bert_small.json')); print('W=$W bert_small', round(d['value']), 'apply-step us', round(d['roofline']['avg_launch_us'],1), 'acc us', round(d['roofline_accumulate']['avg_launch_us'],1), 'parity', d['parity']['max_rel_err'], d['parity']['replicas_identical'], 'e2e', round(d['e2e']['value']), 'with_model', round(d.get('with_model',{}).get('value',0),1)); print(' per_rank', d['per_rank'])"
if [ "$2" = "large" ]; then
timeout 600 $T bench.py --gpus $W --steps 128 --warmup 3 --cpu-budget 0 --workload bert_large --e2e-steps 0 --model-steps 0 2>>gpurun_out/r02_scale${W}_err.log | grep '^{' | tail -1 > gpurun_out/r02_scale${W}_bert_large.json
python -c "
import json; d=json.load(open('gpurun_out/r02_scale${W}_bert_large.json')); print('W=$W bert_large', round(d['value']), 'apply-step us', round(d['roofline']['avg_launch_us'],1), 'acc us', round(d['roofline_accumulate']['avg_launch_us'],1), 'parity', d['parity']['max_rel_err'], d['parity']['replicas_identical']); print(' per_rank', d['per_rank'])"
fi
tail -2 gpurun_out/r02_scale${W}_err.log
