#!/bin/bash
# usage: tools/run_tunes_dp.sh <ngpus> "<tune list>" [extra bench args]
n="$1"; tunes="$2"; shift; shift
for t in $tunes; do
  echo -n "N=$n TUNE=$t $@ : "
  GACCUM_EXPERIMENTS=1 GACCUM_TUNE=$t timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 400 --warmup 10 --e2e-steps 0 --model-steps 0 --cpu-budget 0 "$@" 2>/dev/null | grep '^{' | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), 'ms/step', round(d['ms_per_step']*1e3,1), 'apply-step us', round(d['roofline']['avg_launch_us'],1), 'acc us', round(d['roofline_accumulate']['avg_launch_us'],1))"
done
