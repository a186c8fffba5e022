#!/bin/bash
# usage: tools/run_tunes.sh "<tune list>" [extra bench args]   -- prints value, apply us, frac, acc us, frac
tunes="$1"; shift
for t in $tunes; do
  echo -n "TUNE=$t $@ : "
  GACCUM_EXPERIMENTS=1 GACCUM_TUNE=$t python bench.py --steps 400 --warmup 10 --e2e-steps 0 --model-steps 0 --cpu-budget 0 "$@" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['roofline']['avg_launch_us'],1), round(d['roofline']['frac'],3), round(d['roofline_accumulate']['avg_launch_us'],1), round(d['roofline_accumulate']['frac'],3))"
done
