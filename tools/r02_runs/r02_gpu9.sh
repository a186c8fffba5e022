#!/bin/bash
# same-GPU A/B across workloads: v5 (static pass 1) with 2/3/4 pass-1 slots vs v6c (dynamic both passes)
mkdir -p gpurun_out
L=gradient-accumulation-tf-estimator_b200/csrc
summ() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['value']), 'apply_us', round(d['roofline']['avg_launch_us'],1), 'frac', round(d['roofline']['frac'],3), 'acc_us', round(d['roofline_accumulate']['avg_launch_us'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
for wl in "bert_small --steps 400" "bert_base --steps 320" "bert_large --steps 128"; do
  B="python bench.py --warmup 10 --e2e-steps 0 --model-steps 0 --cpu-budget 0 --parity-steps 0 --workload $wl"
  for v in v5p1s2 v5p1s3 v5; do GACCUM_LIB=$L/libgaccum_$v.so timeout 300 $B 2>/dev/null | summ "$wl $v"; done
  timeout 300 $B 2>/dev/null | summ "$wl v6c"
  GACCUM_LIB=$L/libgaccum_p1s2.so timeout 300 $B 2>/dev/null | summ "$wl v6c-p1s2"
done
