#!/bin/bash
# A/B in ONE run on ONE GPU: v5 (static pass 1, dynamic pass 2) vs v6c (dynamic both passes, exact norm accumulation)
mkdir -p gpurun_out
L=gradient-accumulation-tf-estimator_b200/csrc
B="python bench.py --steps 400 --warmup 10 --e2e-steps 0 --model-steps 0 --cpu-budget 0 --parity-steps 0"
summ() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['value']), 'apply_us', round(d['roofline']['avg_launch_us'],1), 'frac', round(d['roofline']['frac'],3), 'acc_us', round(d['roofline_accumulate']['avg_launch_us'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_golden_and_api.py -m gpu -x -q 2>&1 | tail -3
for rep in 1 2; do
  GACCUM_LIB=$L/libgaccum_v5.so timeout 300 $B 2>/dev/null | summ "rep$rep v5 p1s4"
  GACCUM_LIB=$L/libgaccum_v5p1s2.so timeout 300 $B 2>/dev/null | summ "rep$rep v5 p1s2"
  timeout 300 $B 2>/dev/null | summ "rep$rep v6c p1s4"
  GACCUM_LIB=$L/libgaccum_p1s2.so timeout 300 $B 2>/dev/null | summ "rep$rep v6c p1s2"
done
echo "== timeline v6c"; timeout 300 python tools/cta_timeline.py 2>&1 | tail -12
