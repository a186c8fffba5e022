#!/bin/bash
# hybrid kernel (own tiles static, rest static or ticketed, exact norm): correctness + same-GPU A/B vs v5p1s2
mkdir -p gpurun_out
L=gradient-accumulation-tf-estimator_b200/csrc
summ() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['value']), 'apply_us', round(d['roofline']['avg_launch_us'],1), 'frac', round(d['roofline']['frac'],3), 'acc_us', round(d['roofline_accumulate']['avg_launch_us'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
GACCUM_P1_DYNAMIC=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -2
for wl in "bert_small --steps 400" "bert_base --steps 320" "bert_large --steps 128"; do
  B="python bench.py --warmup 10 --e2e-steps 0 --model-steps 0 --cpu-budget 0 --parity-steps 0 --workload $wl"
  GACCUM_LIB=$L/libgaccum_v5p1s2.so timeout 300 $B 2>/dev/null | summ "$wl v5p1s2"
  GACCUM_P1_DYNAMIC=0 timeout 300 $B 2>/dev/null | summ "$wl hybrid static-p1"
  GACCUM_P1_DYNAMIC=1 timeout 300 $B 2>/dev/null | summ "$wl hybrid dynamic-p1"
  GACCUM_P1_DYNAMIC=0 GACCUM_LIB=$L/libgaccum_p1s4.so timeout 300 $B 2>/dev/null | summ "$wl hybrid static-p1 4slots"
  GACCUM_P1_DYNAMIC=1 GACCUM_LIB=$L/libgaccum_p1s4.so timeout 300 $B 2>/dev/null | summ "$wl hybrid dynamic-p1 4slots"
done
echo "== timeline (auto)"; timeout 300 python tools/cta_timeline.py 2>&1 | tail -12
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/r02j_err_default.log | tee gpurun_out/r02j_bench_default.json | summ "driver-like"
