#!/bin/bash
# instrumented pass-1 analysis of the clip-apply kernel variants
mkdir -p gpurun_out
L=gradient-accumulation-tf-estimator_b200/csrc
B="python bench.py --steps 400 --warmup 10 --e2e-steps 0 --model-steps 0 --cpu-budget 0 --parity-steps 0"
summ() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['value']), 'apply_us', round(d['roofline']['avg_launch_us'],1), 'frac', round(d['roofline']['frac'],3), 'acc_us', round(d['roofline_accumulate']['avg_launch_us'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
GACCUM_LIB=$L/libgaccum_atma.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden_and_api.py -m gpu -x -q 2>&1 | tail -3
for st in 8 4; do echo "== timeline default slots=$st"; GACCUM_STASH_TILES=$st timeout 300 python tools/cta_timeline.py 2>&1 | tail -12; done
for st in 4 3 2; do echo "== timeline atma slots=$st"; GACCUM_STASH_TILES=$st GACCUM_LIB=$L/libgaccum_atma_exp.so timeout 300 python tools/cta_timeline.py 2>&1 | tail -12; done
for st in 4 3 2; do GACCUM_STASH_TILES=$st GACCUM_LIB=$L/libgaccum_atma.so timeout 300 $B 2>/dev/null | summ "atma slots=$st"; done
for st in 5 3; do GACCUM_STASH_TILES=$st timeout 300 $B 2>/dev/null | summ "default slots=$st"; done
