#!/bin/bash
# v6 clip-apply (TMA-fed, dynamic tickets in both passes, exact norm accumulation)
mkdir -p gpurun_out
L=gradient-accumulation-tf-estimator_b200/csrc
B="python bench.py --steps 400 --warmup 10 --e2e-steps 0 --model-steps 0 --cpu-budget 0 --parity-steps 0"
summ() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['value']), 'apply_us', round(d['roofline']['avg_launch_us'],1), 'frac', round(d['roofline']['frac'],3), 'acc_us', round(d['roofline_accumulate']['avg_launch_us'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== timeline"; timeout 300 python tools/cta_timeline.py 2>&1 | tail -12
timeout 300 $B 2>/dev/null | summ "v6 p1slots=4"
GACCUM_LIB=$L/libgaccum_p1s3.so timeout 300 $B 2>/dev/null | summ "v6 p1slots=3"
GACCUM_LIB=$L/libgaccum_p1s2.so timeout 300 $B 2>/dev/null | summ "v6 p1slots=2"
for t in 0 5; do GACCUM_TMEM_TILES=$t timeout 300 $B 2>/dev/null | summ "v6 tmem_tiles=$t"; done
timeout 300 $B --workload bert_base 2>/dev/null | summ "v6 bert_base"
timeout 300 $B --workload bert_large --steps 128 2>/dev/null | summ "v6 bert_large"
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/r02f_err_default.log | tee gpurun_out/r02f_bench_default.json | summ "driver-like"
for tool in memcheck racecheck synccheck; do echo "== compute-sanitizer $tool"; timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "trajectory and bert-1.0 and 3 or unaligned or reproducible" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|error|hazard" | tail -6; done
