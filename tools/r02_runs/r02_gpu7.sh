#!/bin/bash
mkdir -p gpurun_out
L=gradient-accumulation-tf-estimator_b200/csrc
B="python bench.py --steps 400 --warmup 10 --e2e-steps 0 --model-steps 0 --cpu-budget 0 --parity-steps 0"
summ() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['value']), 'apply_us', round(d['roofline']['avg_launch_us'],1), 'frac', round(d['roofline']['frac'],3), 'acc_us', round(d['roofline_accumulate']['avg_launch_us'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'], d.get('parity',{}).get('max_rel_err'))"; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== timeline"; timeout 300 python tools/cta_timeline.py 2>&1 | tail -12
timeout 300 $B 2>/dev/null | summ "v6b p1slots=4"
GACCUM_LIB=$L/libgaccum_p1s3.so timeout 300 $B 2>/dev/null | summ "v6b p1slots=3"
timeout 300 $B --workload bert_base 2>/dev/null | summ "v6b bert_base"
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/r02g_err_default.log | tee gpurun_out/r02g_bench_default.json | summ "driver-like"
echo "== synccheck, Tensor Memory off (no tcgen05.alloc executed)"; GACCUM_TMEM_TILES=0 timeout 600 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "trajectory and bert-1.0 and 3" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Barrier error" | tail -4
echo "== synccheck, default"; timeout 600 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "trajectory and bert-1.0 and 3" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Barrier error|located" | tail -4
