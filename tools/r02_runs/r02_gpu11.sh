#!/bin/bash
mkdir -p gpurun_out
summ() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['value']), 'apply_us', round(d['roofline']['avg_launch_us'],1), 'frac', round(d['roofline']['frac'],3), 'acc_us', round(d['roofline_accumulate']['avg_launch_us'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'], 'traffic', d['roofline']['traffic'])"; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 400 --warmup 10 --e2e-steps 0 --model-steps 0 --cpu-budget 0 --parity-steps 0 2>/dev/null | summ "shipped"
echo "== synccheck"; timeout 600 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "trajectory and bert-1.0 and 3 or reproducible" > gpurun_out/r02_sanitizer_synccheck2.log 2>&1; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/r02_sanitizer_synccheck2.log | tail -3; grep -A4 "Barrier error" gpurun_out/r02_sanitizer_synccheck2.log | head -12
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/r02k_err.log | tee gpurun_out/r02k_bench_default.json | summ "driver-like"
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02k_bench_reference.json 2>/dev/null; head -c 600 gpurun_out/r02k_bench_reference.json
