#!/bin/bash
# round-2 single-GPU check: parity tests, then A/B of the clip-apply kernel variants
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest.log
tail -5 gpurun_out/r02_pytest.log
B="python bench.py --steps 400 --warmup 10 --e2e-steps 0 --model-steps 0 --cpu-budget 0 --parity-steps 0"
summ() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['value']), 'apply_us', round(d['roofline']['avg_launch_us'],1), 'frac', round(d['roofline']['frac'],3), 'acc_us', round(d['roofline_accumulate']['avg_launch_us'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
for f in 0 2 4 6; do GACCUM_FLAGS=$f timeout 300 $B 2>gpurun_out/r02_err_f$f.log | tee gpurun_out/r02_bench_f$f.json | summ "flags=$f"; done
for st in 4 5 6; do GACCUM_STASH_TILES=$st timeout 300 $B 2>/dev/null | tee gpurun_out/r02_bench_stash$st.json | summ "stash=$st"; done
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/r02_err_default.log | tee gpurun_out/r02_bench_default.json | summ "driver-like"
