#!/bin/bash
# round-2 single-GPU check of the "ring is the stash" clip-apply kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
tail -5 gpurun_out/r02b_pytest.log
B="python bench.py --steps 400 --warmup 10 --e2e-steps 0 --model-steps 0 --cpu-budget 0 --parity-steps 0"
summ() { python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$1', round(d['value']), 'apply_us', round(d['roofline']['avg_launch_us'],1), 'frac', round(d['roofline']['frac'],3), 'acc_us', round(d['roofline_accumulate']['avg_launch_us'],1), d['clocks']['sm_mhz'], d['clocks']['reasons'])"; }
for f in 0 2 4 6; do GACCUM_FLAGS=$f timeout 300 $B 2>gpurun_out/r02b_err_f$f.log | tee gpurun_out/r02b_bench_f$f.json | summ "flags=$f"; done
for st in 4 6 7; do GACCUM_STASH_TILES=$st timeout 300 $B 2>/dev/null | tee gpurun_out/r02b_bench_stash$st.json | summ "slots=$st"; done
L=gradient-accumulation-tf-estimator_b200/csrc
for v in pf1 pf2 pf5; do GACCUM_LIB=$L/libgaccum_$v.so timeout 300 $B 2>/dev/null | tee gpurun_out/r02b_bench_$v.json | summ "variant=$v"; done
timeout 300 python tools/cta_timeline.py > gpurun_out/r02b_timeline.txt 2>&1; cat gpurun_out/r02b_timeline.txt
GACCUM_FLAGS=2 timeout 300 python tools/cta_timeline.py > gpurun_out/r02b_timeline_f2.txt 2>&1; cat gpurun_out/r02b_timeline_f2.txt
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/r02b_err_default.log | tee gpurun_out/r02b_bench_default.json | summ "driver-like"
