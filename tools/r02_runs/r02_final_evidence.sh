#!/bin/bash
# final single-GPU evidence with the shipped sources: launch list, ncu --set full, measured DRAM traffic
mkdir -p gpurun_out
BN="python bench.py --steps 20 --warmup 3 --e2e-steps 0 --model-steps 0 --cpu-budget 0 --parity-steps 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 40 --csv --log-file gpurun_out/r02_launches.csv $BN > gpurun_out/r02_launches_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:apply_clip -c 2 -o gpurun_out/r02_prof_apply -f $BN > /dev/null 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:accumulate -c 2 -o gpurun_out/r02_prof_acc -f $BN > /dev/null 2>&1
timeout 1500 python tools/measure_traffic.py bert_small bert_base bert_large 2>&1 | tail -4
cp profiles/traffic.json gpurun_out/traffic.json
timeout 600 python bench.py --steps 20 --warmup 3 2>/dev/null | tee gpurun_out/r02_final_bench_1gpu.json | python -c "import json,sys; d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print(round(d['value']), d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['parity']['max_rel_err'])"
