#!/bin/bash
# round-2 evidence: launch list, ncu --set full of the two kernels, measured DRAM traffic, sanitizers
mkdir -p gpurun_out
BN="python bench.py --steps 20 --warmup 3 --e2e-steps 0 --model-steps 0 --cpu-budget 0 --parity-steps 0"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 40 --csv --log-file gpurun_out/r02_launches.csv $BN > gpurun_out/r02_launches_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:apply_clip -c 2 -o gpurun_out/r02_prof_apply -f $BN > /dev/null 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:accumulate -c 2 -o gpurun_out/r02_prof_acc -f $BN > /dev/null 2>&1
timeout 1500 python tools/measure_traffic.py bert_small bert_base 2>&1 | tail -3
cp profiles/traffic.json gpurun_out/traffic.json
for tool in memcheck racecheck synccheck; do echo "== compute-sanitizer $tool" ; timeout 900 compute-sanitizer --tool $tool python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "trajectory and bert-1.0 and 3 or unaligned or reproducible" > gpurun_out/r02_sanitizer_$tool.log 2>&1; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02_sanitizer_$tool.log | tail -3; done
ls -la gpurun_out/r02_prof_*.ncu-rep gpurun_out/r02_launches.csv
