#!/bin/bash
# usage: tools/run_scaling.sh "<N list>" [bench args]  -- launches bench.py exactly as the driver does
ns="$1"; shift
for n in $ns; do
  if [ "$n" = "1" ]; then
    timeout 600 python bench.py --gpus 1 "$@" > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $n "$@" 2> gpurun_out/scale_$n.err | grep '^{' > gpurun_out/scale_$n.json
  fi
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_$n.json"))
    wm=d.get("with_model") or {}
    print("N=$n value", round(d["value"]), "ms/step", round(d["ms_per_step"]*1e3,1), "apply-step us", round(d["roofline"]["avg_launch_us"],1), "e2e", round((d.get("e2e") or {}).get("value",0)), "with_model", round(wm.get("value",0),1), "share", round(wm.get("train_op_share",0),3))
except Exception as e:
    print("N=$n failed:", e); print(open("gpurun_out/scale_$n.err").read()[-1500:])
PY
done
