#!/usr/bin/env python
"""Measures roofline.traffic: dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel
(ncu, on the GPU box) and stamps the capture with a hash of the kernel sources, so bench.py only quotes it
while it still describes the code that is running (a stale or foreign capture reads as null).

    python tools/measure_traffic.py [workload ...]        # default: bert_small bert_base; rewrites profiles/traffic.json
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

KERNEL = {"bert_small": "apply_clip_kernel", "bert_base": "apply_clip_kernel", "bert_large": "apply_clip_kernel", "mnist_cnn": "apply_kernel"}


def capture(workload):
    kern = KERNEL[workload]
    log = os.path.join(ROOT, "gpurun_out", f"traffic_{workload}.csv")
    os.makedirs(os.path.dirname(log), exist_ok=True)
    n_acc = bench.WORKLOADS[workload][1]
    steps = max(48, 3 * n_acc)                    # at least 3 applies in the timed region, one is skipped
    cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum", "--clock-control", "none",
           "--profile-from-start", "off", "-k", f"regex:{kern}", "-s", "1", "-c", "2" if n_acc > 8 else "4", "--csv", "--log-file", log,
           sys.executable, os.path.join(ROOT, "bench.py"), "--workload", workload, "--steps", str(steps), "--warmup", "3",
           "--e2e-steps", "0", "--model-steps", "0", "--cpu-budget", "0", "--parity-steps", "0"]
    subprocess.run(cmd, check=True, cwd=ROOT, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=1200)
    text = open(log).read()
    text = text[text.index('"ID"'):]
    rows = list(csv.DictReader(io.StringIO(text)))
    per = {}
    for r in rows:
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"].lower()
        mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(unit, 1)
        per.setdefault(r["ID"], {})[r["Metric Name"]] = v * mult
    n = len(per)
    rd = sum(p["dram__bytes_read.sum"] for p in per.values()) / n
    wr = sum(p["dram__bytes_write.sum"] for p in per.values()) / n
    us = sum(p["gpu__time_duration.sum"] for p in per.values()) / n
    return {"kernel": kern, "dram_bytes": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr),
            "launches": n, "ncu_duration_us": us, "stamp": bench.kernel_source_stamp(),
            "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, mean over the captured launches (ncu flushes caches between replays)"}


if __name__ == "__main__":
    wls = sys.argv[1:] or ["bert_small", "bert_base"]
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        doc = json.load(open(path))
        assert "captures" in doc
    except Exception:
        doc = {"captures": {}}
    for w in wls:
        doc["captures"][w] = capture(w)
        print(w, doc["captures"][w])
        json.dump(doc, open(path, "w"), indent=1)
