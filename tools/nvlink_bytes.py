#!/usr/bin/env python
"""NVLink bytes actually moved by dp_apply_kernel, from the driver's per-link data counters:
    python -m torch.distributed.run --nproc-per-node W --master-addr 127.0.0.1 tools/nvlink_bytes.py [workload] [windows]
Every rank reads `nvidia-smi nvlink -gt d -i <gpu>` (KiB per link, Tx and Rx) before and after K whole windows (K apply
launches, K*(N-1) rank-local accumulate launches that must move nothing), and rank 0 prints the per-apply deltas next
to what the algorithm needs: each GPU sends (W-1)/W of the a+G slab to the owners (phase A) and (W-1)/W... of its updated
shard to every peer (phase C) = 2 * (W-1)/W * 4 B * P_padded out, and receives the same."""
import os, re, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import gaccum_b200 as g
from gaccum_b200.manifests import MANIFESTS
from gaccum_b200.distributed import FusedDataParallelTrainOp

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
wl = sys.argv[1] if len(sys.argv) > 1 else "bert_small"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 50
N = 4
man = MANIFESTS[wl]()


def counters():
    out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(local)], capture_output=True, text=True).stdout
    tx = sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", out))
    rx = sum(int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", out))
    return tx * 1024, rx * 1024, len(re.findall(r"Data Tx:", out)), out


params = [torch.randn(s, device=dev) * 0.02 for _, s in man]
dp = FusedDataParallelTrainOp(params, [n for n, _ in man], g.HParams.bert(), N, lambda s: 1e-5, global_step=1)
grads = [torch.randn(s, device=dev) * 1e-3 for _, s in man]
b = dp.bind(grads)
for _ in range(2 * N):                       # warm-up: two windows
    dp.run_bound(b)
torch.cuda.synchronize(); dist.barrier(); time.sleep(0.5)
tx0, rx0, links, raw = counters()
dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(K * N):
    dp.run_bound(b)
e1.record()
torch.cuda.synchronize(); dist.barrier(); time.sleep(0.5)
tx1, rx1, _, _ = counters()
# accumulate-only stretch: N-1 launches per window never touch NVLink
while dp.global_step % N != 1:
    dp.run_bound(b)
torch.cuda.synchronize(); dist.barrier(); time.sleep(0.5)
tx2, rx2, _, _ = counters()
for _ in range(N - 1):
    dp.run_bound(b)
torch.cuda.synchronize(); time.sleep(0.5)
tx3, rx3, _, _ = counters()
padded = dp.plan.padded_size
need = 2 * (world - 1) / world * 4 * padded
t = torch.tensor([tx1 - tx0, rx1 - rx0, tx3 - tx2, rx3 - rx2], dtype=torch.float64, device=dev)
allr = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(allr, t)
if rank == 0:
    if links == 0:
        print("nvidia-smi nvlink -gt d printed no counters on this box:\n" + raw[:2000])
    print(f"{wl}: W={world} N={N} K={K} applies, padded slab {padded * 4 / 1e6:.1f} MB, {links} links per GPU, window {e0.elapsed_time(e1) * 1e3 / K:.1f} us")
    print(f"algorithmic NVLink bytes per GPU per apply: {need / 1e6:.1f} MB out and {need / 1e6:.1f} MB in (phase A + phase C)")
    for r, x in enumerate(allr):
        x = x.tolist()
        print(f"rank {r}: Tx {x[0] / K / 1e6:8.2f} MB/apply ({x[0] / K / need:.3f} x)   Rx {x[1] / K / 1e6:8.2f} MB/apply ({x[1] / K / need:.3f} x)"
              f"   | {N - 1} accumulate launches: Tx {x[2] / 1e6:.3f} MB  Rx {x[3] / 1e6:.3f} MB")
dist.barrier()
dist.destroy_process_group()
