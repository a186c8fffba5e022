"""PCIe probe: what H2D rate can the host-buffer path expect on this box?"""
import time, torch
n = 28764674
big = torch.empty(n, dtype=torch.float32).pin_memory()
dev = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def rate(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return n * 4 * reps / (time.perf_counter() - t0) / 1e9
def one():
    with torch.cuda.stream(s1): dev.copy_(big, non_blocking=True)
h = n // 2
def two():
    with torch.cuda.stream(s1): dev[:h].copy_(big[:h], non_blocking=True)
    with torch.cuda.stream(s2): dev[h:].copy_(big[h:], non_blocking=True)
back = torch.empty(n, dtype=torch.float32).pin_memory()
def both_dirs():
    with torch.cuda.stream(s1): dev.copy_(big, non_blocking=True)
    with torch.cuda.stream(s2): back.copy_(dev, non_blocking=True)
print("H2D one stream      %.1f GB/s" % rate(one))
print("H2D two streams     %.1f GB/s" % rate(two))
print("H2D + D2H together  %.1f GB/s each way" % rate(both_dirs))
import os
print("affinity", len(os.sched_getaffinity(0)))
