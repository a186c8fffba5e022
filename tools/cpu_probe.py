import os, sys, time
os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
sys.path.insert(0, "oracle")
import numpy as np, oracle_c, oracle_np as onp
man = onp.MANIFESTS["bert_small"]()
rng = np.random.default_rng(0)
params = [rng.normal(0, 0.02, s).astype(np.float32) for _, s in man]
op = oracle_c.COracleTrainOp(params, [n for n, _ in man], onp.HParams.bert(), 4, init_lr=2e-5, num_train_steps=207900, num_warmup_steps=20790, global_step=1)
grads = [rng.normal(0, 1e-3, s).astype(np.float32) for _, s in man]
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), "default threads", oracle_c.num_threads())
for nthr in (128, 64, 32, 16):
    oracle_c.set_num_threads(nthr)
    for _ in range(4): op.run(grads)
    t0 = time.perf_counter()
    for _ in range(8): op.run(grads)
    print(nthr, "threads:", round((time.perf_counter() - t0) / 2 * 1e3, 1), "ms/window")
