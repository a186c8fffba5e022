"""Where does the host-buffer step spend its time?  T=73 (BERT-Small) vs the same bytes in 1 tensor."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, gaccum_b200 as g
from gaccum_b200.manifests import MANIFESTS
from gaccum_b200.train_op import HostTrainOp
man73 = MANIFESTS["bert_small"]()
P = sum(int(torch.tensor(s).prod()) for _, s in man73)
for label, man in (("T=73", man73), ("T=1", [("all/kernel", (P,))])):
    for with_params in (True, False):
        hp = [torch.zeros(s).pin_memory() for _, s in man]
        hg = [[torch.zeros(s).pin_memory() for _, s in man] for _ in range(2)]
        op = HostTrainOp(hp, [n for n, _ in man], g.HParams.bert(), 4, lambda s: 1e-5, global_step=1)
        if not with_params:
            op._param_ptrs = None     # no D2H of parameters on apply steps
        b = [op.bind(x) for x in hg]
        for i in range(4): op.run_bound(b[i % 2])
        op.sync()
        t0 = time.perf_counter()
        K = 48
        for i in range(K): op.run_bound(b[i % 2])
        op.sync()
        ms = (time.perf_counter() - t0) * 1e3 / K
        print(f"{label} params_d2h={with_params}: {ms:.3f} ms/step  -> H2D {P*4/ms/1e6:.1f} GB/s")
        del op
