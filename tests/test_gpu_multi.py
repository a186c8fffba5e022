"""2-GPU NCCL test of the data-parallel train_op (skipped on a 1-GPU box)."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAN = [("l0/kernel", (257, 33)), ("l0/bias", (33,)), ("LayerNorm/gamma", (33,)), ("emb", (70001,))]
N, STEPS = 2, 7


def _grads(rank, step, world):
    rng = np.random.Generator(np.random.PCG64(19830610 + 1000 * rank + step))
    return [(rng.normal(0, 0.3, s) / world).astype(np.float32) for _, s in MAN]


def _worker(rank, world, port, outdir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import gaccum_b200 as g
    from gaccum_b200.distributed import DataParallelTrainOp
    from gaccum_b200.train_op import GaccumTrainOp
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    rng = np.random.default_rng(5)
    params = [torch.from_numpy(rng.normal(0, 0.02, s).astype(np.float32)).cuda() for _, s in MAN]
    op = GaccumTrainOp(params, [n for n, _ in MAN], g.HParams.bert(), N, lambda s: 1e-2)
    dp = DataParallelTrainOp(op, None)
    for s in range(STEPS):
        dp.run([torch.from_numpy(x).cuda() for x in _grads(rank, s, world)])
    torch.cuda.synchronize()
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), *[p.cpu().numpy() for p in params],
             m=op.m.cpu().numpy(), v=op.v.cpu().numpy(), allreduces=dp.allreduces)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dp_two_gpus_identical_replicas_and_oracle_parity(tmp_path):
    import torch.multiprocessing as mp
    import oracle_np as onp
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for k in r0.files:
        assert np.array_equal(r0[k], r1[k]), f"replicas differ in {k}"
    assert int(r0["allreduces"]) == 4
    rng = np.random.default_rng(5)
    params = [rng.normal(0, 0.02, s).astype(np.float32) for _, s in MAN]
    ref = onp.ReferenceTrainOp(params, [n for n, _ in MAN], onp.HParams.bert(), N, constant_lr=1e-2)
    for s in range(STEPS):
        ref.run([a + b for a, b in zip(_grads(0, s, world), _grads(1, s, world))])
    for i in range(len(MAN)):
        assert np.allclose(r0[f"arr_{i}"], ref.params[i], rtol=1e-5, atol=1e-7)
