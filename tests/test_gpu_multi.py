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


def _worker_fused(rank, world, port, outdir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import gaccum_b200 as g
    from gaccum_b200.distributed import FusedDataParallelTrainOp
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    rng = np.random.default_rng(5)
    params = [torch.from_numpy(rng.normal(0, 0.02, s).astype(np.float32)).cuda() for _, s in MAN]
    dp = FusedDataParallelTrainOp(params, [n for n, _ in MAN], g.HParams.bert(), N, lambda s: 1e-2)
    for s in range(STEPS):
        dp.run([torch.from_numpy(x).cuda() for x in _grads(rank, s, world)])
    torch.cuda.synchronize()
    st = dp.gather_state()
    np.savez(os.path.join(outdir, f"frank{rank}.npz"), *[p.cpu().numpy() for p in params],
             m=st["m"].cpu().numpy(), v=st["v"].cpu().numpy(), accum=dp.engine.accum.cpu().numpy(),
             exchanges=dp.exchanges, owned=dp.owned_elements, stats=np.array(list(dp.engine.stats().values()), dtype=np.float64))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_dp_kernel_two_gpus(tmp_path):
    """The one-kernel exchange (peer loads / peer stores) gives every rank the same parameters as the
    single-process oracle fed the rank-summed gradient."""
    import torch.multiprocessing as mp
    import oracle_np as onp
    ndev = torch.cuda.device_count()
    world = 8 if ndev >= 8 else 4 if ndev >= 4 else 2       # W=2 -> 4 tiles/iter, W=4 -> 2, W=8 -> 1
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_worker_fused, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rs = [np.load(tmp_path / f"frank{r}.npz") for r in range(world)]
    r0 = rs[0]
    T = len(MAN)
    for r in rs[1:]:
        for i in range(T):
            assert np.array_equal(r0[f"arr_{i}"], r[f"arr_{i}"]), f"replicas differ in tensor {i}"
        assert np.array_equal(r0["m"], r["m"]) and np.array_equal(r0["v"], r["v"])
        assert np.array_equal(r0["stats"], r["stats"])                   # identical gn / clip scale
    assert int(r0["exchanges"]) == 4
    assert sum(int(r["owned"]) for r in rs) == sum(int(np.prod(s)) for _, s in MAN)
    rng = np.random.default_rng(5)
    params = [rng.normal(0, 0.02, s).astype(np.float32) for _, s in MAN]
    ref = onp.ReferenceTrainOp(params, [n for n, _ in MAN], onp.HParams.bert(), N, constant_lr=1e-2)
    for s in range(STEPS):
        gs = [_grads(r, s, world) for r in range(world)]
        ref.run([np.sum([g[i] for g in gs], axis=0, dtype=np.float32) for i in range(T)])
    for i in range(T):
        assert np.allclose(r0[f"arr_{i}"], ref.params[i], rtol=1e-5, atol=1e-7)
    assert all(not r["accum"].any() for r in rs)                         # STEPS-1 is an apply step: all zero


def _worker_host_dp(rank, world, port, outdir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import gaccum_b200 as g
    from gaccum_b200.train_op import HostTrainOp
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    rng = np.random.default_rng(5)
    hp = g.HParams.bert()
    _, host_params = HostTrainOp.pinned_arena([s for _, s in MAN], hp)
    for t, (_, s) in zip(host_params, MAN):
        t.copy_(torch.from_numpy(rng.normal(0, 0.02, s).astype(np.float32)))
    op = HostTrainOp(host_params, [n for n, _ in MAN], hp, N, lambda s: 1e-2, device=rank)
    op.connect_data_parallel()
    for s in range(STEPS):
        _, gv = HostTrainOp.pinned_arena([sh for _, sh in MAN], hp)
        for t, x in zip(gv, _grads(rank, s, world)):
            t.copy_(torch.from_numpy(x))
        op.run(gv)
        op.sync()
    st = op.stats()
    np.savez(os.path.join(outdir, f"hrank{rank}.npz"), *[p.numpy().copy() for p in host_params],
             stats=np.array([st["global_norm"], st["clip_scale"]], dtype=np.float64))
    dist.barrier()
    del op
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_host_buffer_data_parallel_over_cuda_ipc(tmp_path):
    """gaccum_step_host + gaccum_host_session_dp_connect: host-resident tensors, the apply step is the fused
    NVLink exchange + apply kernel over CUDA-IPC mappings (reference 04:46,55,58,62 with CPU tensors)."""
    import torch.multiprocessing as mp
    import oracle_np as onp
    ndev = torch.cuda.device_count()
    world = 4 if ndev >= 4 else 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_worker_host_dp, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rs = [np.load(tmp_path / f"hrank{r}.npz") for r in range(world)]
    T = len(MAN)
    for r in rs[1:]:
        for i in range(T):
            assert np.array_equal(rs[0][f"arr_{i}"], r[f"arr_{i}"]), f"replicas differ in tensor {i}"
        assert np.array_equal(rs[0]["stats"], r["stats"])
    rng = np.random.default_rng(5)
    params = [rng.normal(0, 0.02, s).astype(np.float32) for _, s in MAN]
    ref = onp.ReferenceTrainOp(params, [n for n, _ in MAN], onp.HParams.bert(), N, constant_lr=1e-2)
    for s in range(STEPS):
        gs = [_grads(r, s, world) for r in range(world)]
        ref.run([np.sum([g[i] for g in gs], axis=0, dtype=np.float32) for i in range(T)])
    for i in range(T):
        assert np.allclose(rs[0][f"arr_{i}"], ref.params[i], rtol=1e-5, atol=1e-7)


def _worker_fused_resume(rank, world, port, outdir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    import gaccum_b200 as g
    from gaccum_b200.distributed import FusedDataParallelTrainOp
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    rng = np.random.default_rng(5)
    names = [n for n, _ in MAN]
    params = [torch.from_numpy(rng.normal(0, 0.02, s).astype(np.float32)).cuda() for _, s in MAN]
    dp = FusedDataParallelTrainOp(params, names, g.HParams.bert(), N, lambda s: 1e-2)
    for s in range(4):                                            # stop MID-WINDOW: step 3 only accumulated
        dp.run([torch.from_numpy(x).cuda() for x in _grads(rank, s, world)])
    sd = dp.state_dict()
    assert any(k.endswith("/adam_m") for k in sd) and float(sd["emb/accum_grad"].abs().sum()) > 0
    # ... and through a TensorFlow-format checkpoint on disk under the reference's Saver names: rank 0 writes, all restore
    from gaccum_b200 import tf_checkpoint as ck
    prefix = ck.save(os.path.join(outdir, "model_dir"), dp)
    assert os.path.exists(prefix + ".index") and os.path.basename(prefix) == "model.ckpt-4"
    del dp
    fresh = [torch.full(s, 7.0, device="cuda") for _, s in MAN]   # a new process would start from garbage
    dp2 = FusedDataParallelTrainOp(fresh, names, g.HParams.bert(), N, lambda s: 1e-2)
    if rank == 0:
        dp2.load_state_dict(sd)                                   # rank 0 from the in-memory dictionary,
    else:
        assert ck.restore(os.path.join(outdir, "model_dir"), dp2) == prefix   # rank 1 from the file: they must agree
    assert dp2.global_step == 4
    for s in range(4, STEPS):
        dp2.run([torch.from_numpy(x).cuda() for x in _grads(rank, s, world)])
    torch.cuda.synchronize()
    np.savez(os.path.join(outdir, f"rrank{rank}.npz"), *[p.cpu().numpy() for p in fresh])
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_fused_dp_checkpoint_resume_mid_window(tmp_path):
    """A checkpoint taken mid-window under data parallelism (full moments gathered, accumulators summed over ranks,
    reference names) restores into fresh replicas and continues exactly like the uninterrupted run."""
    import torch.multiprocessing as mp
    import oracle_np as onp
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_worker_fused_resume, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "rrank0.npz"), np.load(tmp_path / "rrank1.npz")
    T = len(MAN)
    rng = np.random.default_rng(5)
    params = [rng.normal(0, 0.02, s).astype(np.float32) for _, s in MAN]
    ref = onp.ReferenceTrainOp(params, [n for n, _ in MAN], onp.HParams.bert(), N, constant_lr=1e-2)
    for s in range(STEPS):
        ref.run([a + b for a, b in zip(_grads(0, s, world), _grads(1, s, world))])
    for i in range(T):
        assert np.array_equal(r0[f"arr_{i}"], r1[f"arr_{i}"])
        assert np.allclose(r0[f"arr_{i}"], ref.params[i], rtol=1e-5, atol=1e-7)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_example_04_recipe_under_torchrun():
    """distributedExample/04 (MultiWorkerMirroredStrategy recipe: 04:13-15, 46, 55-62, 98-121) as a launch script over
    the fused data-parallel train_op: `torchrun --nproc-per-node 2 examples/mnist_gaccum.py` runs and its loss falls."""
    import re
    import subprocess
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(ROOT, "examples", "mnist_gaccum.py"), "--steps", "300", "--lr", "1e-3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    losses = [float(x) for x in re.findall(r"loss ([0-9.]+)", out.stdout)]
    assert len(losses) >= 3 and losses[-1] < 0.5 * losses[0], out.stdout
