"""world_size-2 CPU (gloo) test of the data-parallel wiring (reference distributedExample/04).

The kernels cannot run here, so the per-rank compute engine is the CPU oracle behind the same
engine interface GaccumTrainOp exposes; what is under test is the orchestration in
gaccum_b200/distributed.py: local accumulation, ONE all-reduce of the packed slab per window (on the
apply step only), identical replicas, and equality with a single-process run fed the summed gradients.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAN = [("l0/kernel", (33, 7)), ("l0/bias", (7,)), ("LayerNorm/gamma", (7,)), ("emb", (301,))]
N, STEPS, WORLD = 3, 8, 2


class OracleEngine:
    """Test-only adapter: oracle_np.ReferenceTrainOp with its accumulators aliased to one flat slab."""

    def __init__(self, ref):
        self.ref = ref
        flat = np.zeros(sum(a.size for a in ref.accum), np.float32)
        o = 0
        for i, a in enumerate(ref.accum):
            ref.accum[i] = flat[o:o + a.size].reshape(a.shape); o += a.size
        self.accum = torch.from_numpy(flat)
        self.N = ref.N
        self.names = list(ref.names)                      # what tf_checkpoint.save / restore ask the engine for
        self.hp = type("HP", (), {"variant": 0})()

    @property
    def global_step(self):
        return self.ref.global_step

    @global_step.setter
    def global_step(self, v):
        self.ref.global_step = v

    def run(self, grads):
        return self.ref.run(grads).applied

    def accumulate_only(self, grads):
        for a, g in zip(self.ref.accum, grads):
            if g is not None:
                np.add(a, g, out=a)

    def apply_only(self, grads):
        assert grads is None and self.ref.global_step % self.N == 0
        self.ref.run([None] * len(self.ref.accum))

    # checkpoint surface of GaccumTrainOp (reference names; accumulators read from whatever `accum` currently is)
    def state_dict(self):
        out = {"global_step": torch.tensor(self.ref.global_step)}
        flat = self.accum.numpy()
        o = 0
        for i, n in enumerate(self.ref.names):
            sz = self.ref.params[i].size
            out[n] = torch.from_numpy(self.ref.params[i].copy())
            out[n + "/adam_m"] = torch.from_numpy(self.ref.m[i].copy())
            out[n + "/adam_v"] = torch.from_numpy(self.ref.v[i].copy())
            out[n + "/accum_grad"] = torch.from_numpy(flat[o:o + sz].reshape(self.ref.params[i].shape).copy())
            o += sz
        return out

    def load_state_dict(self, sd, strict=True):
        self.ref.global_step = int(sd["global_step"])
        for i, n in enumerate(self.ref.names):
            self.ref.params[i][...] = sd[n].numpy(); self.ref.m[i][...] = sd[n + "/adam_m"].numpy()
            self.ref.v[i][...] = sd[n + "/adam_v"].numpy(); self.ref.accum[i][...] = sd[n + "/accum_grad"].numpy()


def _grads(rank, step):
    rng = np.random.Generator(np.random.PCG64(19830610 + 1000 * rank + step))
    return [(rng.normal(0, 0.3, s) / WORLD).astype(np.float32) for _, s in MAN]      # 04:46 loss / num_workers


def _worker(rank, port, outdir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import oracle_np as onp
    from gaccum_b200.distributed import DataParallelTrainOp
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=WORLD)
    rng = np.random.default_rng(5)
    params = [rng.normal(0, 0.02, s).astype(np.float32) for _, s in MAN]
    ref = onp.ReferenceTrainOp(params, [n for n, _ in MAN], onp.HParams.bert(), N, constant_lr=1e-2)
    dp = DataParallelTrainOp(OracleEngine(ref), None)
    applied = [dp.run(_grads(rank, s)) for s in range(STEPS)]
    np.savez(os.path.join(outdir, f"rank{rank}.npz"), *ref.params, *ref.m, *ref.v, *ref.accum,
             allreduces=dp.allreduces, applied=np.array(applied), gs=ref.global_step)
    dist.destroy_process_group()


def test_dp_world2_matches_single_process_on_summed_grads(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_worker, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    import oracle_np as onp
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    arr = [k for k in r0.files if k.startswith("arr_")]
    for k in arr:                                   # replicas are bit-identical (params, m, v)
        if int(k[4:]) < 3 * len(MAN):
            assert np.array_equal(r0[k], r1[k]), k
    n_apply = int(r0["applied"].sum())
    assert list(r0["applied"]) == [s % N == 0 for s in range(STEPS)] and r0["gs"] == STEPS
    assert int(r0["allreduces"]) == n_apply == 3    # ONE exchange per window, none on accumulate steps
    # single-process oracle fed the rank-summed gradient
    rng = np.random.default_rng(5)
    params = [rng.normal(0, 0.02, s).astype(np.float32) for _, s in MAN]
    ref = onp.ReferenceTrainOp(params, [n for n, _ in MAN], onp.HParams.bert(), N, constant_lr=1e-2)
    for s in range(STEPS):
        ref.run([a + b for a, b in zip(_grads(0, s), _grads(1, s))])
    T = len(MAN)
    for i in range(T):
        np.testing.assert_allclose(r0[f"arr_{i}"], ref.params[i], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(r0[f"arr_{T + i}"], ref.m[i], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(r0[f"arr_{2 * T + i}"], ref.v[i], rtol=1e-5, atol=1e-9)
    # accumulators are rank-local between applies: after step 7 (two accumulate steps) they differ per rank
    assert not np.array_equal(r0[f"arr_{3 * T}"], r1[f"arr_{3 * T}"])


def _worker_resume(rank, port, outdir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p)
    import oracle_np as onp
    from gaccum_b200.distributed import DataParallelTrainOp
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=WORLD)
    names = [n for n, _ in MAN]

    def fresh(fill=None):
        rng = np.random.default_rng(5)
        params = [rng.normal(0, 0.02, s).astype(np.float32) if fill is None else np.full(s, fill, np.float32) for _, s in MAN]
        return onp.ReferenceTrainOp(params, names, onp.HParams.bert(), N, constant_lr=1e-2)
    ref = fresh()
    dp = DataParallelTrainOp(OracleEngine(ref), None)
    for s in range(5):                                   # N = 3: step 3 applied, step 4 only accumulated -> mid-window
        dp.run(_grads(rank, s))
    sd = dp.state_dict()                                 # collective: accumulators summed over ranks (04:55 aggregation=SUM)
    local = float(sum(np.abs(a).sum() for a in ref.accum))
    summed = float(sum(float(sd[n + "/accum_grad"].abs().sum()) for n in names))
    # the same state as a TensorFlow-format checkpoint on disk: every rank calls save(), rank 0 alone writes
    from gaccum_b200 import tf_checkpoint as ck
    model_dir = os.path.join(outdir, "model_dir")
    prefix = ck.save(model_dir, dp)
    assert os.path.exists(prefix + ".index") and sorted(os.listdir(model_dir)) == ["checkpoint", "model.ckpt-5.data-00000-of-00001", "model.ckpt-5.index"]
    ref2 = fresh(fill=9.0)                               # a restarted worker holds garbage until it restores
    dp2 = DataParallelTrainOp(OracleEngine(ref2), None)
    if rank == 0:
        dp2.load_state_dict(sd)                          # rank 0 from the in-memory dictionary, rank 1 from the file
    else:
        assert ck.restore(model_dir, dp2) == prefix
    for s in range(5, STEPS):
        dp2.run(_grads(rank, s))
    np.savez(os.path.join(outdir, f"res{rank}.npz"), *ref2.params, local=local, summed=summed, gs=ref2.global_step)
    dist.destroy_process_group()


def test_dp_checkpoint_mid_window_sums_accumulators_and_resumes(tmp_path):
    """Checkpoint compatibility under data parallelism: the saved accumulators are the SUM over ranks
    (what the reference's aggregation=SUM variables hold), restored into rank 0 only, and training continues exactly."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_worker_resume, args=(port, str(tmp_path)), nprocs=WORLD, join=True)
    import oracle_np as onp
    r0, r1 = np.load(tmp_path / "res0.npz"), np.load(tmp_path / "res1.npz")
    assert float(r0["summed"]) == float(r1["summed"]) and float(r0["summed"]) > max(float(r0["local"]), float(r1["local"]))
    rng = np.random.default_rng(5)
    params = [rng.normal(0, 0.02, s).astype(np.float32) for _, s in MAN]
    ref = onp.ReferenceTrainOp(params, [n for n, _ in MAN], onp.HParams.bert(), N, constant_lr=1e-2)
    for s in range(STEPS):
        ref.run([a + b for a, b in zip(_grads(0, s), _grads(1, s))])
    assert int(r0["gs"]) == STEPS
    for i in range(len(MAN)):
        assert np.array_equal(r0[f"arr_{i}"], r1[f"arr_{i}"])
        np.testing.assert_allclose(r0[f"arr_{i}"], ref.params[i], rtol=1e-5, atol=1e-7)
