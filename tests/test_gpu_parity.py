"""CUDA path (through the C ABI) vs the CPU oracle on identical seeded inputs."""
import numpy as np
import pytest

import oracle_np as onp
from common import make_grads, make_params, oracle_for, rel_err

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

BERT_SCHED = dict(init_lr=2e-5, num_train_steps=207900, num_warmup_steps=20790)

TOY = [("bert/embeddings/word_embeddings", (1000, 33)), ("bert/embeddings/LayerNorm/beta", (33,)),
       ("bert/embeddings/LayerNorm/gamma", (33,)), ("l0/attention/self/query/kernel", (64, 64)),
       ("l0/attention/self/query/bias", (64,)), ("big/kernel", (3, 4099)), ("output_bias", (2,)),
       ("one", (1,)), ("l1/dense/kernel", (2048,)), ("l1/dense/bias", (2049,))]


def _gpu_op(manifest, params, hp_kind, N, sched):
    import gaccum_b200 as g
    from gaccum_b200.train_op import GaccumTrainOp
    dev = torch.device("cuda:0")
    tp = [torch.from_numpy(p.copy()).to(dev) for p in params]
    if hp_kind == "bert":
        hp = g.HParams.bert()
    elif hp_kind == "bert_noclip":
        hp = g.HParams.bert(); hp.clip_norm = 0.0
    else:
        hp = g.HParams.tf_adam()
    if "constant_lr" in sched:
        lr_fn = lambda s: sched["constant_lr"]
    else:
        lr_fn = lambda s: g.learning_rate(sched["init_lr"], sched["num_train_steps"], sched["num_warmup_steps"], s)
    return GaccumTrainOp(tp, [n for n, _ in manifest], hp, N, lr_fn), tp


def _oracle_hp(kind):
    if kind == "bert":
        return onp.HParams.bert()
    if kind == "bert_noclip":
        hp = onp.HParams.bert(); hp.clip_norm = 0.0; return hp
    return onp.HParams.tf_adam()


def _compare(op, tp, ref, bitexact, tol=1e-5):
    """accum is ALWAYS bit-identical (one fp32 add per micro-step, zeroed on apply); p, m, v are
    bit-identical until the first apply whose clip scale != 1 (the norm's summation order differs
    from the oracle's), after which they are held to the north-star tolerance."""
    worst = 0.0
    for i in range(len(tp)):
        for name, got, exp in (("p", tp[i].cpu().numpy(), ref.params[i]),
                               ("m", op.m_view(i).cpu().numpy(), ref.m[i]),
                               ("v", op.v_view(i).cpu().numpy(), ref.v[i]),
                               ("a", op.accum_view(i).cpu().numpy(), ref.accum[i])):
            if bitexact or name == "a":
                assert np.array_equal(got, exp, equal_nan=True), f"tensor {i} {name} not bit-identical"
            else:
                e = rel_err(got, exp)
                worst = max(worst, e)
                assert e <= tol, f"tensor {i} {name}: rel err {e}"
                assert np.allclose(got, exp, rtol=tol, atol=1e-8, equal_nan=True)
    return worst


@pytest.mark.parametrize("hp_kind,sigma,sched", [
    ("bert", 1e-4, BERT_SCHED),                      # unclipped: scale == 1.0 -> bit-exact
    ("bert", 1.0, BERT_SCHED),                       # clipped
    ("bert", 1e-2, dict(BERT_SCHED, num_warmup_steps=0)),
    ("bert_noclip", 1e-2, dict(BERT_SCHED, num_warmup_steps=0)),
    ("adam", 1e-2, dict(constant_lr=1e-4)),          # distributedExample/02 optimizer
])
@pytest.mark.parametrize("N", [1, 3, 4])
def test_trajectory_matches_oracle(hp_kind, sigma, sched, N):
    rng = np.random.default_rng(7)
    params = make_params(TOY, rng)
    ref = oracle_for(TOY, params, _oracle_hp(hp_kind), N, **sched)
    op, tp = _gpu_op(TOY, params, hp_kind, N, sched)
    exact = True
    for step in range(2 * N + 2):
        grads = make_grads(TOY, sigma, 0, step)
        info = ref.run(grads)
        applied = op.run([torch.from_numpy(g).cuda() for g in grads])
        assert applied == info.applied
        st = op.stats()
        assert st["applied"] == info.applied and np.float32(st["lr"]) == info.lr
        clipping = hp_kind == "bert"
        if info.applied and clipping:
            assert abs(st["global_norm"] - float(info.global_norm)) <= 2e-6 * float(info.global_norm)
        # bit-exact for as long as every clip scale so far was exactly 1 on both sides
        if info.applied and clipping and not (float(info.clip_scale) == 1.0 and st["clip_scale"] == 1.0):
            exact = False
        _compare(op, tp, ref, bitexact=exact)
    assert op.global_step == ref.global_step


def test_unaligned_and_missing_grads():
    """Views at odd element offsets (4-byte aligned only) and a tensor without gradient."""
    import gaccum_b200 as g
    from gaccum_b200.train_op import GaccumTrainOp
    rng = np.random.default_rng(3)
    man = [("a/kernel", (777,)), ("a/bias", (5,)), ("b/kernel", (4097,)), ("c/kernel", (100,))]
    params = make_params(man, rng)
    flat = torch.zeros(1 + sum(p.size for p in params) + 16, device="cuda")
    tp, o = [], 1                                 # start at element 1 -> pointers are 4 B aligned only
    for p in params:
        t = flat[o:o + p.size]; t.copy_(torch.from_numpy(p.ravel())); tp.append(t); o += p.size
    hp = g.HParams.bert()
    op = GaccumTrainOp(tp, [n for n, _ in man], hp, 2, lambda s: 1e-3)
    ref = oracle_for(man, [p.ravel() for p in params], onp.HParams.bert(), 2, constant_lr=1e-3)
    gflat = torch.zeros_like(flat)
    for step in range(5):
        grads = make_grads(man, 0.5, 0, step)
        grads[3] = None
        tg, o = [], 3                             # different misalignment for grads
        for gr in grads:
            if gr is None:
                tg.append(None); continue
            t = gflat[o:o + gr.size]; t.copy_(torch.from_numpy(gr.ravel())); tg.append(t); o += gr.size
        info = ref.run([None if x is None else x.ravel() for x in grads])
        op.run(tg)
        _compare(op, tp, ref, bitexact=False)


def test_nan_gradient_propagates():
    """TF 1.15 clip_by_global_norm: a non-finite norm makes every update NaN."""
    import gaccum_b200 as g
    from gaccum_b200.train_op import GaccumTrainOp
    man = [("w/kernel", (300,)), ("w/bias", (7,))]
    params = make_params(man, np.random.default_rng(0))
    tp = [torch.from_numpy(p.copy()).cuda() for p in params]
    op = GaccumTrainOp(tp, [n for n, _ in man], g.HParams.bert(), 1, lambda s: 1e-3)
    grads = [torch.zeros(300, device="cuda"), torch.zeros(7, device="cuda")]
    grads[0][5] = float("nan")
    op.run(grads)
    assert all(torch.isnan(t).all() for t in tp)
    assert torch.isnan(op.m_view(1)).all() and (op.accum == 0).all()


def test_split_branches_equal_fused_step():
    """accumulate_only + apply_only(None) (what the data-parallel drivers call) == run() on the apply step."""
    import gaccum_b200 as g
    from gaccum_b200.train_op import GaccumTrainOp
    rng = np.random.default_rng(11)
    params = make_params(TOY, rng)
    names = [n for n, _ in TOY]
    mk = lambda: GaccumTrainOp([torch.from_numpy(p.copy()).cuda() for p in params], names, g.HParams.bert(), 3, lambda s: 2e-3)
    a, b = mk(), mk()
    for step in range(7):
        grads = [torch.from_numpy(x).cuda() for x in make_grads(TOY, 0.7, 0, step)]
        a.run(grads)
        if b.global_step % b.N == 0:
            b.accumulate_only(grads)
            b.apply_only(None)
            b.global_step += 1
        else:
            b.run(grads)
        for x, y in zip(a.params, b.params):
            assert torch.equal(x, y)
        assert torch.equal(a.m, b.m) and torch.equal(a.v, b.v) and torch.equal(a.accum, b.accum)
        assert a.stats() == b.stats() or not a.last_applied


def test_empty_single_element_and_full_pointer_table():
    """Edge shapes: a zero-size tensor, 1-element tensors, and T = 1920 (the largest pointer table)."""
    import gaccum_b200 as g
    from gaccum_b200.train_op import GaccumTrainOp
    rng = np.random.default_rng(12)
    man = [("empty/kernel", (0,)), ("one/kernel", (1,)), ("one/bias", (1,))]
    man += [(f"t{i}/{'bias' if i % 3 == 0 else 'kernel'}", (int(rng.integers(1, 70)),)) for i in range(1917)]
    assert len(man) == 1920
    params = [rng.normal(0, 0.02, s).astype(np.float32) for _, s in man]
    ref = oracle_for(man, params, onp.HParams.bert(), 2, constant_lr=1e-4)
    flat = torch.zeros(sum(p.size for p in params) + 8, device="cuda")
    tp, o = [], 0
    for p in params:                          # views into one flat buffer: arbitrary 4-byte alignment
        t = flat[o:o + p.size]; t.copy_(torch.from_numpy(p)); tp.append(t); o += p.size
    op = GaccumTrainOp(tp, [n for n, _ in man], g.HParams.bert(), 2, lambda s: 1e-4)
    assert op.plan.T == 1920 and op.plan.num_tiles == 1919
    for step in range(3):
        grads = [rng.normal(0, 0.05, s).astype(np.float32) for _, s in man]
        info = ref.run(grads)
        gflat = torch.from_numpy(np.concatenate([x.ravel() for x in grads])).cuda()
        tg, o = [], 0
        for x in grads:
            tg.append(gflat[o:o + x.size]); o += x.size
        assert op.run(tg) == info.applied
        _compare(op, tp, ref, bitexact=False)
    with pytest.raises(g.GaccumError):
        g.Plan([1] * 1921, None, g.HParams.bert(), device=0)


def test_steps_are_cuda_graph_capturable():
    """The ABI promises asynchronous, sync-free launches: a whole window (N-1 accumulates + the
    cooperative apply) must record into a CUDA graph and replay to the same bits as eager launches."""
    import gaccum_b200 as g
    from gaccum_b200.train_op import GaccumTrainOp
    rng = np.random.default_rng(21)
    params = make_params(TOY, rng)
    names = [n for n, _ in TOY]
    mk = lambda: GaccumTrainOp([torch.from_numpy(p.copy()).cuda() for p in params], names, g.HParams.bert(), 3, lambda s: 1e-3,
                               global_step=1)
    eager, graphed = mk(), mk()
    grads = [[torch.from_numpy(x).cuda() for x in make_grads(TOY, 0.3, 0, s)] for s in range(3)]
    tables = [graphed.bind(gl) for gl in grads]
    side = torch.cuda.Stream()
    cg = torch.cuda.CUDAGraph()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for t in tables:                      # warm-up outside capture (lazy occupancy queries, attribute sets)
            graphed.run_bound(t, side.cuda_stream)
        graphed.load_state_dict(eager.state_dict())
        torch.cuda.synchronize()
        with torch.cuda.graph(cg, stream=side):
            gs = graphed.global_step
            for t in tables:                  # steps 1, 2 accumulate; step 3 applies (3 % 3 == 0)
                graphed.run_bound(t, side.cuda_stream)
            graphed.global_step = gs          # the captured launches carry steps 1..3 baked in
    for _ in range(3):
        graphed.global_step = 1
        cg.replay()
        for gl in grads:
            eager.run(gl)
        eager.global_step = 1
        torch.cuda.synchronize()
        for x, y in zip(eager.params, graphed.params):
            assert torch.equal(x, y)
        assert torch.equal(eager.m, graphed.m) and torch.equal(eager.v, graphed.v) and torch.equal(eager.accum, graphed.accum)


def test_clip_apply_is_bit_reproducible_although_tiles_are_scheduled_dynamically():
    """The clip-apply kernel hands tiles to SMs through atomic tickets, so which SM reduces which tile changes from
    launch to launch.  The global norm is accumulated EXACTLY (fixed-point integer atomics), so results must still be
    bit-identical run after run -- the property data-parallel replicas rely on."""
    import gaccum_b200 as g
    from gaccum_b200.train_op import GaccumTrainOp
    man = [("w%d/kernel" % i, (257, 129 + i)) for i in range(40)] + [("emb", (50000, 33)), ("b/bias", (7,)), ("LayerNorm/gamma", (3,))]
    rng = np.random.default_rng(3)
    p0 = [rng.normal(0, 0.02, s).astype(np.float32) for _, s in man]
    grads = [[torch.from_numpy(rng.normal(0, 0.05, s).astype(np.float32)).cuda() for _, s in man] for _ in range(3)]
    outs = []
    for rep in range(4):
        tp = [torch.from_numpy(p.copy()).cuda() for p in p0]
        op = GaccumTrainOp(tp, [n for n, _ in man], g.HParams.bert(), 2, lambda s: 1e-3, global_step=1)
        stats = []
        for i in range(3):                      # accumulate, apply (clipped), accumulate
            op.run(grads[i])
            stats.append(op.stats())
        assert stats[1]["applied"] and stats[1]["clip_scale"] < 1.0
        outs.append(([t.cpu().numpy() for t in tp], op.m.cpu().numpy(), op.v.cpu().numpy(), stats[1]))
    for o in outs[1:]:
        assert o[3] == outs[0][3]
        assert all(np.array_equal(a, b) for a, b in zip(o[0], outs[0][0])) and np.array_equal(o[1], outs[0][1]) and np.array_equal(o[2], outs[0][2])
