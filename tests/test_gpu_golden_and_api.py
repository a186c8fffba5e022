"""GPU: the CUDA path against the committed golden fixtures (reference's own optimization.py run),
the drop-in Python API (create_optimizer & co.), the packed and host-buffer entry points."""
import numpy as np
import pytest

import oracle_np as onp
from common import make_grads, make_params, oracle_for, rel_err
from golden_util import Golden, cases

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", cases())
def test_cuda_path_reproduces_reference_fixtures(case):
    import gaccum_b200 as g
    from gaccum_b200.train_op import GaccumTrainOp
    gd = Golden(case)
    tp = [torch.from_numpy(p).cuda() for p in gd.init()]
    op = GaccumTrainOp(tp, gd.names, g.HParams.bert(), gd.N,
                       lambda s: g.learning_rate(gd.init_lr, gd.num_train_steps, gd.num_warmup_steps, s))
    exact = True
    for s in range(gd.steps):
        applied = op.run([torch.from_numpy(x).cuda() for x in gd.grads(s)])
        if applied and op.stats()["clip_scale"] != 1.0:
            exact = False                      # norm summation order differs from the fixture's
        assert op.global_step == gd.global_step(s)
        for kind, view in (("param", lambda i: tp[i]), ("accum", op.accum_view), ("m", op.m_view), ("v", op.v_view)):
            for i, exp in enumerate(gd.state(s, kind)):
                got = view(i).cpu().numpy()
                if exact or kind == "accum":
                    assert np.array_equal(got, exp), f"{case} step {s} {kind} {gd.names[i]}"
                else:
                    assert np.allclose(got, exp, rtol=1e-5, atol=1e-8), f"{case} step {s} {kind} {gd.names[i]}"
    if case == "warmup_unclipped":
        assert exact                           # the whole trajectory was bit-identical


def test_cuda_path_reproduces_recipe_fixtures():
    """Variant B (a14) against the AST-lifted recipes of 02 / 04 / another-example executed over the stub:
    tf.train.AdamOptimizer inside the accumulation window, N from params, no clip.  No reduction is involved
    (no global norm), every op is one correctly rounded fp32 op: the CUDA path must be bit-identical."""
    import gaccum_b200 as g
    from gaccum_b200.train_op import GaccumTrainOp
    from golden_util import RecipeGolden, recipe_cases
    assert len(recipe_cases()) >= 4
    for case in recipe_cases():
        gd = RecipeGolden(case)
        tp = [torch.from_numpy(p).cuda() for p in gd.init()]
        op = GaccumTrainOp(tp, gd.names, g.HParams.tf_adam(), gd.N, lambda s, _lr=gd.lr: _lr)
        for s in range(gd.steps):
            applied = op.run([torch.from_numpy(x).cuda() for x in gd.grads(s)])
            assert applied == (s % gd.N == 0) and op.global_step == int(gd.z[f"global_step/{s}"])
            if s in gd.recorded:
                for i, n in enumerate(gd.names):
                    gd.check(f"param/{s}/{n}", tp[i].cpu().numpy()); gd.check(f"accum/{s}/{n}", op.accum_view(i).cpu().numpy())
                    gd.check(f"m/{s}/{n}", op.m_view(i).cpu().numpy()); gd.check(f"v/{s}/{n}", op.v_view(i).cpu().numpy())
                assert np.float32(op.beta1_power) == gd.z[f"beta1_power/{s}"] and np.float32(op.beta2_power) == gd.z[f"beta2_power/{s}"]


def test_direct_apply_gradients_skips_none_pairs_like_the_reference():
    """optimization.AdamWeightDecayOptimizer(...).apply_gradients(zip(grads, tvars)) with one grad None:
    the reference skips the pair (optimization.py:132-133): no slots, no weight decay, parameter untouched."""
    from gaccum_b200 import graph, optimization as opt
    from golden_util import DirectApplyGolden
    gd = DirectApplyGolden()
    graph.reset_default_graph()
    tvars = [graph.add_variable(n, torch.from_numpy(gd.z[f"init/{n}"].copy()).cuda()) for n in gd.names]
    optimizer = opt.AdamWeightDecayOptimizer(learning_rate=gd.lr, weight_decay_rate=0.01, beta_1=0.9, beta_2=0.999, epsilon=1e-6,
                                             exclude_from_weight_decay=["LayerNorm", "layer_norm", "bias"])
    for s in range(gd.steps):
        grads = [None if i == gd.none_at else torch.from_numpy(gd.z[f"grad/{s}/{n}"]).cuda() for i, n in enumerate(gd.names)]
        optimizer.apply_gradients(zip(grads, tvars))
        for i, n in enumerate(gd.names):
            assert np.array_equal(tvars[i].tensor.cpu().numpy(), gd.z[f"param/{s}/{n}"]), f"step {s} {n}"


def test_host_session_coalesces_arena_copies_and_matches_scattered_buffers():
    """gaccum_step_host with gradients / parameters in a pinned arena laid out like the device slabs (one copy per
    direction) gives exactly what separately allocated host tensors give (one copy per tensor)."""
    import gaccum_b200 as g
    from gaccum_b200.train_op import HostTrainOp
    man = [("l0/kernel", (257, 33)), ("l0/bias", (33,)), ("LayerNorm/gamma", (33,)), ("emb", (70001,)), ("tail", (5,))]
    names, shapes = [n for n, _ in man], [s for _, s in man]
    rng = np.random.default_rng(11)
    p0 = [rng.normal(0, 0.02, s).astype(np.float32) for s in shapes]
    gl = [[rng.normal(0, 0.3, s).astype(np.float32) for s in shapes] for _ in range(5)]
    hp = g.HParams.bert()
    results = []
    for arena in (False, True):
        if arena:
            _, params = HostTrainOp.pinned_arena(shapes, hp)
            for t, x in zip(params, p0):
                t.copy_(torch.from_numpy(x))
        else:
            params = [torch.from_numpy(x.copy()).pin_memory() for x in p0]
        op = HostTrainOp(params, names, hp, 2, lambda s: 1e-2)
        for gs in gl:
            if arena:
                _, gv = HostTrainOp.pinned_arena(shapes, hp)
                for t, x in zip(gv, gs):
                    t.copy_(torch.from_numpy(x))
            else:
                gv = [torch.from_numpy(x.copy()).pin_memory() for x in gs]
            op.run(gv)
            op.sync()
        results.append([p.numpy().copy() for p in params])
    assert all(np.array_equal(a, b) for a, b in zip(*results))
    assert not all(np.array_equal(a, b) for a, b in zip(results[0], p0))


def test_packed_inplace_handoff_matches_the_scattered_path_bit_for_bit():
    """p.grad views of the packed accumulator (autograd accumulates in place, no accumulate launch,
    apply without a gradient stream) give exactly what the scattered pointer-table path gives on the same micro-batches,
    and both follow the oracle."""
    import gaccum_b200 as g
    from gaccum_b200.train_op import GaccumTrainOp, PackedTrainOp
    torch.manual_seed(1)
    names = ["dense/kernel", "dense/bias", "LayerNorm/gamma", "LayerNorm/beta", "output_weights", "output_bias"]

    def make():
        torch.manual_seed(1)
        return torch.nn.Sequential(torch.nn.Linear(24, 40), torch.nn.LayerNorm(40), torch.nn.Tanh(), torch.nn.Linear(40, 3)).cuda()
    xs = [torch.randn(16, 24, device="cuda") for _ in range(9)]
    ys = [torch.randint(0, 3, (16,), device="cuda") for _ in range(9)]
    N, lr_fn = 4, (lambda s: 5e-3)
    ma, mb = make(), make()
    pa, pb = list(ma.parameters()), list(mb.parameters())
    ref = onp.ReferenceTrainOp([p.detach().cpu().numpy() for p in pa], names, onp.HParams.bert(), N, constant_lr=5e-3)
    scattered = GaccumTrainOp(pa, names, g.HParams.bert(), N, lr_fn)
    packed = PackedTrainOp(pb, names, g.HParams.bert(), N, lr_fn)
    for i in range(9):
        la = torch.nn.functional.cross_entropy(ma(xs[i]), ys[i]) * 30.0
        ga = torch.autograd.grad(la, pa)
        ref.run([x.cpu().numpy() for x in ga])
        scattered.run([x.contiguous() for x in ga])
        lb = torch.nn.functional.cross_entropy(mb(xs[i]), ys[i]) * 30.0
        lb.backward()                                       # accumulates into the packed slab
        assert packed.step() == (i % N == 0)
        for k in range(len(names)):
            assert torch.equal(pa[k], pb[k]), f"step {i} {names[k]}"
            assert torch.equal(scattered.accum_view(k), packed.accum_view(k)) and torch.equal(scattered.m_view(k), packed.m_view(k))
            assert np.allclose(pb[k].detach().cpu().numpy(), ref.params[k], rtol=1e-5, atol=1e-8)
    assert packed.launches == 3 and packed.stats()["clip_scale"] < 1.0      # 9 micro-steps, 3 applies, nothing else launched


def _tiny_model():
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.LayerNorm(32), torch.nn.Tanh(), torch.nn.Linear(32, 2)).cuda()
    names = {"0.weight": "dense/kernel", "0.bias": "dense/bias", "1.weight": "LayerNorm/gamma",
             "1.bias": "LayerNorm/beta", "3.weight": "output_weights", "3.bias": "output_bias"}
    return m, (lambda n: names[n])


def test_create_optimizer_drop_in_matches_oracle():
    """create_optimizer(loss, init_lr, num_train_steps, num_warmup_steps, use_tpu) -> train_op.run()"""
    from gaccum_b200 import graph, optimization as opt
    graph.reset_default_graph()
    model, rename = _tiny_model()
    tvars = graph.register_module(model, rename)
    xs = [torch.randn(8, 16, device="cuda") for _ in range(20)]
    ys = [torch.randint(0, 2, (8,), device="cuda") for _ in range(20)]
    it = {"i": 0}

    def loss():
        i = it["i"]; it["i"] += 1
        return torch.nn.functional.cross_entropy(model(xs[i]), ys[i])

    train_op = opt.create_optimizer(loss, 1e-3, 50, 4, False)
    assert train_op.accum_n == 8 and train_op.engine.hp.clip_norm == 1.0
    ref = onp.ReferenceTrainOp([v.tensor.detach().cpu().numpy() for v in tvars], [v.name for v in tvars],
                               onp.HParams.bert(), 8, init_lr=1e-3, num_train_steps=50, num_warmup_steps=4)
    assert train_op.engine.decay == ref.decay == [True, False, False, False, True, False]
    for step in range(18):
        grads = train_op.gradients()
        ref.run([g.cpu().numpy() for g in grads])
        train_op.run_with_grads(grads)
        assert int(graph.get_global_step()) == ref.global_step == step + 1
        for v, exp in zip(tvars, ref.params):
            assert np.allclose(v.tensor.detach().cpu().numpy(), exp, rtol=1e-5, atol=1e-8)
    # and the all-in-one call
    l = train_op.run()
    assert l is not None and torch.isfinite(l) and int(graph.get_global_step()) == 19


def test_state_dict_uses_the_reference_slot_names_through_the_public_api():
    """graph.Variable appends ':0' like TF; the reference names slots after _get_variable_name(param.name)
    (optimization.py:135-148, 189-194): 'scope/kernel/adam_m', never 'scope/kernel:0/adam_m'."""
    from gaccum_b200 import graph, optimization as opt
    graph.reset_default_graph()
    model, rename = _tiny_model()
    tvars = graph.register_module(model, rename)
    assert all(v.name.endswith(":0") for v in tvars)
    train_op = opt.create_optimizer(lambda: model(torch.randn(4, 16, device="cuda")).sum(), 1e-3, 50, 4, False)
    train_op.run()
    keys = set(train_op.state_dict())
    assert {"dense/kernel", "dense/kernel/adam_m", "dense/kernel/adam_v", "dense/kernel/accum_grad", "global_step"} <= keys
    assert not any(":0" in k for k in keys)


def test_inline_recipe_with_tf_adam_like_example_02():
    """distributedExample/02:47-73 -- AdamOptimizer(1e-4), N from params, no clip."""
    from gaccum_b200 import graph, optimization as opt
    graph.reset_default_graph()
    man = onp.MANIFESTS["mnist_cnn"]()
    params = make_params(man, np.random.default_rng(1))
    tv = [graph.add_variable(n, torch.from_numpy(p.copy()).cuda()) for (n, _), p in zip(man, params)]
    train_op = opt.gradient_accumulation_train_op(None, opt.AdamOptimizer(learning_rate=1e-4), 2)
    ref = oracle_for(man, params, onp.HParams.tf_adam(), 2, constant_lr=1e-4)
    for step in range(7):
        grads = make_grads(man, 0.1, 0, step)
        ref.run(grads)
        train_op.run_with_grads([torch.from_numpy(x).cuda() for x in grads])
        for v, exp in zip(tv, ref.params):
            assert np.array_equal(v.tensor.cpu().numpy(), exp)           # no clip -> bit-exact
    assert np.float32(train_op.engine.beta1_power) == ref.beta1_power
    assert np.float32(train_op.engine.beta2_power) == ref.beta2_power


def test_optimizer_apply_gradients_direct():
    """AdamWeightDecayOptimizer.apply_gradients(zip(grads, tvars)) -- optimization.py:128-177."""
    from gaccum_b200 import graph, optimization as opt
    graph.reset_default_graph()
    man = [("w/kernel", (300,)), ("w/bias", (9,)), ("skip/kernel", (5,))]
    params = make_params(man, np.random.default_rng(2))
    tv = [graph.add_variable(n, torch.from_numpy(p.copy()).cuda()) for (n, _), p in zip(man, params)]
    o = opt.AdamWeightDecayOptimizer(learning_rate=0.05, weight_decay_rate=0.01, beta_1=0.9, beta_2=0.999,
                                     epsilon=1e-6, exclude_from_weight_decay=["LayerNorm", "layer_norm", "bias"])
    p, m, v = [x.copy() for x in params], [np.zeros_like(x) for x in params], [np.zeros_like(x) for x in params]
    for step in range(3):
        grads = make_grads(man, 0.2, 0, step)
        o.apply_gradients(zip([torch.from_numpy(grads[0]).cuda(), torch.from_numpy(grads[1]).cuda(), None], tv))
        for i in (0, 1):
            p[i], m[i], v[i] = onp.adam_weight_decay_update(p[i], m[i], v[i], grads[i], 0.05, 0.9, 0.999, 1e-6, 0.01, i == 0)
        for i in range(3):
            assert np.array_equal(tv[i].tensor.cpu().numpy(), p[i])        # tensor 2 (grad None) untouched


def test_packed_entry_point_matches_table_entry_point():
    import gaccum_b200 as g
    from gaccum_b200.train_op import GaccumTrainOp
    man = [("a/kernel", (5000,)), ("a/bias", (3,)), ("LayerNorm/gamma", (70,)), ("b/kernel", (2049,))]
    params = make_params(man, np.random.default_rng(4))
    hp = g.HParams.bert()
    op1 = GaccumTrainOp([torch.from_numpy(p.copy()).cuda() for p in params], [n for n, _ in man], hp, 2, lambda s: 1e-2)
    plan = op1.plan
    n = plan.padded_size
    pslab = torch.zeros(n, device="cuda"); gslab = torch.zeros(n, device="cuda")
    acc, m, v = (torch.zeros(n, device="cuda") for _ in range(3))
    for p, o in zip(params, plan.offsets):
        pslab[o:o + p.size] = torch.from_numpy(p)
    for step in range(5):
        grads = make_grads(man, 1.0, 0, step)
        op1.run([torch.from_numpy(x).cuda() for x in grads])
        gslab.zero_()
        for x, o in zip(grads, plan.offsets):
            gslab[o:o + x.size] = torch.from_numpy(x)
        plan.step_packed(gslab.data_ptr(), pslab.data_ptr(), acc.data_ptr(), m.data_ptr(), v.data_ptr(),
                         g.StepArgs(step, 2, 0, 1e-2, 0.9, 0.999, 0.0), -1, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        for i, o in enumerate(plan.offsets):
            sz = params[i].size
            assert torch.equal(pslab[o:o + sz], op1.params[i]) and torch.equal(m[o:o + sz], op1.m_view(i).ravel())
            assert torch.equal(acc[o:o + sz], op1.accum_view(i).ravel())


def test_host_buffer_entry_point():
    import gaccum_b200 as g
    from gaccum_b200.train_op import GaccumTrainOp
    man = [("a/kernel", (40000,)), ("a/bias", (17,)), ("b/kernel", (5, 999))]
    params = make_params(man, np.random.default_rng(6))
    op = GaccumTrainOp([torch.from_numpy(p.copy()).cuda() for p in params], [n for n, _ in man], g.HParams.bert(), 3, lambda s: 1e-2)
    ref = oracle_for(man, params, onp.HParams.bert(), 3, constant_lr=1e-2)
    host_out = [torch.empty(s).pin_memory() for _, s in man]
    for step in range(8):
        grads = make_grads(man, 0.5, 0, step)
        hg = [torch.from_numpy(x).pin_memory() for x in grads]
        info = ref.run(grads)
        applied = op.run_host(hg, host_out)
        torch.cuda.synchronize()
        assert applied == info.applied
        if applied:
            for h, exp in zip(host_out, ref.params):
                assert rel_err(h.numpy(), exp) <= 1e-5
        for i in range(len(man)):
            assert np.array_equal(op.accum_view(i).cpu().numpy(), ref.accum[i])


def test_state_dict_roundtrip_mid_window():
    """Checkpoint under the reference's variable names (optimization.py:78,137-148), resume mid-window."""
    import gaccum_b200 as g
    from gaccum_b200.train_op import GaccumTrainOp
    man = [("w/kernel", (700,)), ("w/bias", (3,))]
    params = make_params(man, np.random.default_rng(8))
    mk = lambda: GaccumTrainOp([torch.from_numpy(p.copy()).cuda() for p in params], [n for n, _ in man], g.HParams.bert(), 4, lambda s: 1e-2)
    a, b = mk(), mk()
    gr = [[torch.from_numpy(x).cuda() for x in make_grads(man, 0.3, 0, s)] for s in range(11)]
    for s in range(6):
        a.run(gr[s])
    sd = a.state_dict()
    assert "w/kernel/adam_m" in sd and "w/bias/adam_v" in sd and int(sd["global_step"]) == 6
    b.load_state_dict(sd)
    for s in range(6, 11):
        a.run(gr[s]); b.run(gr[s])
    for x, y in zip(a.params, b.params):
        assert torch.equal(x, y)
    assert torch.equal(a.m, b.m) and torch.equal(a.v, b.v) and torch.equal(a.accum, b.accum)


def test_c_abi_host_session_matches_oracle():
    """gaccum_step_host: host-resident parameters/gradients, packed device state owned by the session."""
    import gaccum_b200 as g
    from gaccum_b200.train_op import HostTrainOp
    man = [("a/kernel", (40000,)), ("a/bias", (17,)), ("LayerNorm/gamma", (33,)), ("b/kernel", (5, 999))]
    params = make_params(man, np.random.default_rng(6))
    host = [torch.from_numpy(p.copy()).pin_memory() for p in params]
    op = HostTrainOp(host, [n for n, _ in man], g.HParams.bert(), 3, lambda s: 1e-2)
    ref = oracle_for(man, params, onp.HParams.bert(), 3, constant_lr=1e-2)
    for step in range(8):
        grads = make_grads(man, 0.5, 0, step)
        if step == 4:
            grads[1] = None
        hg = [None if x is None else torch.from_numpy(x).pin_memory() for x in grads]
        info = ref.run(grads)
        assert op.run(hg) == info.applied
        st = op.stats()                      # syncs
        assert st["applied"] == info.applied
        if info.applied:
            assert abs(st["global_norm"] - float(info.global_norm)) <= 2e-6 * float(info.global_norm)
        for h, exp in zip(host, ref.params):
            assert rel_err(h.numpy(), exp) <= 1e-5


def test_example_script_trains():
    """examples/mnist_gaccum.py (the 02 recipe) runs and its loss falls."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "examples", "mnist_gaccum.py"), "--steps", "300", "--lr", "1e-3"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    losses = [float(l.split("loss")[1]) for l in out.stdout.splitlines() if "loss" in l]
    assert len(losses) >= 3 and losses[-1] < 0.5 * losses[0], out.stdout


def test_estimator_contract_train_eval_and_midwindow_resume(tmp_path):
    """model_fn -> EstimatorSpec(train_op) -> Estimator.train/evaluate; checkpoint mid-window and resume
    bit-exactly (the reference gets this from the Saver seeing accum/adam_m/adam_v/global_step)."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "examples"))
    import importlib
    from gaccum_b200 import estimator as est, graph
    ex = importlib.import_module("02_single_worker_with_estimator_gaccum")
    hp = {'learning_rate': 1e-3, 'batch_size': 64, 'gradient_accumulation_multiplier': 4}
    # uninterrupted: 10 micro-steps
    a = est.Estimator(ex.model_fn, est.RunConfig(tf_random_seed=5, log_step_count_steps=0), hp)
    a.train(ex.input_fn("train", 10, 64))
    pa = [v.tensor.detach().clone() for v in graph.trainable_variables()]
    # interrupted after 6 (mid-window: 6 % 4 == 2) and resumed from the checkpoint
    d = str(tmp_path / "ckpt")
    b1 = est.Estimator(ex.model_fn, est.RunConfig(model_dir=d, tf_random_seed=5, log_step_count_steps=0), hp)
    b1.train(ex.input_fn("train", 6, 64))
    assert os.path.exists(os.path.join(d, "model.ckpt.pt"))
    saved = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in b1._spec.train_op.engine.state_dict().items()}
    assert any(float(v.abs().max()) > 0 for k, v in saved.items() if k.endswith("/accum_grad"))   # really mid-window
    b2 = est.Estimator(ex.model_fn, est.RunConfig(model_dir=d, tf_random_seed=99, log_step_count_steps=0), hp)
    b2._build(est.ModeKeys.TRAIN)                       # different init seed: everything must come from the checkpoint
    restored = b2._spec.train_op.engine.state_dict()
    assert set(restored) == set(saved) and int(restored["global_step"]) == 6
    for k, v in saved.items():
        assert torch.equal(torch.as_tensor(v).cpu(), torch.as_tensor(restored[k]).cpu()), k   # bit-exact state, incl. beta powers
    b2.train(ex.input_fn("train", 4, 64, start=6))
    assert int(graph.get_global_step()) == 10
    for x, v in zip(pa, graph.trainable_variables()):    # cuDNN's conv backward is not bit-reproducible run to run
        assert torch.allclose(x, v.tensor, rtol=1e-4, atol=1e-6), v.name
    # and it learns: train longer, evaluate
    c = est.Estimator(ex.model_fn, est.RunConfig(tf_random_seed=5, log_step_count_steps=50), hp)
    res = est.train_and_evaluate(c, ex.input_fn("train", 400, 64), ex.input_fn("eval", 5, 512, seed=1))
    assert res["accuracy"] > 0.9 and res["global_step"] == 400 and len(c.log) == 8
