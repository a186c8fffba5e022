"""Pins the CPU oracle (numpy and C) against fixtures produced by the reference's own
optimization.py executed over oracle/tf_stub (tests/golden/make_golden.py).  Bit-exact."""
import numpy as np
import pytest

import oracle_c
import oracle_np as onp
from golden_util import Golden, cases


def _mk(cls, g):
    return cls(g.init(), g.names, onp.HParams.bert(), g.N, init_lr=g.init_lr,
               num_train_steps=g.num_train_steps, num_warmup_steps=g.num_warmup_steps)


def test_fixtures_present():
    assert set(cases()) >= {"warmup_unclipped", "nowarmup_clipped", "mixed_short_decay"}


@pytest.mark.parametrize("case", cases())
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_oracle_reproduces_reference_trajectory(case, impl):
    g = Golden(case)
    op = _mk(onp.ReferenceTrainOp if impl == "numpy" else oracle_c.COracleTrainOp, g)
    applied_steps = 0
    for s in range(g.steps):
        info = op.run(g.grads(s))
        applied_steps += bool(info.applied)
        assert op.global_step == g.global_step(s)
        for kind, mine in (("param", op.params), ("accum", op.accum), ("m", op.m), ("v", op.v)):
            for name, a, b in zip(g.names, mine, g.state(s, kind)):
                assert np.array_equal(a, b), f"{case}/{impl}: step {s} {kind} {name} differs from the reference run"
    # windows {0}, {1..N}, {N+1..2N}, ...: the pre-increment predicate applies at steps 0, N, 2N, ...
    assert applied_steps == len(range(0, g.steps, g.N))


@pytest.mark.parametrize("case", cases())
def test_reference_semantics_visible_in_fixtures(case):
    """Properties SURVEY.md 0 calls out, read straight off the reference's own run."""
    g = Golden(case)
    p0 = g.init()
    # step 0 applies a single micro-batch (scaled by 1/N) and zeroes the accumulators
    assert all(not a.any() for a in g.state(0, "accum"))
    assert any(m.any() for m in g.state(0, "m"))
    if g.num_warmup_steps:
        # lr(0) = 0 with warm-up: m, v move, params do not
        assert all(np.array_equal(a, b) for a, b in zip(p0, g.state(0, "param")))
    # steps 1..N-1 only accumulate
    for s in range(1, g.N):
        assert all(np.array_equal(a, b) for a, b in zip(g.state(s, "param"), g.state(0, "param")))
        assert any(a.any() for a in g.state(s, "accum"))
    assert all(not a.any() for a in g.state(g.N, "accum"))


# ---- variant B: the inline recipes of 02 / 04 / another-example (AST-lifted, tf.train.AdamOptimizer) ----
from golden_util import DirectApplyGolden, RecipeGolden, recipe_cases


def test_recipe_fixtures_present():
    assert set(recipe_cases()) >= {"02_small_n2", "02_mnist_n4", "04_small_n2", "another_example_n3"}
    assert {"n4_clipped", "n3_warmup"} <= set(cases())


@pytest.mark.parametrize("case", recipe_cases())
@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_oracle_reproduces_recipe_trajectory(case, impl):
    """Pins variant B (a14): window logic with global_step=None, no clip, N from params, ApplyAdam, beta powers."""
    g = RecipeGolden(case)
    cls = onp.ReferenceTrainOp if impl == "numpy" else oracle_c.COracleTrainOp
    op = cls(g.init(), g.names, onp.HParams.tf_adam(), g.N, constant_lr=g.lr)
    for s in range(g.steps):
        info = op.run(g.grads(s))
        assert bool(info.applied) == (s % g.N == 0)
        assert op.global_step == int(g.z[f"global_step/{s}"])
        if s in g.recorded:
            for i, n in enumerate(g.names):
                g.check(f"param/{s}/{n}", op.params[i]); g.check(f"accum/{s}/{n}", op.accum[i])
                g.check(f"m/{s}/{n}", op.m[i]); g.check(f"v/{s}/{n}", op.v[i])
            assert np.float32(op.beta1_power) == g.z[f"beta1_power/{s}"] and np.float32(op.beta2_power) == g.z[f"beta2_power/{s}"]


@pytest.mark.parametrize("impl", ["numpy", "c"])
def test_oracle_reproduces_direct_apply_with_none_gradient(impl):
    """AdamWeightDecayOptimizer.apply_gradients(zip(grads, tvars)) with one grad None (optimization.py:132-133):
    the pair is skipped altogether -- no slots, no weight decay.  The oracle runs over the remaining pairs."""
    g = DirectApplyGolden()
    keep = [i for i in range(len(g.names)) if i != g.none_at]
    cls = onp.ReferenceTrainOp if impl == "numpy" else oracle_c.COracleTrainOp
    hp = onp.HParams.bert(); hp.clip_norm = 0.0
    op = cls([g.z[f"init/{g.names[i]}"].copy() for i in keep], [g.names[i] for i in keep], hp, 1, constant_lr=g.lr)
    for s in range(g.steps):
        op.run([g.z[f"grad/{s}/{g.names[i]}"] for i in keep])
        for k, i in enumerate(keep):
            n = g.names[i]
            assert np.array_equal(op.params[k], g.z[f"param/{s}/{n}"]) and np.array_equal(op.m[k], g.z[f"m/{s}/{n}"])
            assert np.array_equal(op.v[k], g.z[f"v/{s}/{n}"])
        skipped = g.names[g.none_at]
        assert np.array_equal(g.z[f"param/{s}/{skipped}"], g.z[f"init/{skipped}"])


def test_c_oracle_reproduces_the_reference_at_bert_small_size():
    """BASELINE config 2 shapes (T=73, P=28.8 M), accum x4, clipped: the reference's optimization.py over the stub vs the
    C oracle, bit for bit on the stored subsample and fp64 sums (the numpy oracle is bit-identical to the C one)."""
    from golden_util import FullsizeGolden
    g = FullsizeGolden("bert_small_n4")
    assert len(g.names) == 73 and sum(int(np.prod(s)) for s in g.shapes) == 28764674
    op = oracle_c.COracleTrainOp(g.init(), g.names, onp.HParams.bert(), g.N, init_lr=g.init_lr,
                                 num_train_steps=g.num_train_steps, num_warmup_steps=g.num_warmup_steps)
    clipped = 0
    for s in range(g.steps):
        info = op.run(g.grads(s))
        clipped += bool(info.applied and float(info.clip_scale) < 1.0)
        assert op.global_step == int(g.z[f"global_step/{s}"])
        if s in g.recorded:
            for i, n in enumerate(g.names):
                g.check(f"param/{s}/{n}", op.params[i]); g.check(f"accum/{s}/{n}", op.accum[i])
                g.check(f"m/{s}/{n}", op.m[i]); g.check(f"v/{s}/{n}", op.v[i])
    assert clipped >= 2
