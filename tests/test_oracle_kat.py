"""Hand-derived known answers for the oracle: the reference ships no tests."""
import numpy as np
import pytest

import oracle_c
import oracle_np as onp

F = np.float32
IMPLS = [onp.ReferenceTrainOp, oracle_c.COracleTrainOp]


@pytest.mark.parametrize("cls", IMPLS)
def test_single_element_variant_a_two_windows(cls):
    """T=1, numel=1, N=2, no warm-up, lr constant: follow the arithmetic by hand in fp32."""
    p0, lr, N = F(0.5), F(0.1), 2
    op = cls([np.array([p0], F)], ["w/kernel"], onp.HParams.bert(), N, constant_lr=float(lr))
    g = [F(0.4), F(0.2), F(-0.6), F(1.0), F(0.25)]
    # step 0 (g%2==0): apply with a = g0, n = g0/2 = 0.2, ||n|| = 0.2 <= 1 -> scale exactly 1
    info = op.run([np.array([g[0]], F)])
    n = F(g[0] / F(2))
    m1 = F(F(F(0.9) * F(0)) + F(F(0.1) * n))
    v1 = F(F(F(0.999) * F(0)) + F(F(0.001) * F(n * n)))
    u = F(m1 / F(np.sqrt(v1) + F(1e-6)))
    u = F(u + F(F(0.01) * p0))
    p1 = F(p0 - F(lr * u))
    assert info.applied and info.clip_scale == F(1.0)
    assert op.params[0][0] == p1 and op.m[0][0] == m1 and op.v[0][0] == v1 and op.accum[0][0] == 0
    # step 1: accumulate only
    info = op.run([np.array([g[1]], F)])
    assert not info.applied and op.accum[0][0] == g[1] and op.params[0][0] == p1
    # step 2: a = g1 + g2 = -0.4, n = -0.2
    op.run([np.array([g[2]], F)])
    n = F(F(g[1] + g[2]) / F(2))
    m2 = F(F(F(0.9) * m1) + F(F(0.1) * n))
    v2 = F(F(F(0.999) * v1) + F(F(0.001) * F(n * n)))
    u = F(F(m2 / F(np.sqrt(v2) + F(1e-6))) + F(F(0.01) * p1))
    p2 = F(p1 - F(lr * u))
    assert op.params[0][0] == p2 and op.m[0][0] == m2 and op.v[0][0] == v2
    assert op.global_step == 3


@pytest.mark.parametrize("cls", IMPLS)
def test_single_element_variant_b(cls):
    """tf.train.AdamOptimizer (TF1 ApplyAdam) by hand, N=1, two applies (bias-corrected alpha)."""
    p0, lr = F(1.0), F(1e-3)
    op = cls([np.array([p0], F)], ["dense/kernel"], onp.HParams.tf_adam(), 1, constant_lr=float(lr))
    b1, b2, eps = F(0.9), F(0.999), F(1e-8)
    p, m, v, b1p, b2p = p0, F(0), F(0), b1, b2
    for g in (F(0.3), F(-0.7)):
        op.run([np.array([g], F)])
        alpha = F(F(lr * np.sqrt(F(F(1) - b2p))) / F(F(1) - b1p))
        m = F(m + F(F(g - m) * F(F(1) - b1)))
        v = F(v + F(F(F(g * g) - v) * F(F(1) - b2)))
        p = F(p - F(F(m * alpha) / F(np.sqrt(v) + eps)))
        b1p, b2p = F(b1p * b1), F(b2p * b2)
        assert op.params[0][0] == p and op.m[0][0] == m and op.v[0][0] == v
    # first Adam step moves by ~lr regardless of gradient scale
    assert abs(float(p0) - float(op.params[0][0])) < 2.1e-3


@pytest.mark.parametrize("cls", IMPLS)
def test_zero_grad_is_pure_decay_on_decayed_tensors_only(cls):
    names = ["a/kernel", "a/bias", "x/LayerNorm/gamma", "y/layer_norm/w"]
    ps = [np.full((4,), 2.0, F) for _ in names]
    op = cls([p.copy() for p in ps], names, onp.HParams.bert(), 1, constant_lr=0.5)
    op.run([np.zeros(4, F) for _ in names])
    expect = F(F(2.0) - F(F(0.5) * F(F(0.01) * F(2.0))))          # p - lr*(0/(0+eps) + wd*p)
    assert np.all(op.params[0] == expect)
    for i in (1, 2, 3):
        assert np.all(op.params[i] == F(2.0))


def test_clip_scale_known_values():
    for fn in (onp.clip_scale, lambda g, c: F(oracle_c.lib().oracle_clip_scale(float(g), float(c)))):
        assert fn(F(0.5), 1.0) == F(1.0)           # ||n|| < clip -> exactly 1
        assert fn(F(1.0), 1.0) == F(1.0)
        assert fn(F(2.0), 1.0) == F(0.5)
        assert fn(F(0.0), 1.0) == F(1.0)           # 1/0 = inf -> min picks 1/clip; finite
        assert np.isnan(fn(F(np.inf), 1.0))        # TF 1.15: + (gn - gn)
        assert np.isnan(fn(F(np.nan), 1.0))
        assert fn(F(8.0), 2.0) == F(0.25)


def test_global_norm_of_3_4_is_5():
    assert onp.global_norm([np.array([3.0], F), np.array([4.0], F)]) == F(5.0)


@pytest.mark.parametrize("cls", IMPLS)
def test_warmup_step0_leaves_params_bit_identical(cls):
    p = np.array([0.3, -0.2], F)
    op = cls([p.copy()], ["w/kernel"], onp.HParams.bert(), 4, init_lr=2e-5, num_train_steps=100, num_warmup_steps=10)
    op.run([np.array([0.5, 0.25], F)])
    n = np.array([0.5, 0.25], F) / F(4)            # step-0 window: ONE micro-batch, still divided by N
    assert np.array_equal(op.params[0], p)
    assert np.array_equal(op.m[0], (F(0.1) * n).astype(F))
    assert np.array_equal(op.v[0], (F(0.001) * (n * n).astype(F)).astype(F))


@pytest.mark.parametrize("cls", IMPLS)
def test_n3_uses_true_division(cls):
    """(1.0*a)/N is a division by fp32(3), not a multiply by fp32(1/3): they differ in the last ulp."""
    vals = np.arange(1, 200, dtype=F) * F(0.37)
    assert np.any((vals / F(3)) != (vals * F(1.0 / 3.0)))
    hp = onp.HParams.tf_adam()
    op = cls([np.zeros_like(vals)], ["w"], hp, 3, constant_lr=0.0)
    ref_m = (F(0) + ((vals / F(3)) - F(0)) * (F(1) - F(0.9))).astype(F)
    op.run([vals])
    assert np.array_equal(op.m[0], ref_m)


@pytest.mark.parametrize("cls", IMPLS)
def test_nan_gradient_poisons_everything_on_apply(cls):
    op = cls([np.ones(3, F), np.ones(2, F)], ["a/kernel", "b/bias"], onp.HParams.bert(), 1, constant_lr=0.1)
    op.run([np.array([0, np.nan, 0], F), np.zeros(2, F)])
    assert np.isnan(op.params[0]).all() and np.isnan(op.params[1]).all()
    assert not op.accum[0].any()


def test_window_structure_pre_increment():
    op = onp.ReferenceTrainOp([np.zeros(1, F)], ["w"], onp.HParams.tf_adam(), 4, constant_lr=0.0)
    applied = [op.run([np.ones(1, F)]).applied for _ in range(10)]
    assert applied == [True, False, False, False, True, False, False, False, True, False]


def test_lr_schedule_points():
    lr = onp.learning_rate
    assert lr(2e-5, 207900, 20790, 0) == F(0.0)
    assert lr(2e-5, 207900, 0, 0) == F(2e-5)
    assert lr(2e-5, 207900, 0, 207900) == F(0.0) and lr(2e-5, 207900, 0, 10**7) == F(0.0)
    # warm-up: init_lr * (g/W) in fp32
    assert lr(2e-5, 207900, 20790, 2079) == F(F(2e-5) * F(F(2079) / F(20790)))
    # after warm-up: polynomial decay value
    assert lr(2e-5, 207900, 20790, 50000) == F(F(2e-5) * F(F(1) - F(F(50000) / F(207900))))
