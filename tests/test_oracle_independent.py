"""Independent cross-checks of the oracle's restated TensorFlow primitives (`ApplyAdam`, `clip_by_global_norm`,
`polynomial_decay` are restated from TF 1.15, which cannot be executed here) against implementations that DO exist in this
image -- PyTorch's, written by other people from the same papers -- in the regimes where the two definitions coincide.
These do not replace the fixtures (which pin the reference's Python); they bound the risk that a restated primitive is wrong
in a way the stub and the oracle share.  CPU only; torch is used as a checker here, never by the product path."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle_np as onp  # noqa: E402


def test_variant_b_window_equals_torch_adam_on_the_mean_gradient_when_epsilon_is_negligible():
    """TF1 ApplyAdam (`alpha = lr*sqrt(1-b2^t)/(1-b1^t); p -= alpha*m/(sqrt(v)+eps)`, "epsilon hat") and torch.optim.Adam
    (`p -= lr/(1-b1^t) * m/(sqrt(v)/sqrt(1-b2^t)+eps)`) are the same update when eps << sqrt(v).  The reference's recipe
    (another-example.py:126-155) with window N is then Adam on the mean of the window's gradients -- except for its first
    window, which holds ONE micro-batch and is still divided by N (pre-increment predicate)."""
    rng = np.random.default_rng(7)
    shapes, N, steps, lr = [(17, 5), (5,)], 3, 13, 1e-3
    init = [rng.standard_normal(s).astype(np.float32) for s in shapes]
    grads = [[(rng.standard_normal(s) * 1e-1).astype(np.float32) for s in shapes] for _ in range(steps)]
    hp = onp.HParams.tf_adam()
    hp.epsilon = 1e-12
    ref = onp.ReferenceTrainOp([x.copy() for x in init], ["w", "b"], hp, N, constant_lr=lr)
    tp = [torch.tensor(x.astype(np.float64), requires_grad=True) for x in init]
    # ApplyAdam holds beta1/beta2 as fp32 tensors and computes 1 - beta in fp32 (1 - fp32(0.999) is 1.3e-5 below 0.001):
    # give torch the same two numbers, everything else of its update runs in fp64
    opt = torch.optim.Adam(tp, lr=lr, betas=(float(np.float32(0.9)), float(np.float32(0.999))), eps=1e-12)
    window = [np.zeros(s, np.float64) for s in shapes]
    for s in range(steps):
        info = ref.run(grads[s])
        for w, g in zip(window, grads[s]):
            w += g.astype(np.float64)
        assert info.applied == (s % N == 0)
        if info.applied:
            for p, w in zip(tp, window):
                p.grad = torch.from_numpy(w / N)                    # (1.0 * accum) / N, whatever the window held
            opt.step()
            window = [np.zeros(s_, np.float64) for s_ in shapes]
            for mine, theirs in zip(ref.params, tp):
                np.testing.assert_allclose(mine, theirs.detach().numpy(), rtol=2e-5, atol=2e-7)
    st = opt.state[tp[0]]
    np.testing.assert_allclose(ref.m[0], st["exp_avg"].numpy(), rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(ref.v[0], st["exp_avg_sq"].numpy(), rtol=1e-5, atol=1e-12)
    assert int(st["step"]) == 5 and abs(float(ref.beta1_power) - 0.9 ** 6) < 1e-6   # powers advance once per APPLY


@pytest.mark.parametrize("scale", [1e-3, 1.0, 30.0])
def test_global_norm_and_clip_scale_against_torch_clip_grad_norm(scale):
    """tf.clip_by_global_norm: gn = sqrt(sum ||t||^2), scale = clip*min(1/gn, 1/clip); torch: total_norm likewise,
    coefficient clamp(max_norm/(total_norm+1e-6), max=1) -- equal up to the 1e-6 in the denominator."""
    rng = np.random.default_rng(3)
    ts = [(rng.standard_normal(s) * scale).astype(np.float32) for s in [(64, 33), (7,), (1,), (129,)]]
    gn = onp.global_norm(ts)
    tt = [torch.zeros(t.shape, dtype=torch.float64, requires_grad=True) for t in ts]
    for p, t in zip(tt, ts):
        p.grad = torch.from_numpy(t.astype(np.float64))
    total = float(torch.nn.utils.clip_grad_norm_(tt, max_norm=1.0))
    assert abs(float(gn) - total) <= 1e-6 * total
    s = float(onp.clip_scale(gn, 1.0))
    want = min(1.0 / (total + 1e-6), 1.0)
    assert abs(s - want) <= 3e-6 * want
    if total < 1.0:
        assert s == 1.0                                              # exactly: clip*min(1/gn, 1/clip) with gn < clip
    for p, t in zip(tt, ts):                                         # and the clipped tensors themselves
        np.testing.assert_allclose(t * np.float32(s), p.grad.numpy(), rtol=5e-6, atol=1e-12)


def test_learning_rate_schedule_against_torch_polynomial_lr_and_linear_warmup():
    """tf.train.polynomial_decay(power=1, end=0, cycle=False) == torch PolynomialLR(power=1); BERT's warm-up is
    init_lr * g / W below W (optimization.py:45-54) and switches to the DECAY value (not a re-based one) at W."""
    init_lr, T, W = 2e-5, 1000, 100
    p = torch.zeros(1, requires_grad=True)
    opt = torch.optim.SGD([p], lr=init_lr)
    sched = torch.optim.lr_scheduler.PolynomialLR(opt, total_iters=T, power=1.0)
    for g in range(T + 50):
        theirs = opt.param_groups[0]["lr"]
        mine = float(onp.learning_rate(init_lr, T, 0, g))
        assert abs(mine - theirs) <= 2e-7 * init_lr + 1e-12, (g, mine, theirs)
        with_warmup = float(onp.learning_rate(init_lr, T, W, g))
        if g < W:
            assert abs(with_warmup - init_lr * g / W) <= 2e-7 * init_lr
        else:
            assert with_warmup == mine
        opt.step(); sched.step()
    assert float(onp.learning_rate(init_lr, T, W, 0)) == 0.0 and float(onp.learning_rate(init_lr, T, 0, T + 10)) == 0.0


def test_adam_weight_decay_is_adamw_without_bias_correction_closed_form_fp64():
    """optimization.py:151-171 against an independent fp64 evaluation of the paper's update with the two deviations the
    reference makes: no bias correction, and decay added to the update BEFORE the learning rate (decoupled, lr-scaled)."""
    rng = np.random.default_rng(11)
    p0 = rng.standard_normal((40, 9)); g_seq = rng.standard_normal((6, 40, 9)) * 1e-2
    lr, b1, b2, eps, wd = 1e-3, 0.9, 0.999, 1e-6, 0.01
    p, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    P, M, V = p0.astype(np.float32), np.zeros(p0.shape, np.float32), np.zeros(p0.shape, np.float32)
    for g in g_seq:
        m = b1 * m + (1 - b1) * g
        v = b2 * v + (1 - b2) * g * g
        p = p - lr * (m / (np.sqrt(v) + eps) + wd * p)
        P, M, V = onp.adam_weight_decay_update(P, M, V, g.astype(np.float32), np.float32(lr), b1, b2, eps, wd, True)
    np.testing.assert_allclose(P, p, rtol=3e-6, atol=1e-7)
    np.testing.assert_allclose(M, m, rtol=3e-6, atol=1e-9)
    np.testing.assert_allclose(V, v, rtol=3e-6, atol=1e-12)
    Pn, _, _ = onp.adam_weight_decay_update(p0.astype(np.float32), np.zeros(p0.shape, np.float32), np.zeros(p0.shape, np.float32),
                                            np.zeros(p0.shape, np.float32), np.float32(lr), b1, b2, eps, wd, False)
    assert np.array_equal(Pn, p0.astype(np.float32))                 # excluded from decay + zero gradient: untouched
