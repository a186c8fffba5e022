"""Randomised cross-check of the two CPU restatements (numpy vs C): bit-identical on arbitrary
shapes, window lengths, both optimizer variants, clipping on/off, missing gradients."""
import numpy as np
import pytest

import oracle_c
import oracle_np as onp

NAMES = ["a/kernel", "a/bias", "LayerNorm/gamma", "emb/word_embeddings", "x/layer_norm/beta", "out/kernel", "out/bias"]


@pytest.mark.parametrize("seed", range(24))
def test_numpy_and_c_oracles_agree_bitwise(seed):
    rng = np.random.default_rng(1000 + seed)
    T = int(rng.integers(1, 7))
    shapes = [tuple(int(x) for x in rng.integers(1, 40, size=int(rng.integers(1, 3)))) for _ in range(T)]
    if seed % 5 == 0:
        shapes[0] = (0,)                       # empty tensor
    names = list(rng.permutation(NAMES)[:T])
    N = int(rng.integers(1, 6))
    variant_b = bool(seed % 2)
    hp = onp.HParams.tf_adam() if variant_b else onp.HParams.bert()
    if seed % 3 == 0:
        hp.clip_norm = 0.0 if not variant_b else 0.5      # also: variant B WITH clipping, variant A without
    hp.beta1 = float(rng.choice([0.9, 0.8, 0.95]))
    hp.weight_decay_rate = float(rng.choice([0.01, 0.0, 0.1])) if not variant_b else 0.0
    kw = dict(constant_lr=float(rng.choice([1e-4, 1e-2]))) if seed % 4 else \
        dict(init_lr=1e-3, num_train_steps=int(rng.integers(5, 40)), num_warmup_steps=int(rng.integers(0, 6)))
    params = [rng.normal(0, 0.05, s).astype(np.float32) for s in shapes]
    a = onp.ReferenceTrainOp([p.copy() for p in params], names, hp, N, **kw)
    b = oracle_c.COracleTrainOp([p.copy() for p in params], names, hp, N, **kw)
    sigma = float(rng.choice([1e-3, 0.3, 5.0]))
    for step in range(2 * N + 3):
        grads = [rng.normal(0, sigma, s).astype(np.float32) for s in shapes]
        if step % 4 == 3 and T > 1:
            grads[int(rng.integers(0, T))] = None
        ia, ib = a.run(grads), b.run(grads)
        assert ia.applied == ib.applied and ia.lr == ib.lr
        assert np.array_equal(np.float32(ia.global_norm), np.float32(ib.global_norm))
        assert np.array_equal(np.float32(ia.clip_scale), np.float32(ib.clip_scale))
        for la, lb in ((a.params, b.params), (a.m, b.m), (a.v, b.v), (a.accum, b.accum)):
            for x, y in zip(la, lb):
                assert np.array_equal(x, y, equal_nan=True)
    if variant_b:
        assert a.beta1_power == b.beta1_power and a.beta2_power == b.beta2_power
