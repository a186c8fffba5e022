"""Randomised GPU parity: shapes around the 2048-element tile boundary, views at arbitrary 4-byte
offsets, both optimizer variants, clipping on/off, odd hyper-parameters, missing gradients."""
import numpy as np
import pytest

import oracle_np as onp

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

SIZES = [1, 2, 3, 5, 31, 32, 33, 127, 2047, 2048, 2049, 4095, 4096, 4097, 6143, 10000, 70001]
NAMES = ["a/kernel", "a/bias", "LayerNorm/gamma", "emb/word_embeddings", "x/layer_norm/beta", "out/kernel", "out/bias", "c/kernel"]


@pytest.mark.parametrize("seed", range(16))
def test_random_configuration_matches_oracle(seed):
    import gaccum_b200 as g
    from gaccum_b200.train_op import GaccumTrainOp
    rng = np.random.default_rng(500 + seed)
    T = int(rng.integers(1, 9))
    sizes = [int(rng.choice(SIZES)) for _ in range(T)]
    names = [NAMES[i] for i in rng.permutation(len(NAMES))[:T]]
    N = int(rng.integers(1, 6))
    variant_b = bool(seed % 2)
    hp_o = onp.HParams.tf_adam() if variant_b else onp.HParams.bert()
    hp_g = g.HParams.tf_adam() if variant_b else g.HParams.bert()
    if seed % 3 == 0:
        clip = 0.0 if not variant_b else 0.5
        hp_o.clip_norm = clip; hp_g.clip_norm = clip
    b1 = float(rng.choice([0.9, 0.8, 0.95]))
    hp_o.beta1 = b1; hp_g.beta1 = b1
    if not variant_b:
        wd = float(rng.choice([0.01, 0.0, 0.1]))
        hp_o.weight_decay_rate = wd; hp_g.weight_decay_rate = wd
    lr = float(rng.choice([1e-4, 1e-3]))
    params = [rng.normal(0, 0.05, s).astype(np.float32) for s in sizes]
    ref = onp.ReferenceTrainOp([p.copy() for p in params], names, hp_o, N, constant_lr=lr)
    aligned = bool(seed % 4 == 1)
    # parameters and gradients as views of flat buffers at random element offsets (4-byte alignment only)
    pflat = torch.zeros(sum(sizes) + 64 + 4 * T, device="cuda")
    gflat = torch.zeros_like(pflat)
    tp, offs, o = [], [], int(rng.integers(0, 4))
    for p in params:
        if aligned:
            o = (o + 3) // 4 * 4
        t = pflat[o:o + p.size]; t.copy_(torch.from_numpy(p)); tp.append(t); offs.append(o)
        o += p.size + int(rng.integers(0, 4))
    op = GaccumTrainOp(tp, names, hp_g, N, lambda s: lr)
    sigma = float(rng.choice([1e-3, 0.3, 5.0]))
    exact = True
    clipping = hp_o.clip_norm > 0
    for step in range(2 * N + 3):
        grads = [rng.normal(0, sigma, s).astype(np.float32) for s in sizes]
        if step % 4 == 3 and T > 1:
            grads[int(rng.integers(0, T))] = None
        tg = []
        for x, off in zip(grads, offs):
            if x is None:
                tg.append(None); continue
            v = gflat[off:off + x.size]; v.copy_(torch.from_numpy(x)); tg.append(v)
        info = ref.run(grads)
        assert op.run(tg) == info.applied
        if info.applied and clipping:
            st = op.stats()
            assert abs(st["global_norm"] - float(info.global_norm)) <= 3e-6 * max(float(info.global_norm), 1e-30)
            if not (st["clip_scale"] == 1.0 and float(info.clip_scale) == 1.0):
                exact = False
        for i in range(T):
            for name, got, exp in (("p", tp[i], ref.params[i]), ("m", op.m_view(i), ref.m[i]),
                                   ("v", op.v_view(i), ref.v[i]), ("a", op.accum_view(i), ref.accum[i])):
                got = got.cpu().numpy()
                if exact or name == "a":
                    assert np.array_equal(got, exp, equal_nan=True), (seed, step, i, name)
                else:
                    d = np.max(np.abs(got.astype(np.float64) - exp)) / max(float(np.max(np.abs(exp))), 1e-30)
                    assert d <= 1e-5, (seed, step, i, name, d)
    if variant_b:
        assert np.float32(op.beta1_power) == ref.beta1_power and np.float32(op.beta2_power) == ref.beta2_power
