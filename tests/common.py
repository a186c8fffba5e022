"""Shared helpers for the parity tests: seeded synthetic cases."""
import numpy as np

import oracle_np as onp

SEED0 = 19830610   # the reference's own tf_random_seed (02:106, another-example.py:285)


def make_params(manifest, rng):
    """BERT-style init: kernels/embeddings N(0, 0.02^2), LayerNorm gamma 1 / beta 0, biases 0."""
    out = []
    for name, shape in manifest:
        if name.endswith("gamma"):
            a = np.ones(shape, np.float32)
        elif name.endswith("beta") or "bias" in name:
            a = np.zeros(shape, np.float32)
        else:
            a = rng.normal(0.0, 0.02, size=shape).astype(np.float32)
        out.append(a)
    return out


def make_grads(manifest, sigma, rank, step):
    rng = np.random.Generator(np.random.PCG64(SEED0 + 1000 * rank + step))
    return [rng.normal(0.0, sigma, size=shape).astype(np.float32) for _, shape in manifest]


def rel_err(x, ref):
    x = np.asarray(x, np.float64); ref = np.asarray(ref, np.float64)
    d = np.max(np.abs(x - ref)) if x.size else 0.0
    return d / max(np.max(np.abs(ref)) if ref.size else 0.0, 1e-30)


def oracle_for(manifest, params, hp, N, **kw):
    return onp.ReferenceTrainOp([p.copy() for p in params], [n for n, _ in manifest], hp, N, **kw)
