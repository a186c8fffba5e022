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
    # windows {0}, {1..8}, {9..16}: three applies in 18 micro-steps (pre-increment predicate)
    assert applied_steps == 3


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
    # steps 1..7 only accumulate
    for s in range(1, 8):
        assert all(np.array_equal(a, b) for a, b in zip(g.state(s, "param"), g.state(0, "param")))
        assert any(a.any() for a in g.state(s, "accum"))
    assert all(not a.any() for a in g.state(8, "accum"))
