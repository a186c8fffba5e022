"""Executes the PYTHON half of the TensorFlow shim (tf_shim/optimization.py -- the file a maintainer drops in place of the
reference's optimization.py) without TensorFlow: over oracle/tf_stub, whose ``tf.load_op_library`` returns an emulation of
the ``GaccumStep`` node with the C ABI's ``gaccum_step`` semantics (evaluated by the CPU oracle's per-op functions).

What this pins: the graph the shim builds around the ONE custom-op node -- discovery of variables and gradients, the
packed state variables, the in-graph learning-rate schedule, which value of ``global_step`` the node / the beta-power
update / the increment see, skipping of (None, var) pairs, the 5-argument signature -- reproduces the reference's own
runs (tests/golden/, produced by the reference's code) bit for bit, in BOTH evaluation orders of unordered op sets
(``tf.group`` inputs forwards and backwards: a missing control dependency shows up as a difference).
The C++ adapter gaccum_tf_op.cc is covered separately: tests/test_tf_op_adapter.py (compiled against tests/tf_mock)."""
import importlib.util
import inspect
import os
import sys

import numpy as np
import pytest

from golden_util import DirectApplyGolden, Golden, RecipeGolden, cases, recipe_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(ROOT, "oracle", "tf_stub")


@pytest.fixture()
def shim():
    saved = sys.modules.get("tensorflow")
    sys.path.insert(0, STUB)
    sys.modules.pop("tensorflow", None)
    import tensorflow as tf
    assert "stub" in tf.__version__
    spec = importlib.util.spec_from_file_location(
        "gaccum_tf_shim_optimization", os.path.join(ROOT, "gradient-accumulation-tf-estimator_b200", "tf_shim", "optimization.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    yield tf, mod
    tf.REVERSE_UNORDERED = False
    sys.path.remove(STUB)
    sys.modules.pop("tensorflow", None)
    if saved is not None:
        sys.modules["tensorflow"] = saved


def test_shim_exports_the_reference_surface(shim):
    _, mod = shim
    assert list(inspect.signature(mod.create_optimizer).parameters) == ["loss", "init_lr", "num_train_steps", "num_warmup_steps", "use_tpu"]
    opt = mod.AdamWeightDecayOptimizer(learning_rate=1e-3, weight_decay_rate=0.01, exclude_from_weight_decay=["LayerNorm", "layer_norm", "bias"])
    assert opt._do_use_weight_decay("bert/encoder/layer_0/output/dense/kernel") and not opt._do_use_weight_decay("a/LayerNorm/gamma")
    assert opt._get_variable_name("scope/kernel:0") == "scope/kernel"
    with pytest.raises(ValueError):
        mod.create_optimizer(None, 1e-3, 10, 0, True)                     # use_tpu is refused, not ignored


@pytest.mark.parametrize("reverse", [False, True])
@pytest.mark.parametrize("case", cases())
def test_shim_create_optimizer_reproduces_the_reference_runs(shim, case, reverse):
    tf, mod = shim
    g = Golden(case)
    tf.reset_default_graph()
    tf.REVERSE_UNORDERED = reverse
    for n, v in zip(g.names, g.init()):
        tf.get_variable(n, shape=list(v.shape), dtype=tf.float32, initializer=v)
    mod.gradient_accumulation_multiplier = g.N                              # optimization.py:76 is a literal; the shim makes it a module attribute
    train_op = mod.create_optimizer(tf.constant(0.0), g.init_lr, g.num_train_steps, g.num_warmup_steps, False)
    ph = {p.name[len("grad/"):]: p for p in tf._g.placeholders}
    by = {v.name: v for v in tf.global_variables()}
    assert {"gaccum/accum_grads:0", "gaccum/adam_m:0", "gaccum/adam_v:0"} <= set(by)      # Saver-visible state, three packed variables
    sess = tf.Session()
    for s in range(g.steps):
        sess.run(train_op, feed_dict={ph[n + ":0"]: x for n, x in zip(g.names, g.grads(s))})
        assert int(by["global_step:0"].value) == g.global_step(s)
        for n, exp in zip(g.names, g.state(s, "param")):
            assert np.array_equal(by[n + ":0"].value, exp), f"{case} reverse={reverse} step {s} {n}"
    # the packed accumulator is zero after the last apply and holds per-tensor views at 32-element aligned offsets
    off = 0
    for n, exp in zip(g.names, g.state(g.steps - 1, "accum")):
        assert np.array_equal(by["gaccum/accum_grads:0"].value[off:off + exp.size].reshape(exp.shape), exp)
        off += (exp.size + 31) // 32 * 32


@pytest.mark.parametrize("reverse", [False, True])
@pytest.mark.parametrize("case", [c for c in recipe_cases() if "mnist" not in c])
def test_shim_inline_recipe_variant_b_reproduces_the_reference_runs(shim, case, reverse):
    """gaccum_train_op(variant=1): tf.train.AdamOptimizer inside the window, beta powers advance on apply steps only."""
    tf, mod = shim
    g = RecipeGolden(case)
    tf.reset_default_graph()
    tf.REVERSE_UNORDERED = reverse
    for n, v in zip(g.names, g.init()):
        tf.get_variable(n, shape=list(v.shape), dtype=tf.float32, initializer=v)
    tf.train.get_or_create_global_step()
    train_op = mod.gaccum_train_op(tf.constant(0.0), g.lr, g.N, variant=1, epsilon=1e-8, weight_decay_rate=0.0, clip_norm=None)
    ph = {p.name[len("grad/"):]: p for p in tf._g.placeholders}
    by = {v.name: v for v in tf.global_variables()}
    sess = tf.Session()
    for s in range(g.steps):
        sess.run(train_op, feed_dict={ph[n + ":0"]: x for n, x in zip(g.names, g.grads(s))})
        assert int(by["global_step:0"].value) == int(g.z[f"global_step/{s}"])
        if s in g.recorded:
            for n in g.names:
                g.check(f"param/{s}/{n}", by[n + ":0"].value)
            bp = by["gaccum/beta_powers:0"].value
            assert np.float32(bp[0]) == g.z[f"beta1_power/{s}"] and np.float32(bp[1]) == g.z[f"beta2_power/{s}"]


def test_shim_direct_apply_gradients_skips_none_pairs(shim):
    """optimization.AdamWeightDecayOptimizer(...).apply_gradients(zip(grads, tvars)) -- legal against the reference
    (optimization.py:128-177) -- works against the shim and skips (None, var) pairs (:132-133)."""
    tf, mod = shim
    gd = DirectApplyGolden()
    tf.reset_default_graph()
    tvars = [tf.get_variable(n, shape=list(gd.z[f"init/{n}"].shape), dtype=tf.float32, initializer=gd.z[f"init/{n}"]) for n in gd.names]
    opt = mod.AdamWeightDecayOptimizer(learning_rate=gd.lr, weight_decay_rate=0.01, beta_1=0.9, beta_2=0.999, epsilon=1e-6,
                                       exclude_from_weight_decay=["LayerNorm", "layer_norm", "bias"])
    phs = [None if i == gd.none_at else tf.placeholder(tf.float32, v.shape, name="g/" + v.name) for i, v in enumerate(tvars)]
    train_op = opt.apply_gradients(zip(phs, tvars))
    sess = tf.Session()
    for s in range(gd.steps):
        sess.run(train_op, feed_dict={phs[i]: gd.z[f"grad/{s}/{n}"] for i, n in enumerate(gd.names) if i != gd.none_at})
        for i, n in enumerate(gd.names):
            assert np.array_equal(tvars[i].value, gd.z[f"param/{s}/{n}"]), f"step {s} {n}"
    assert tf.train.get_global_step() is None or int(tf.train.get_global_step().value) == 0     # :99-101: not incremented here


# ---- checkpoint interchange between a reference run and the shim's graph ---------------------------------
def _pick_mid_window(g):
    """a micro-step after which the reference's accumulators are non-zero (the checkpoint is taken mid-window)"""
    for s in range(g.N + 1, g.steps - 2):
        if any(np.any(a != 0) for a in g.state(s, "accum")):
            return s
    raise AssertionError("fixture has no mid-window step")


def _reference_checkpoint(g, s):
    """what the reference's Saver holds after micro-step s of the fixture's run (keys: optimization.py:78, 137-148)"""
    ref = {"global_step": np.asarray(g.global_step(s), np.int64)}
    for i, n in enumerate(g.names):
        ref[n] = g.z[f"param/{s}/{n}"]
        ref[n + "/adam_m"], ref[n + "/adam_v"] = g.z[f"m/{s}/{n}"], g.z[f"v/{s}/{n}"]
        ref["Variable" if i == 0 else f"Variable_{i}"] = g.z[f"accum/{s}/{n}"]
    return ref


@pytest.mark.parametrize("case", cases())
def test_reference_checkpoint_restores_into_the_shim_graph_mid_window_and_back(shim, case, tmp_path):
    """A TF-format checkpoint holding the REFERENCE's state mid-window (taken from the fixture the reference's own code
    produced) is read back, repacked into the shim's three slab variables and restored into the shim's graph: the rest of
    the run reproduces the reference's run bit for bit.  And the other way: the shim's variables after the run, mapped
    to the reference's Saver keys, equal the reference's final state."""
    from gaccum_b200 import tf_checkpoint as ck
    tf, mod = shim
    g = Golden(case)
    cut = _pick_mid_window(g)
    prefix = str(tmp_path / f"model.ckpt-{g.global_step(cut)}")
    ck.write_bundle(prefix, _reference_checkpoint(g, cut))                  # the file a reference run would have left behind
    restored = ck.to_shim_names(ck.read_bundle(prefix), g.names, ck.ADAM_WEIGHT_DECAY)
    tf.reset_default_graph()
    for n, v in zip(g.names, g.init()):
        tf.get_variable(n, shape=list(v.shape), dtype=tf.float32, initializer=np.full_like(v, 7.0))     # garbage: everything must come from the file
    mod.gradient_accumulation_multiplier = g.N
    train_op = mod.create_optimizer(tf.constant(0.0), g.init_lr, g.num_train_steps, g.num_warmup_steps, False)
    ph = {p.name[len("grad/"):]: p for p in tf._g.placeholders}
    by = {v.name: v for v in tf.global_variables()}
    assert set(restored) == {k[:-2] for k in by}                             # exactly the variables a Saver over this graph has
    for k, arr in restored.items():                                          # == saver.restore(sess, prefix)
        assert by[k + ":0"].value.shape == arr.shape, k
        by[k + ":0"].value = arr.astype(by[k + ":0"].value.dtype)
    sess = tf.Session()
    for s in range(cut + 1, g.steps):
        sess.run(train_op, feed_dict={ph[n + ":0"]: x for n, x in zip(g.names, g.grads(s))})
        for n, exp in zip(g.names, g.state(s, "param")):
            assert np.array_equal(by[n + ":0"].value, exp), f"{case} resumed at {cut}, step {s} {n}"
    back = ck.from_shim_names({k[:-2]: v.value for k, v in by.items()}, g.names, ck.ADAM_WEIGHT_DECAY)
    want = _reference_checkpoint(g, g.steps - 1)
    assert set(back) == set(want)
    for k in want:
        assert np.array_equal(back[k], want[k]), k
