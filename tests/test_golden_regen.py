"""The committed golden fixtures are exactly what tests/golden/make_golden.py produces from the
reference's own optimization.py (only checkable where /root/reference exists: the authoring container)."""
import importlib.util
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = "/root/reference/optimization.py"


@pytest.mark.skipif(not os.path.exists(REFERENCE), reason="/root/reference is not present (GPU box)")
def test_fixtures_regenerate_bit_identically(tmp_path, monkeypatch):
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    monkeypatch.setattr(mg, "HERE", str(tmp_path))
    saved_path, saved_mods = list(sys.path), {k: sys.modules.get(k) for k in ("tensorflow", "optimization")}
    try:
        for name, cfg in mg.CASES.items():
            path, steps = mg.run_case(name, *cfg)
            new = np.load(path)
            old = np.load(os.path.join(HERE, "golden", os.path.basename(path)))
            assert set(new.files) == set(old.files)
            for k in new.files:
                assert np.array_equal(new[k], old[k]), f"{name}: {k} differs from the committed fixture"
    finally:
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.skipif(not os.path.exists(REFERENCE), reason="/root/reference is not present (GPU box)")
def test_recipe_fixtures_regenerate_bit_identically(tmp_path, monkeypatch):
    """Same for tests/golden/make_golden_recipes.py: the AST-lifted statements of 02 / 04 / another-example and the
    N-patched optimization.py produce exactly the committed ref_recipe_* / ref_optimization_n* / ref_direct_* files."""
    spec = importlib.util.spec_from_file_location("make_golden_recipes", os.path.join(HERE, "golden", "make_golden_recipes.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    monkeypatch.setattr(mg, "HERE", str(tmp_path))
    saved_path, saved_mods = list(sys.path), {k: sys.modules.get(k) for k in ("tensorflow", "optimization")}
    try:
        made = []
        path, lifted = mg.run_recipe("02_small_n2", "distributedExample/02_single_worker_with_estimator_gaccum.py", ["model_fn"],
                                     [(38, 41), (47, 73)],
                                     lambda tf: {"params": {"learning_rate": 1e-4, "batch_size": 100, "gradient_accumulation_multiplier": 2},
                                                 "loss": tf.constant(0.0)}, mg.SMALL, 2, 1e-4, 0.5, 21, 9, None)
        assert lifted[0][0] == 38 and lifted[-1][1] == 73
        made.append(path)
        path, _ = mg.run_recipe("another_example_n3", "another-example.py", ["model_fn"], [(126, 155)],
                                lambda tf: {"gradient_accumulation_multiplier": 3}, mg.SMALL, 3, 1e-3, 0.5, 24, 11,
                                lambda ns, tf: ns["_train_op_fn"](tf.constant(0.0)))
        made.append(path)
        path, edits = mg.run_optimization_case("n3_warmup", 3, 1e-2, 40, 5, 0.2, 32, 8)
        assert edits == [(76, 8, 3)]
        made.append(path)
        made.append(mg.run_direct_apply_with_none("direct_apply_none_grad", 1e-3, 0.3, 33, 3))
        for path in made:
            new = np.load(path)
            old = np.load(os.path.join(HERE, "golden", os.path.basename(path)))
            assert set(new.files) == set(old.files), os.path.basename(path)
            for k in new.files:
                assert np.array_equal(new[k], old[k]), f"{os.path.basename(path)}: {k} differs from the committed fixture"
    finally:
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
