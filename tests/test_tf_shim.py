"""Activates only where TensorFlow exists (it does not in this image): the real-TF drop-in."""
import os

import pytest

tf = pytest.importorskip("tensorflow", reason="TensorFlow is not installable in this image (DESIGN.md)")


def test_tf_shim_signature():
    import importlib.util
    import inspect
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location(
        "gaccum_tf_optimization", os.path.join(here, "gradient-accumulation-tf-estimator_b200", "tf_shim", "optimization.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert list(inspect.signature(mod.create_optimizer).parameters) == \
        ["loss", "init_lr", "num_train_steps", "num_warmup_steps", "use_tpu"]
