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


def test_tf_checkpoint_interchange_with_real_tensorflow(tmp_path):
    """The cross-check this image cannot run: gaccum_b200.tf_checkpoint against TensorFlow's own BundleReader / Saver.
    (a) a bundle written here is readable by tf.train.load_checkpoint, tensor for tensor; (b) a checkpoint written by
    TensorFlow's Saver is readable here."""
    import numpy as np
    from gaccum_b200 import tf_checkpoint as ck
    rng = np.random.default_rng(0)
    tensors = {"bert/embeddings/word_embeddings": rng.standard_normal((50, 8)).astype(np.float32),
               "bert/embeddings/word_embeddings/adam_m": rng.standard_normal((50, 8)).astype(np.float32),
               "Variable": np.zeros((50, 8), np.float32), "Variable_1": rng.standard_normal((8,)).astype(np.float32),
               "global_step": np.asarray(123, np.int64)}
    prefix = str(tmp_path / "model.ckpt-123")
    ck.write_bundle(prefix, tensors)
    reader = tf.train.load_checkpoint(prefix)
    shapes = reader.get_variable_to_shape_map()
    assert set(shapes) == set(tensors)
    for k, v in tensors.items():
        got = reader.get_tensor(k)
        assert list(shapes[k]) == list(v.shape) and np.array_equal(got, v), k
    # (b) TensorFlow writes, this module reads
    v1 = tf.compat.v1
    g = tf.Graph()
    with g.as_default():
        vs = [v1.get_variable(k.replace("/", "_"), initializer=tf.constant(val)) for k, val in tensors.items() if val.dtype == np.float32]
        saver = v1.train.Saver()
        with v1.Session(graph=g) as sess:
            sess.run(v1.global_variables_initializer())
            p2 = saver.save(sess, str(tmp_path / "tf_written"), write_meta_graph=False)
    back = ck.read_bundle(p2)
    for k, val in tensors.items():
        if val.dtype == np.float32:
            assert np.array_equal(back[k.replace("/", "_")], val), k
