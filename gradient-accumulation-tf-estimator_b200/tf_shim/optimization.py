# coding=utf-8
"""Drop-in replacement for the reference's ``optimization.py`` on a machine that HAS TensorFlow.

Put this file where ``run_classifier.py`` imports ``optimization`` from (or the inline recipes of
distributedExample/02,04 call ``gaccum_train_op``) -- nothing else in the caller changes:

    train_op = optimization.create_optimizer(total_loss, learning_rate, num_train_steps,
                                             num_warmup_steps, use_tpu)

Same five arguments, same return type (a ``tf.Operation``), same checkpointed state (variables),
but instead of the ``tf.cond`` over 3T ``assign_add`` / ~17T un-fused Adam ops
(reference optimization.py:76-94, 128-177) the graph contains ONE stateful custom-op node,
``GaccumStep`` (gaccum_tf_op.cc), that calls ``gaccum_step`` in libgaccum.so.

CANNOT BE EXECUTED IN THIS REPOSITORY'S IMAGE (TensorFlow is not installable; tests/
test_tf_shim.py activates wherever ``import tensorflow`` works).  Written against
``tf.compat.v1`` so it runs on TF 1.15 and TF 2.x alike.
"""
from __future__ import absolute_import, division, print_function

import os
import re

import tensorflow as _tf

tf = _tf.compat.v1 if hasattr(_tf, "compat") and hasattr(_tf.compat, "v1") else _tf

_HERE = os.path.dirname(os.path.abspath(__file__))
_ops = None

# optimization.py:76 hard-codes 8 (README.md:17,34 says 4); no argument exists for it.
gradient_accumulation_multiplier = int(os.environ.get("GACCUM_MULTIPLIER", "8"))


def _load():
    global _ops
    if _ops is None:
        _ops = tf.load_op_library(os.path.join(_HERE, "libgaccum_tf.so"))
    return _ops


def _decay_mask(names, weight_decay_rate, exclude):
    """optimization.py:179-194."""
    out = []
    for n in names:
        m = re.match("^(.*):\\d+$", n)
        n = m.group(1) if m is not None else n
        use = bool(weight_decay_rate)
        for r in (exclude or []):
            if re.search(r, n) is not None:
                use = False
        out.append(use)
    return out


def _slab_size(tvars):
    return sum((int(v.shape.num_elements()) + 31) // 32 * 32 for v in tvars)


def gaccum_train_op(loss, learning_rate, accum_n, variant=0, beta1=0.9, beta2=0.999, epsilon=1e-6,
                    weight_decay_rate=0.01, clip_norm=1.0, exclude_from_weight_decay=None,
                    global_step=None, increment_global_step=True):
    """The accumulate-then-apply train_op as one node (the recipe of 02:47-73 / 04:48-74 /
    optimization.py:70-104)."""
    global_step = global_step if global_step is not None else tf.train.get_or_create_global_step()
    tvars = tf.trainable_variables()                                   # optimization.py:70
    grads = tf.gradients(loss, tvars)                                  # optimization.py:71
    pairs = [(g, v) for g, v in zip(grads, tvars) if g is not None]    # optimization.py:132
    grads, tvars = [g for g, _ in pairs], [v for _, v in pairs]
    n = _slab_size(tvars)
    with tf.variable_scope("gaccum", reuse=tf.AUTO_REUSE):
        mk = lambda name: tf.get_variable(name, shape=[n], dtype=tf.float32, trainable=False,
                                          initializer=tf.zeros_initializer(), use_resource=False)
        accum, m, v = mk("accum_grads"), mk("adam_m"), mk("adam_v")   # :78, :137-148 packed
        beta_powers = tf.get_variable("beta_powers", dtype=tf.float32, trainable=False, use_resource=False,
                                      initializer=tf.constant([beta1, beta2], dtype=tf.float32))
    mask = _decay_mask([v_.name for v_ in tvars], weight_decay_rate if variant == 0 else 0.0,
                       exclude_from_weight_decay)
    step_op = _load().gaccum_step(
        params=tvars, grads=grads, accum=accum, m=m, v=v, global_step=global_step,
        lr=tf.cast(learning_rate, tf.float32), beta_powers=beta_powers, accum_n=accum_n, variant=variant,
        beta1=beta1, beta2=beta2, epsilon=epsilon, weight_decay_rate=weight_decay_rate,
        clip_norm=clip_norm or 0.0, decay_mask=mask)
    ops = [step_op]
    if variant == 1:
        # TF1 AdamOptimizer._finish: beta powers advance after every APPLY (pre-increment predicate)
        with tf.control_dependencies([step_op]):
            is_apply = tf.equal(tf.cast(global_step, tf.int32) % accum_n, 0)
            ops.append(tf.cond(is_apply,
                               lambda: beta_powers.assign(beta_powers * tf.constant([beta1, beta2], tf.float32)),
                               lambda: tf.identity(beta_powers)))
    if increment_global_step:
        with tf.control_dependencies([step_op]):        # defined ordering: pre-increment
            ops.append(global_step.assign(global_step + 1))           # optimization.py:102-103
    return tf.group(*ops)


def create_optimizer(loss, init_lr, num_train_steps, num_warmup_steps, use_tpu):
    """Creates an optimizer training op (same signature as the reference, optimization.py:25)."""
    if use_tpu:
        raise ValueError("use_tpu=True is not supported by the B200 train_op")
    global_step = tf.train.get_or_create_global_step()
    # optimization.py:29-54 verbatim semantics (this IS TensorFlow, so the schedule stays in-graph)
    learning_rate = tf.constant(value=init_lr, shape=[], dtype=tf.float32)
    learning_rate = tf.train.polynomial_decay(learning_rate, global_step, num_train_steps,
                                              end_learning_rate=0.0, power=1.0, cycle=False)
    if num_warmup_steps:
        global_steps_int = tf.cast(global_step, tf.int32)
        warmup_steps_int = tf.constant(num_warmup_steps, dtype=tf.int32)
        warmup_percent_done = tf.cast(global_steps_int, tf.float32) / tf.cast(warmup_steps_int, tf.float32)
        warmup_learning_rate = init_lr * warmup_percent_done
        is_warmup = tf.cast(global_steps_int < warmup_steps_int, tf.float32)
        learning_rate = (1.0 - is_warmup) * learning_rate + is_warmup * warmup_learning_rate
    return gaccum_train_op(loss, learning_rate, gradient_accumulation_multiplier, variant=0,
                           beta1=0.9, beta2=0.999, epsilon=1e-6, weight_decay_rate=0.01, clip_norm=1.0,
                           exclude_from_weight_decay=["LayerNorm", "layer_norm", "bias"],
                           global_step=global_step)


class AdamWeightDecayOptimizer(object):
    """Kept for import compatibility (reference optimization.py:107); the math lives in the kernel."""

    def __init__(self, learning_rate, weight_decay_rate=0.0, beta_1=0.9, beta_2=0.999, epsilon=1e-6,
                 exclude_from_weight_decay=None, name="AdamWeightDecayOptimizer"):
        self.learning_rate, self.weight_decay_rate = learning_rate, weight_decay_rate
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
        self.exclude_from_weight_decay, self.name = exclude_from_weight_decay, name
