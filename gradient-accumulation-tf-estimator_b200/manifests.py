"""Variable manifests (TF-style names + shapes, ``tf.trainable_variables()`` creation order) of the
models BASELINE.json names.  Names matter: the weight-decay mask is name-driven
(reference optimization.py:65,179-187).

* ``mnist_cnn``  -- distributedExample/02:22-28: Conv2D(32,3) -> MaxPool -> Flatten -> Dense(64) -> Dense(10)
* ``bert_*``     -- upstream google-research/bert ``modeling.py`` (referenced by the reference's
                    README.md:14) with the classifier head of ``run_classifier.py``: vocab 30522,
                    512 positions, 2 token types, 2 labels.  T = 5 + 16 L + 4.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

Manifest = List[Tuple[str, Tuple[int, ...]]]


def bert_manifest(num_layers: int, hidden: int, intermediate: Optional[int] = None, vocab: int = 30522,
                  max_pos: int = 512, type_vocab: int = 2, num_labels: int = 2) -> Manifest:
    inter = intermediate or 4 * hidden
    e = "bert/embeddings/"
    out: Manifest = [(e + "word_embeddings", (vocab, hidden)), (e + "token_type_embeddings", (type_vocab, hidden)),
                     (e + "position_embeddings", (max_pos, hidden)),
                     (e + "LayerNorm/beta", (hidden,)), (e + "LayerNorm/gamma", (hidden,))]
    for l in range(num_layers):
        b = f"bert/encoder/layer_{l}/"
        for nm in ("query", "key", "value"):
            out += [(b + f"attention/self/{nm}/kernel", (hidden, hidden)), (b + f"attention/self/{nm}/bias", (hidden,))]
        out += [(b + "attention/output/dense/kernel", (hidden, hidden)), (b + "attention/output/dense/bias", (hidden,)),
                (b + "attention/output/LayerNorm/beta", (hidden,)), (b + "attention/output/LayerNorm/gamma", (hidden,)),
                (b + "intermediate/dense/kernel", (hidden, inter)), (b + "intermediate/dense/bias", (inter,)),
                (b + "output/dense/kernel", (inter, hidden)), (b + "output/dense/bias", (hidden,)),
                (b + "output/LayerNorm/beta", (hidden,)), (b + "output/LayerNorm/gamma", (hidden,))]
    out += [("bert/pooler/dense/kernel", (hidden, hidden)), ("bert/pooler/dense/bias", (hidden,)),
            ("output_weights", (num_labels, hidden)), ("output_bias", (num_labels,))]
    return out


def mnist_cnn_manifest() -> Manifest:
    return [("conv2d/kernel", (3, 3, 1, 32)), ("conv2d/bias", (32,)), ("dense/kernel", (5408, 64)),
            ("dense/bias", (64,)), ("dense_1/kernel", (64, 10)), ("dense_1/bias", (10,))]


MANIFESTS = {
    "mnist_cnn": mnist_cnn_manifest,
    "bert_small": lambda: bert_manifest(4, 512),      # L4_H512_A8
    "bert_base": lambda: bert_manifest(12, 768),      # L12_H768_A12
    "bert_large": lambda: bert_manifest(24, 1024),    # L24_H1024_A16
}
