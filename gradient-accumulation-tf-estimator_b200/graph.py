"""TF1-style graph collections for the PyTorch host side.

The reference finds its inputs through global collections: ``tf.trainable_variables()``
(optimization.py:70) and ``tf.train.get_or_create_global_step()`` (optimization.py:27).
``create_optimizer`` keeps its 5-argument signature, so the host mirror needs the same two
lookups.  This module is that registry -- nothing more (no ops, no sessions).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch


class Variable:
    """A named tensor.  ``name`` follows TF conventions (``scope/kernel:0``)."""

    def __init__(self, name: str, tensor: torch.Tensor, trainable: bool = True):
        self.name = name if ":" in name else name + ":0"
        self.tensor = tensor
        self.trainable = trainable

    def __repr__(self):
        return f"<Variable {self.name} shape={tuple(self.tensor.shape)} trainable={self.trainable}>"


class GlobalStep:
    """int64 scalar, incremented once per micro-step (optimization.py:102-103)."""

    def __init__(self, value: int = 0):
        self.value = int(value)

    def __int__(self):
        return self.value

    def assign(self, v: int):
        self.value = int(v)


class Graph:
    def __init__(self):
        self.variables: List[Variable] = []
        self.by_name: Dict[str, Variable] = {}
        self.global_step: Optional[GlobalStep] = None


_default = Graph()


def get_default_graph() -> Graph:
    return _default


def reset_default_graph() -> None:
    global _default
    _default = Graph()


def add_variable(name: str, tensor: torch.Tensor, trainable: bool = True) -> Variable:
    v = Variable(name, tensor, trainable)
    if v.name in _default.by_name:
        raise ValueError(f"variable {v.name} already exists")
    _default.variables.append(v)
    _default.by_name[v.name] = v
    return v


def get_variable(name: str, shape=None, initializer: Optional[Callable] = None, trainable: bool = True,
                 device=None, dtype=torch.float32) -> Variable:
    """``tf.get_variable`` (optimization.py:137-148): create-or-return by name."""
    key = name if ":" in name else name + ":0"
    if key in _default.by_name:
        return _default.by_name[key]
    t = torch.zeros(tuple(shape or ()), dtype=dtype, device=device)
    if initializer is not None:
        initializer(t)
    return add_variable(name, t, trainable)


def trainable_variables() -> List[Variable]:
    """``tf.trainable_variables()``: creation order (optimization.py:70)."""
    return [v for v in _default.variables if v.trainable]


def get_or_create_global_step() -> GlobalStep:
    if _default.global_step is None:
        _default.global_step = GlobalStep(0)
    return _default.global_step


def get_global_step() -> Optional[GlobalStep]:
    return _default.global_step


def register_module(module: torch.nn.Module, rename: Optional[Callable[[str], str]] = None) -> List[Variable]:
    """Register every ``requires_grad`` parameter of a module as a trainable variable, in
    ``named_parameters()`` order.  ``rename`` maps torch names to TF names (the decay mask of
    optimization.py:179-187 is name-driven: ``LayerNorm`` / ``layer_norm`` / ``bias``)."""
    out = []
    for n, p in module.named_parameters():
        if not p.requires_grad:
            continue
        tf_name = rename(n) if rename else n.replace(".", "/")
        out.append(add_variable(tf_name, p, trainable=True))
    return out
