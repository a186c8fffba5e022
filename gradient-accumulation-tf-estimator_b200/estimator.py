"""The slice of the ``tf.estimator`` contract the reference scripts rely on, for the PyTorch host.

The reference's example scripts hand a ``model_fn(features, labels, mode, params)`` to
``tf.estimator.Estimator`` and get their ``train_op`` run once per micro-step by the Estimator's
session loop (distributedExample/02:20-94,112-140; 04:20-95,123-153; another-example.py:98-195).
Only what touches the train_op is mirrored here:

* ``model_fn`` is called ONCE per mode (graph construction).  ``features`` / ``labels`` are
  :class:`Placeholder` objects -- the eager stand-in for TF's symbolic tensors -- whose ``.value`` the
  Estimator refreshes before every ``train_op.run()``.
* ``EstimatorSpec(mode, loss, train_op, eval_metric_ops, predictions)`` (02:87-94).
* ``Estimator.train(input_fn, steps, max_steps)`` runs ``train_op.run()`` per batch, logs every
  ``log_step_count_steps`` (02:105), and checkpoints the variables the reference's Saver would see
  (params, ``name/adam_m``, ``name/adam_v``, accumulators, ``global_step``; optimization.py:78,137-148)
  into ``model_dir`` so that training resumes mid-window exactly.
* ``Estimator.evaluate(input_fn, steps)`` averages the callables in ``eval_metric_ops``.

Datasets, feature columns, exporters and hooks are out of scope.
"""
from __future__ import annotations

import os
import time
from typing import Any, Callable, Dict, Iterable, Optional

import torch

from . import graph


class ModeKeys:
    TRAIN, EVAL, PREDICT = "train", "eval", "infer"


class Placeholder:
    """Symbolic input: ``model_fn`` closes over it, the Estimator assigns ``.value`` per batch."""

    def __init__(self, name: str):
        self.name = name
        self.value: Any = None

    def get(self):
        if self.value is None:
            raise RuntimeError(f"placeholder {self.name} has not been fed")
        return self.value


class EstimatorSpec:
    def __init__(self, mode, loss=None, train_op=None, eval_metric_ops: Optional[Dict[str, Callable[[], torch.Tensor]]] = None,
                 predictions=None):
        if mode == ModeKeys.TRAIN and train_op is None:
            raise ValueError("EstimatorSpec in TRAIN mode needs a train_op")
        self.mode, self.loss, self.train_op = mode, loss, train_op
        self.eval_metric_ops = eval_metric_ops or {}
        self.predictions = predictions


class RunConfig:
    def __init__(self, model_dir: Optional[str] = None, tf_random_seed: Optional[int] = None,
                 log_step_count_steps: int = 100, save_checkpoints_steps: Optional[int] = None,
                 train_distribute=None):
        self.model_dir, self.tf_random_seed = model_dir, tf_random_seed
        self.log_step_count_steps, self.save_checkpoints_steps = log_step_count_steps, save_checkpoints_steps
        self.train_distribute = train_distribute       # a torch.distributed process group (or True) == 04:106,113-119


class Estimator:
    def __init__(self, model_fn: Callable, config: Optional[RunConfig] = None, params: Optional[dict] = None,
                 model_dir: Optional[str] = None):
        self.model_fn, self.params = model_fn, params or {}
        self.config = config or RunConfig(model_dir=model_dir)
        if model_dir and not self.config.model_dir:
            self.config.model_dir = model_dir
        self._spec = None
        self._features, self._labels = Placeholder("features"), Placeholder("labels")
        self.log: list = []

    # -- graph construction (once) ------------------------------------------------------------
    def _build(self, mode) -> EstimatorSpec:
        if self._spec is not None:
            return self._spec
        if self.config.tf_random_seed is not None:
            torch.manual_seed(self.config.tf_random_seed)
        graph.reset_default_graph()
        import inspect
        if len(inspect.signature(self.model_fn).parameters) >= 5:      # model_fn(features, labels, mode, params, config)
            spec = self.model_fn(self._features, self._labels, mode, self.params, self.config)
        else:
            spec = self.model_fn(self._features, self._labels, mode, self.params)
        self._spec = spec
        self._maybe_restore()
        return spec

    def _ckpt_path(self) -> Optional[str]:
        return os.path.join(self.config.model_dir, "model.ckpt.pt") if self.config.model_dir else None

    def _maybe_restore(self) -> None:
        path = self._ckpt_path()
        if path and os.path.exists(path) and self._spec.train_op is not None:
            sd = torch.load(path, map_location="cpu")
            op = self._spec.train_op
            op.load_state_dict({k: v.to(op.engine.device) if torch.is_tensor(v) and v.dim() else v for k, v in sd.items()})

    def save_checkpoint(self) -> Optional[str]:
        path = self._ckpt_path()
        if path and self._spec is not None and self._spec.train_op is not None:
            import torch.distributed as dist
            sd = self._spec.train_op.state_dict()          # collective under data parallelism: full moments, summed accumulators
            multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
            if not multi or dist.get_rank() == 0:          # one writer (04: the chief saves)
                os.makedirs(self.config.model_dir, exist_ok=True)
                torch.save({k: (v.cpu() if torch.is_tensor(v) else v) for k, v in sd.items()}, path)
            if multi:
                dist.barrier()
        return path

    # -- the loops --------------------------------------------------------------------------------
    def train(self, input_fn: Callable[[], Iterable], steps: Optional[int] = None, max_steps: Optional[int] = None):
        spec = self._build(ModeKeys.TRAIN)
        gs = graph.get_or_create_global_step()
        done, t0, last = 0, time.time(), int(gs)
        for features, labels in input_fn():
            if (steps is not None and done >= steps) or (max_steps is not None and int(gs) >= max_steps):
                break
            self._features.value, self._labels.value = features, labels
            loss = spec.train_op.run()                     # one session.run(train_op): one micro-step
            done += 1
            n = self.config.log_step_count_steps
            if n and int(gs) % n == 0:
                dt = time.time() - t0
                rate = (int(gs) - last) / dt if dt > 0 else float("nan")
                self.log.append((int(gs), float(loss) if loss is not None else None, rate))
                t0, last = time.time(), int(gs)
            s = self.config.save_checkpoints_steps
            if s and int(gs) % s == 0:
                self.save_checkpoint()
        self.save_checkpoint()
        return self

    def evaluate(self, input_fn: Callable[[], Iterable], steps: Optional[int] = None) -> Dict[str, float]:
        spec = self._build(ModeKeys.TRAIN if self._spec is None else self._spec.mode)
        sums: Dict[str, float] = {}
        count = 0
        with torch.no_grad():
            for features, labels in input_fn():
                if steps is not None and count >= steps:
                    break
                self._features.value, self._labels.value = features, labels
                for k, fn in spec.eval_metric_ops.items():
                    sums[k] = sums.get(k, 0.0) + float(fn())
                if callable(spec.loss):
                    sums["loss"] = sums.get("loss", 0.0) + float(spec.loss())
                count += 1
        out = {k: v / max(count, 1) for k, v in sums.items()}
        out["global_step"] = int(graph.get_or_create_global_step())
        return out


def train_and_evaluate(estimator: Estimator, train_input_fn, eval_input_fn, train_steps: Optional[int] = None,
                       eval_steps: Optional[int] = None) -> Dict[str, float]:
    """``tf.estimator.train_and_evaluate`` reduced to: train, then evaluate once (02:136-140)."""
    estimator.train(train_input_fn, steps=train_steps)
    return estimator.evaluate(eval_input_fn, steps=eval_steps)
