"""GaccumTrainOp -- the train_op object ``create_optimizer`` returns (host side, PyTorch plumbing).

Holds what the reference's graph holds as variables: ``accum_grads`` (optimization.py:78),
``adam_m`` / ``adam_v`` (optimization.py:137-148), ``global_step`` (optimization.py:27) and, for
the plain-Adam variant, TF1's ``beta{1,2}_power`` non-slot variables -- but accum/m/v are three
packed fp32 slabs in HBM instead of 3T separate variables.  ``run(grads)`` is one
``session.run(train_op)``: exactly one kernel launch through the C ABI, then ``global_step += 1``
(optimization.py:102-103).  PyTorch supplies device memory and the CUDA stream; all arithmetic is
in csrc/.
"""
from __future__ import annotations

import ctypes as C
import re
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import ADAM, HParams, Plan, StepArgs, Stats


def _f32(x) -> float:
    return float(np.float32(x))


class GaccumTrainOp:
    def __init__(self, params: Sequence[torch.Tensor], names: Sequence[str], hp: HParams, accum_n: int,
                 lr_fn: Callable[[int], float],
                 exclude_from_weight_decay: Optional[Sequence[str]] = ("LayerNorm", "layer_norm", "bias"),
                 global_step: int = 0, accum: Optional[torch.Tensor] = None):
        if len(params) != len(names):
            raise ValueError("params and names differ in length")
        if not params:
            raise ValueError("no trainable variables")
        dev = params[0].device
        if dev.type != "cuda":
            raise _lib.GaccumError(_lib.ENODEVICE, "parameters must live on a CUDA device: the train_op has no CPU fallback")
        for p in params:
            if p.dtype != torch.float32 or not p.is_contiguous() or p.device != dev:
                raise ValueError("parameters must be contiguous fp32 tensors on one device")
        self.device = dev
        self.params = list(params)
        # optimization.py:135-143 names slots after _get_variable_name(param.name): the ":0" tensor suffix is stripped
        self.names = [re.sub(r":\d+$", "", n) for n in names]
        self.hp = hp
        self.N = int(accum_n)
        self.lr_fn = lr_fn
        self.global_step = int(global_step)
        self.decay = _lib.decay_mask(self.names, hp.weight_decay_rate, exclude_from_weight_decay) \
            if hp.variant == _lib.ADAM_WEIGHT_DECAY else [False] * len(params)
        self.plan = Plan([p.numel() for p in params], self.decay, hp, device=dev.index or 0)
        n = max(self.plan.padded_size, 32)
        if accum is not None:          # caller-provided slab (e.g. NVLink peer-mapped symmetric memory)
            if accum.numel() < n or accum.dtype != torch.float32 or accum.device != dev:
                raise ValueError("accum must be an fp32 tensor of at least plan.padded_size elements on the params' device")
            self.accum = accum
        else:
            self.accum = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        # TF1 AdamOptimizer._create_slots: beta powers start at beta
        self.beta1_power = _f32(hp.beta1)
        self.beta2_power = _f32(hp.beta2)
        self._accum_ptr, self._m_ptr, self._v_ptr = self.accum.data_ptr(), self.m.data_ptr(), self.v.data_ptr()
        self._param_ptrs = Plan.ptr_array([p.data_ptr() for p in self.params])
        self._param_key = tuple(p.data_ptr() for p in self.params)
        self._grad_key = None
        self._grad_ptrs = None
        self._stats_host = torch.zeros(4, dtype=torch.float32).pin_memory() if torch.cuda.is_available() else None
        self.last_lr = 0.0
        self.last_applied = False

    # -- views under the reference's variable names -----------------------------------------
    def _view(self, slab: torch.Tensor, i: int) -> torch.Tensor:
        o = self.plan.offsets[i]
        return slab[o:o + self.params[i].numel()].view(self.params[i].shape)

    def accum_view(self, i: int) -> torch.Tensor:
        return self._view(self.accum, i)

    def m_view(self, i: int) -> torch.Tensor:
        return self._view(self.m, i)

    def v_view(self, i: int) -> torch.Tensor:
        return self._view(self.v, i)

    def _args(self, lr: float) -> StepArgs:
        return StepArgs(self.global_step, self.N, 0, lr, self.beta1_power, self.beta2_power, 0.0)

    def _grad_table(self, grads: Sequence[Optional[torch.Tensor]]):
        key = tuple(0 if g is None else g.data_ptr() for g in grads)
        if key != self._grad_key:
            for g, p in zip(grads, self.params):
                if g is None:
                    continue
                if g.dtype != torch.float32 or not g.is_contiguous() or g.numel() != p.numel() or g.device != self.device:
                    raise ValueError("gradients must be contiguous fp32 tensors shaped like their parameter")
            self._grad_ptrs = Plan.ptr_array(list(key))
            self._grad_key = key
        return self._grad_ptrs

    def _refresh_param_table(self):
        key = tuple(p.data_ptr() for p in self.params)
        if key != self._param_key:
            self._param_ptrs = Plan.ptr_array(list(key))
            self._param_key = key

    # -- one micro-step -------------------------------------------------------------------
    def run(self, grads: Sequence[Optional[torch.Tensor]]) -> bool:
        """One ``session.run(train_op)`` (optimization.py:91-104).  Returns True if it applied."""
        self._refresh_param_table()
        return self.run_bound(self._grad_table(grads))

    def bind(self, grads: Sequence[Optional[torch.Tensor]]):
        """Validate a gradient list once and return its device-pointer table.  A graph-mode caller
        (TF hands the op raw pointers) pays nothing per step; Python callers whose gradient buffers
        are persistent can do the same with ``run_bound(bind(grads))``."""
        self._grad_key = None
        table = self._grad_table(grads)
        self._grad_key = None
        return table

    def run_bound(self, grad_table, stream: Optional[int] = None) -> bool:
        """``run`` with a pointer table from ``bind`` (no per-step Python work over T tensors).
        Parameters must not have been re-allocated since construction / the last ``run``."""
        g = self.global_step
        lr = self.lr_fn(g)
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        self.plan.step(grad_table, self._param_ptrs, self._accum_ptr, self._m_ptr, self._v_ptr,
                       StepArgs(g, self.N, 0, lr, self.beta1_power, self.beta2_power, 0.0), stream)
        applied = (g % self.N) == 0 if 0 <= g < 2**31 else _lib.is_apply_step(g, self.N)
        self._after(applied, lr)
        return applied

    def accumulate_only(self, grads: Sequence[Optional[torch.Tensor]]) -> None:
        """The false branch alone (used by the data-parallel driver); does not touch global_step."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.plan.accumulate(self._grad_table(grads), self.accum.data_ptr(), stream)

    def apply_only(self, grads: Optional[Sequence[Optional[torch.Tensor]]] = None) -> None:
        """The true branch alone; ``grads=None`` applies the accumulators as they are."""
        lr = _f32(self.lr_fn(self.global_step))
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._refresh_param_table()
        self.plan.apply(self._grad_table(grads) if grads is not None else None, self._param_ptrs,
                        self.accum.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self._args(lr), stream)
        self.last_lr, self.last_applied = lr, True
        if self.hp.variant == ADAM:
            self.beta1_power = _f32(np.float32(self.beta1_power) * np.float32(self.hp.beta1))
            self.beta2_power = _f32(np.float32(self.beta2_power) * np.float32(self.hp.beta2))

    def _after(self, applied: bool, lr: float) -> None:
        self.last_lr, self.last_applied = lr, applied
        if applied and self.hp.variant == ADAM:
            # TF1 AdamOptimizer._finish: beta_power <- beta_power * beta (fp32) after the updates
            self.beta1_power = _f32(np.float32(self.beta1_power) * np.float32(self.hp.beta1))
            self.beta2_power = _f32(np.float32(self.beta2_power) * np.float32(self.hp.beta2))
        self.global_step += 1                                   # optimization.py:102-103

    __call__ = run

    # -- host-buffer entry point (what a CPU-resident caller such as the TF CPU graph sees) ----
    def _host_init(self):
        n = self.accum.numel()
        self._stage = [torch.empty(n, dtype=torch.float32, device=self.device) for _ in range(2)]
        self._stage_views = [[self._view(s, i) for i in range(len(self.params))] for s in self._stage]
        self._copy_stream = torch.cuda.Stream(self.device)
        self._d2h_stream = torch.cuda.Stream(self.device)
        self._buf_free = [torch.cuda.Event() for _ in range(2)]
        self._h2d_done = [torch.cuda.Event() for _ in range(2)]
        self._k_done, self._d2h_done = torch.cuda.Event(), torch.cuda.Event()
        self._host_calls = 0
        self._stats_ring = torch.zeros(4, dtype=torch.float32).pin_memory()

    def run_host(self, host_grads: Sequence[Optional[torch.Tensor]],
                 host_params_out: Optional[Sequence[torch.Tensor]] = None) -> bool:
        """One micro-step whose gradients live in (pinned) HOST memory.

        H2D of this step's gradients (copy stream, double-buffered so it overlaps the previous
        kernel), one kernel launch, D2H of the 16-byte stats block, and -- on apply steps -- D2H of
        the updated parameters into ``host_params_out``.  Everything is stream-ordered; the caller's
        current stream observes completion.  Returns True if the step applied.
        """
        if not hasattr(self, "_stage"):
            self._host_init()
        main = torch.cuda.current_stream(self.device)
        b = self._host_calls & 1
        self._host_calls += 1
        self._copy_stream.wait_event(self._buf_free[b])
        with torch.cuda.stream(self._copy_stream):
            for view, hg in zip(self._stage_views[b], host_grads):
                if hg is not None:
                    view.copy_(hg.view(view.shape), non_blocking=True)
            self._h2d_done[b].record(self._copy_stream)
        main.wait_event(self._h2d_done[b])
        grads = [v if hg is not None else None for v, hg in zip(self._stage_views[b], host_grads)]
        applied = self.run(grads)
        self._buf_free[b].record(main)
        self.plan.read_stats(self._stats_ring.data_ptr(), main.cuda_stream)
        if applied and host_params_out is not None:
            self._k_done.record(main)
            self._d2h_stream.wait_event(self._k_done)
            with torch.cuda.stream(self._d2h_stream):
                for hp_, p in zip(host_params_out, self.params):
                    hp_.copy_(p, non_blocking=True)
                self._d2h_done.record(self._d2h_stream)
            main.wait_event(self._d2h_done)
        return applied

    def stats(self) -> Dict[str, float]:
        """Synchronously read what the last step computed (applied, lr, global_norm, clip_scale)."""
        stream = torch.cuda.current_stream(self.device)
        self.plan.read_stats(self._stats_host.data_ptr(), stream.cuda_stream)
        stream.synchronize()
        a, lr, gn, s = self._stats_host.tolist()
        return {"applied": bool(a), "lr": lr, "global_norm": gn, "clip_scale": s}

    # -- checkpoint compatibility: the reference's Saver sees per-variable tensors --------
    def state_dict(self) -> Dict[str, torch.Tensor]:
        """Per-tensor copies under the reference's names (optimization.py:78, 137-148)."""
        out: Dict[str, torch.Tensor] = {"global_step": torch.tensor(self.global_step, dtype=torch.int64)}
        for i, n in enumerate(self.names):
            out[n] = self.params[i].detach().clone()
            out[n + "/adam_m"] = self.m_view(i).clone()
            out[n + "/adam_v"] = self.v_view(i).clone()
            out[n + "/accum_grad"] = self.accum_view(i).clone()
        if self.hp.variant == ADAM:
            out["beta1_power"] = torch.tensor(self.beta1_power, dtype=torch.float32)
            out["beta2_power"] = torch.tensor(self.beta2_power, dtype=torch.float32)
        return out

    def load_state_dict(self, sd: Dict[str, torch.Tensor], strict: bool = True) -> None:
        self.global_step = int(sd["global_step"])
        for i, n in enumerate(self.names):
            with torch.no_grad():
                if n in sd:
                    self.params[i].copy_(sd[n])
                for suffix, view in (("/adam_m", self.m_view), ("/adam_v", self.v_view), ("/accum_grad", self.accum_view)):
                    if n + suffix in sd:
                        view(i).copy_(sd[n + suffix])
                    elif strict:
                        raise KeyError(n + suffix)
        if self.hp.variant == ADAM and "beta1_power" in sd:
            self.beta1_power = float(sd["beta1_power"])
            self.beta2_power = float(sd["beta2_power"])


class PackedTrainOp:
    """SURVEY.md 8(f) #2 -- gradient hand-off without the scatter.

    The reference keeps T parameters, T gradients and T ``accum_grads`` as separate tensors
    (optimization.py:70-71, 78) and adds gradient to accumulator with T ``assign_add`` ops per micro-step (:81, 93).
    Here the parameters are re-pointed at views of ONE flat slab in the plan's layout and every ``p.grad`` is a view
    of the packed accumulator slab, so the producer (autograd's ``AccumulateGrad``: ``grad += new_grad``, one fp32
    rounding per micro-step -- the same arithmetic as ``assign_add``) accumulates IN PLACE, straight into ``accum``:
      * accumulate micro-steps launch nothing from this library (12 -> 0 B/param of train_op traffic),
      * the apply micro-step runs the slab-to-slab kernel with no gradient stream and no pointer table
        (``gaccum_step_packed(grad_slab=NULL)``: 32 B/param instead of 36) and leaves ``accum`` == ``p.grad`` zeroed
        for the next window.
    Call ``step()`` once per micro-step, after ``loss.backward()``.  Never set the gradients to ``None``
    (``zero_grad(set_to_none=True)``) -- the views are the accumulator.
    """

    def __init__(self, params: Sequence[torch.Tensor], names: Sequence[str], hp: HParams, accum_n: int,
                 lr_fn: Callable[[int], float],
                 exclude_from_weight_decay: Optional[Sequence[str]] = ("LayerNorm", "layer_norm", "bias"),
                 global_step: int = 0):
        if not params:
            raise ValueError("no trainable variables")
        dev = params[0].device
        if dev.type != "cuda":
            raise _lib.GaccumError(_lib.ENODEVICE, "parameters must live on a CUDA device: the train_op has no CPU fallback")
        self.device, self.params, self.names = dev, list(params), list(names)
        self.hp, self.N, self.lr_fn, self.global_step = hp, int(accum_n), lr_fn, int(global_step)
        self.decay = _lib.decay_mask(self.names, hp.weight_decay_rate, exclude_from_weight_decay) \
            if hp.variant == _lib.ADAM_WEIGHT_DECAY else [False] * len(params)
        self.plan = Plan([p.numel() for p in params], self.decay, hp, device=dev.index or 0)
        n = max(self.plan.padded_size, 32)
        self.param_slab = torch.zeros(n, dtype=torch.float32, device=dev)
        self.accum = torch.zeros(n, dtype=torch.float32, device=dev)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, self.plan.offsets):
                if p.dtype != torch.float32:
                    raise ValueError("parameters must be fp32")
                view = self.param_slab[o:o + p.numel()].view(p.shape)
                view.copy_(p)
                p.data = view                                            # the caller's tensors now alias the slab
                p.grad = self.accum[o:o + p.numel()].view(p.shape)      # autograd accumulates straight into accum_grads
        self.beta1_power, self.beta2_power = _f32(hp.beta1), _f32(hp.beta2)
        self._stats_host = torch.zeros(4, dtype=torch.float32).pin_memory()
        self.launches = 0

    def _view(self, slab, i):
        o = self.plan.offsets[i]
        return slab[o:o + self.params[i].numel()].view(self.params[i].shape)

    def accum_view(self, i): return self._view(self.accum, i)
    def m_view(self, i): return self._view(self.m, i)
    def v_view(self, i): return self._view(self.v, i)

    def step(self, stream: Optional[int] = None) -> bool:
        """One ``session.run(train_op)`` AFTER the backward pass has added this micro-batch's gradient into
        ``p.grad`` (== accum).  Returns True if it applied (optimization.py:91, pre-increment predicate)."""
        g = self.global_step
        applied = _lib.is_apply_step(g, self.N)
        if applied:
            for p, o in zip(self.params, self.plan.offsets):     # a caller that dropped the views broke the contract
                if p.grad is None or p.grad.data_ptr() != self.accum.data_ptr() + 4 * o:
                    raise RuntimeError("p.grad no longer aliases the packed accumulator (zero_grad(set_to_none=True)?)")
            lr = self.lr_fn(g)
            if stream is None:
                stream = torch.cuda.current_stream(self.device).cuda_stream
            self.plan.step_packed(0, self.param_slab.data_ptr(), self.accum.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                  StepArgs(g, self.N, 0, lr, self.beta1_power, self.beta2_power, 0.0), 1, stream)
            self.launches += 1
            if self.hp.variant == ADAM:
                self.beta1_power = _f32(np.float32(self.beta1_power) * np.float32(self.hp.beta1))
                self.beta2_power = _f32(np.float32(self.beta2_power) * np.float32(self.hp.beta2))
        self.global_step = g + 1
        return applied

    def stats(self) -> Dict[str, float]:
        stream = torch.cuda.current_stream(self.device)
        self.plan.read_stats(self._stats_host.data_ptr(), stream.cuda_stream)
        stream.synchronize()
        a, lr, gn, s = self._stats_host.tolist()
        return {"applied": bool(a), "lr": lr, "global_norm": gn, "clip_scale": s}


class HostTrainOp:
    """The train_op for a caller whose parameters and gradients live in HOST memory (the reference's
    CPU placement, distributedExample/02 "1 worker CPU").  A thin state holder over the C ABI's
    ``gaccum_host_session``: gradients go H2D every micro-step (double-buffered), the kernel runs,
    updated parameters come back D2H on apply steps, the 16-byte stats block every step.
    ``host_params`` are updated in place; pin them (and the gradients) for asynchronous copies."""

    @staticmethod
    def pinned_arena(shapes: Sequence[Sequence[int]], hp: HParams):
        """One pinned host buffer laid out like the device slabs (plan offsets, 128-byte aligned tensors) and the
        per-tensor views into it.  Gradients / parameters kept in such an arena cross PCIe with ONE copy per
        direction (gaccum_step_host coalesces tensors whose host spacing equals their slab spacing)."""
        numels = [int(np.prod(s)) for s in shapes]
        layout = Plan(numels, None, hp, device=-1)
        flat = torch.zeros(max(layout.padded_size, 32), dtype=torch.float32)
        try:
            flat = flat.pin_memory()
        except Exception:
            pass
        views = [flat[o:o + n].view(tuple(s)) for o, n, s in zip(layout.offsets, numels, shapes)]
        return flat, views

    def connect_data_parallel(self, process_group=None) -> None:
        """04's MultiWorkerMirroredStrategy for host-resident tensors: exchange the sessions' CUDA-IPC records over
        torch.distributed (any transport would do: they are plain bytes) and switch the apply step to the fused
        exchange + apply kernel.  Every rank must hold identical parameters and step in lock-step."""
        import torch.distributed as dist
        world, rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        if world < 2:
            return
        mine = self.session.dp_export(world)
        records = [None] * world
        dist.all_gather_object(records, mine, group=process_group)
        self.session.dp_connect(rank, world, records)
        dist.barrier(group=process_group)
        self.world = world

    def __init__(self, host_params: Sequence[torch.Tensor], names: Sequence[str], hp: HParams, accum_n: int,
                 lr_fn: Callable[[int], float],
                 exclude_from_weight_decay: Optional[Sequence[str]] = ("LayerNorm", "layer_norm", "bias"),
                 global_step: int = 0, device: int = 0):
        self.world = 1
        for p in host_params:
            if p.device.type != "cpu" or p.dtype != torch.float32 or not p.is_contiguous():
                raise ValueError("host_params must be contiguous fp32 CPU tensors")
        self.host_params = list(host_params)
        self.names, self.hp, self.N, self.lr_fn = list(names), hp, int(accum_n), lr_fn
        self.global_step = int(global_step)
        self.decay = _lib.decay_mask(self.names, hp.weight_decay_rate, exclude_from_weight_decay) \
            if hp.variant == _lib.ADAM_WEIGHT_DECAY else [False] * len(host_params)
        self.plan = Plan([p.numel() for p in host_params], self.decay, hp, device=device)
        self.session = _lib.HostSession(self.plan)
        self._param_ptrs = Plan.ptr_array([p.data_ptr() for p in self.host_params])
        self.session.set_params(self._param_ptrs)
        self.beta1_power, self.beta2_power = _f32(hp.beta1), _f32(hp.beta2)
        self.stats_host = torch.zeros(4, dtype=torch.float32)
        try:
            self.stats_host = self.stats_host.pin_memory()
        except Exception:
            pass

    def bind(self, host_grads: Sequence[Optional[torch.Tensor]]):
        for g, p in zip(host_grads, self.host_params):
            if g is not None and (g.device.type != "cpu" or g.dtype != torch.float32 or not g.is_contiguous() or g.numel() != p.numel()):
                raise ValueError("host gradients must be contiguous fp32 CPU tensors shaped like their parameter")
        return Plan.ptr_array([0 if g is None else g.data_ptr() for g in host_grads])

    def run(self, host_grads: Sequence[Optional[torch.Tensor]]) -> bool:
        return self.run_bound(self.bind(host_grads))

    def run_bound(self, grad_ptrs) -> bool:
        g = self.global_step
        lr = self.lr_fn(g)
        self.session.step(grad_ptrs, self._param_ptrs, StepArgs(g, self.N, 0, lr, self.beta1_power, self.beta2_power, 0.0),
                          self.stats_host.data_ptr())
        applied = _lib.is_apply_step(g, self.N)
        if applied and self.hp.variant == ADAM:
            self.beta1_power = _f32(np.float32(self.beta1_power) * np.float32(self.hp.beta1))
            self.beta2_power = _f32(np.float32(self.beta2_power) * np.float32(self.hp.beta2))
        self.global_step = g + 1
        return applied

    def sync(self) -> None:
        self.session.sync()

    def stats(self) -> Dict[str, float]:
        self.sync()
        a, lr, gn, s = self.stats_host.tolist()
        return {"applied": bool(a), "lr": lr, "global_norm": gn, "clip_scale": s}
