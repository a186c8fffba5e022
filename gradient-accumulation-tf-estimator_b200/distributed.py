"""Data-parallel wiring of the train_op (reference distributedExample/04, MultiWorkerMirroredStrategy).

The reference marks every accumulator ``aggregation=SUM`` (04:55), so ``assign_add`` all-reduces
every gradient tensor on EVERY micro-step over a gRPC ring (04:58,70,106), and pre-divides the loss
by ``num_workers`` (04:46).  Summation is linear, so here each rank accumulates locally and the
packed accumulator slab is exchanged ONCE per window, on the apply step, over NCCL / NVLink:

    accumulate steps : local kernel only, no communication
    apply step       : a += G (local)  ->  all-reduce(a)  ->  apply kernel without a gradient

Every rank then runs the identical deterministic apply on identical inputs, so replicas stay
bit-identical without an extra norm exchange.  The producer keeps 04's convention of dividing the
loss by ``num_workers``.  (04's second reduction inside ``apply_gradients`` -- SURVEY.md 5.8 -- is
not reproduced.)
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.distributed as dist


class DataParallelTrainOp:
    """Wraps an engine exposing accumulate_only / apply_only / run / accum / N / global_step."""

    def __init__(self, engine, process_group=None):
        self.engine = engine
        self.group = process_group if process_group is not None and process_group is not True else None
        self.world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        self.allreduces = 0

    def run(self, grads: Sequence[Optional[torch.Tensor]]) -> bool:
        e = self.engine
        if self.world == 1:
            return e.run(grads)
        g = e.global_step
        if (g % e.N) != 0:                       # optimization.py:91 predicate, accumulate branch
            return e.run(grads)
        e.accumulate_only(grads)                 # 04:58
        dist.all_reduce(e.accum, op=dist.ReduceOp.SUM, group=self.group)   # 04:55, once per window
        self.allreduces += 1
        e.apply_only(None)                       # 04:59-66 on the summed accumulators
        e.global_step = g + 1                    # 04:74
        return True
