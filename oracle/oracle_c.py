"""ctypes front-end of oracle/oracle.c (TEST INFRASTRUCTURE ONLY -- see that file's header).

Same interface as ``oracle_np.ReferenceTrainOp`` so tests can run the two restatements side by
side, and so bench.py can time "the reference's CPU path" on all host cores.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence

import numpy as np

from oracle_np import (DEFAULT_EXCLUDE, HParams, StepInfo, do_use_weight_decay,
                       get_variable_name)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-B", "liboracle.so"], cwd=_HERE)
    return _LIB_PATH


class _HP(C.Structure):
    _fields_ = [("variant", C.c_int32), ("_pad", C.c_int32), ("beta1", C.c_double),
                ("beta2", C.c_double), ("epsilon", C.c_double), ("weight_decay_rate", C.c_double),
                ("clip_norm", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_learning_rate.restype = C.c_float
        _lib.oracle_learning_rate.argtypes = [C.c_double, C.c_int64, C.c_int64, C.c_int64]
        _lib.oracle_l2_loss.restype = C.c_float
        _lib.oracle_l2_loss.argtypes = [C.c_void_p, C.c_int64]
        _lib.oracle_clip_scale.restype = C.c_float
        _lib.oracle_clip_scale.argtypes = [C.c_float, C.c_float]
        _lib.oracle_step.restype = C.c_int
        _lib.oracle_step.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(_HP), C.c_int64,
                                     C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_num_threads.restype = C.c_int
        _lib.oracle_set_num_threads.argtypes = [C.c_int]
    return _lib


def learning_rate(init_lr, num_train_steps, num_warmup_steps, global_step) -> np.float32:
    return np.float32(lib().oracle_learning_rate(float(init_lr), int(num_train_steps),
                                                 int(num_warmup_steps or 0), int(global_step)))


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(int(n))


def _ptr_array(arrs: Sequence[Optional[np.ndarray]]):
    out = (C.c_void_p * len(arrs))()
    for i, a in enumerate(arrs):
        out[i] = None if a is None else a.ctypes.data
    return out


class COracleTrainOp:
    def __init__(self, params: List[np.ndarray], names: Sequence[str], hp: HParams, accum_n: int,
                 init_lr: float = 0.0, num_train_steps: int = 1, num_warmup_steps: Optional[int] = 0,
                 constant_lr: Optional[float] = None, exclude=DEFAULT_EXCLUDE, global_step: int = 0):
        self.params = [np.ascontiguousarray(p, dtype=np.float32) for p in params]
        self.names = [get_variable_name(n) for n in names]
        self.hp, self.N = hp, int(accum_n)
        self.init_lr, self.num_train_steps, self.num_warmup_steps = init_lr, num_train_steps, num_warmup_steps
        self.constant_lr = constant_lr
        self.global_step = int(global_step)
        self.accum = [np.zeros_like(p) for p in self.params]
        self.m = [np.zeros_like(p) for p in self.params]
        self.v = [np.zeros_like(p) for p in self.params]
        self.decay = np.array([do_use_weight_decay(n, hp.weight_decay_rate, exclude) for n in self.names],
                              dtype=np.uint8)
        self.beta_pow = np.array([hp.beta1, hp.beta2], dtype=np.float32)
        self._numel = np.array([p.size for p in self.params], dtype=np.int64)
        self._scratch = np.empty(6 * int(self._numel.max(initial=1)), dtype=np.float32)
        self._hp = _HP(hp.variant, 0, hp.beta1, hp.beta2, hp.epsilon, hp.weight_decay_rate,
                       hp.clip_norm if hp.clip_norm else 0.0)
        self._info = np.zeros(4, dtype=np.float32)
        self._pp, self._pa = _ptr_array(self.params), _ptr_array(self.accum)
        self._pm, self._pv = _ptr_array(self.m), _ptr_array(self.v)

    def lr(self, g: int) -> np.float32:
        if self.constant_lr is not None:
            return np.float32(self.constant_lr)
        return learning_rate(self.init_lr, self.num_train_steps, self.num_warmup_steps, g)

    @property
    def beta1_power(self):
        return self.beta_pow[0]

    @property
    def beta2_power(self):
        return self.beta_pow[1]

    def run(self, grads: Sequence[Optional[np.ndarray]]) -> StepInfo:
        g = self.global_step
        keep = [None if x is None else np.ascontiguousarray(x, dtype=np.float32) for x in grads]
        lr = self.lr(g)
        rc = lib().oracle_step(len(self.params), self._numel.ctypes.data, self._pp, _ptr_array(keep),
                               self._pa, self._pm, self._pv, self.decay.ctypes.data, C.byref(self._hp),
                               g, self.N, float(lr), self.beta_pow.ctypes.data,
                               self._scratch.ctypes.data, self._info.ctypes.data)
        if rc != 0:
            raise RuntimeError("oracle_step failed")
        self.global_step = g + 1
        return StepInfo(global_step=g, applied=bool(self._info[0]), lr=np.float32(self._info[1]),
                        global_norm=np.float32(self._info[2]), clip_scale=np.float32(self._info[3]))
