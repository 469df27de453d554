"""ctypes binding for oracle/libglove_oracle.so (the C port of the oracle).  TEST / BASELINE INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libglove_oracle.so")
_SO64 = os.path.join(_HERE, "libglove_oracle64.so")   # fp64 shadow: same source with -DORACLE_F64

HEADS = {"glove": 0, "logistic": 1}
OPTIMIZERS = {"Adam": 0, "Adagrad": 1, "SGD": 2}
ADAM_MODES = {"keras_dense": 0, "lazy": 1}

def _state_type(creal):
    fp = ctypes.POINTER(creal)

    class _State(ctypes.Structure):
        _fields_ = ([("V", ctypes.c_int32), ("d", ctypes.c_int32)] +
                    [(n, fp) for n in ("R", "C", "rb", "cb", "R_s0", "R_s1", "C_s0", "C_s1",
                                       "rb_s0", "rb_s1", "cb_s0", "cb_s1")] +
                    [("g", creal), ("g_s0", creal), ("g_s1", creal), ("step", ctypes.c_int32)])
    return _State, fp


_TYPES = {np.float32: (ctypes.c_float,) + _state_type(ctypes.c_float),
          np.float64: (ctypes.c_double,) + _state_type(ctypes.c_double)}


def build(force=False):
    src = os.path.join(_HERE, "glove_oracle.c")
    for so in (_SO, _SO64):
        if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s", os.path.basename(so)])
    return _SO


_libs = {}


def lib(dtype=np.float32):
    if dtype not in _libs:
        build()
        l = ctypes.CDLL(_SO if dtype == np.float32 else _SO64)
        l.glove_oracle_num_threads.restype = ctypes.c_int
        l.glove_oracle_train.restype = ctypes.c_int
        _libs[dtype] = l
    return _libs[dtype]


def num_threads():
    return lib().glove_oracle_num_threads()


class COracle:
    """Holds reference-layout tables (R, C [V,d]; rb, cb [V]) + optimizer slots as fp32 numpy arrays."""

    def __init__(self, R, C, rb, cb, g=0.0, optimizer="Adam", dtype=np.float32):
        """dtype=np.float64 runs the fp64 shadow (libglove_oracle64.so): same inputs (fp32 values), every variable and
        operation in double."""
        self.dtype = dtype
        self._creal, _State, _fp = _TYPES[dtype]
        self._fp = _fp
        self.R, self.C = np.ascontiguousarray(R, dtype).copy(), np.ascontiguousarray(C, dtype).copy()
        self.rb, self.cb = np.ascontiguousarray(rb, dtype).copy(), np.ascontiguousarray(cb, dtype).copy()
        init = float(np.float32(0.1)) if optimizer == "Adagrad" else 0.0   # Keras initial_accumulator_value (fp32 value)
        self.slots = {n + s: np.full_like(getattr(self, n), init) for n in ("R", "C", "rb", "cb") for s in ("_s0", "_s1")}
        self.optimizer = optimizer
        self.st = _State()
        self.st.V, self.st.d = self.R.shape
        for n in ("R", "C", "rb", "cb"):
            setattr(self.st, n, getattr(self, n).ctypes.data_as(_fp))
        for n, a in self.slots.items():
            setattr(self.st, n, a.ctypes.data_as(_fp))
        self.st.g = g
        self.st.g_s0 = init
        self.st.g_s1 = 0.0
        self.st.step = 0

    @property
    def g(self):
        return self.dtype(self.st.g)

    @property
    def step(self):
        return int(self.st.step)

    def train(self, coo, batch_idx, *, head="glove", learning_rate=0.001, l2_reg=0.01, reg_scale=2.0,
              neg_factor=1.0, adam_mode="keras_dense", alpha=None):
        batch_idx = np.ascontiguousarray(batch_idx, np.int64)
        n_steps, B = batch_idx.shape
        row = np.ascontiguousarray(coo["row"], np.int32)
        col = np.ascontiguousarray(coo["col"], np.int32)
        a_name, b_name = ("target", "weight") if head == "glove" else ("pos", "neg")
        _fp, creal, dt = self._fp, self._creal, self.dtype
        colA = np.ascontiguousarray(np.asarray(coo[a_name], np.float32), dt)     # fp32 VALUES in either precision
        colB = np.ascontiguousarray(np.asarray(coo[b_name], np.float32), dt)
        if alpha is None:
            from . import glove_oracle as _o
            alpha = _o.alpha_table(learning_rate, self.st.step + n_steps)
        alpha = np.ascontiguousarray(np.asarray(alpha, np.float32), dt)
        assert len(alpha) >= self.st.step + n_steps
        losses = np.zeros(n_steps, dt)
        f = lambda x: creal(float(np.float32(x)))
        rc = lib(dt).glove_oracle_train(
            ctypes.byref(self.st), row.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
            col.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), colA.ctypes.data_as(_fp), colB.ctypes.data_as(_fp),
            batch_idx.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), ctypes.c_int32(n_steps), ctypes.c_int32(B),
            ctypes.c_int32(HEADS[head]), ctypes.c_int32(OPTIMIZERS[self.optimizer]), f(learning_rate),
            f(l2_reg), f(reg_scale), f(neg_factor),
            ctypes.c_int32(ADAM_MODES[adam_mode]), alpha.ctypes.data_as(_fp), losses.ctypes.data_as(_fp))
        if rc != 0:
            raise RuntimeError("glove_oracle_train failed: %d" % rc)
        return losses
