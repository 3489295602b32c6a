"""ctypes wrapper of oracle/libpgbart_oracle.so.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import
this module; nothing under pymc_bart_b200/ does.  PARITY UNPINNED: see the
header of pgbart_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from pymc_bart_b200 import _cabi
from pymc_bart_b200.settings import SamplerSettings

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpgbart_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pgbart_oracle.c")
    hdrs = [os.path.join(_HERE, "..", "include", h) for h in ("bk_spec.h", "pgbart_b200.h")]
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(f) for f in [src] + hdrs if os.path.exists(f))
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    lib.bko_create.argtypes = [C.POINTER(_cabi.BkSettings), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.bko_destroy.argtypes = [C.c_void_p]
    lib.bko_destroy.restype = None
    lib.bko_step.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p]
    lib.bko_sum_trees.argtypes = [C.c_void_p, C.c_void_p]
    lib.bko_read_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.bko_export_forest.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.bko_export_leaf_ids.argtypes = [C.c_void_p, C.c_void_p]
    lib.bko_export_leaf_values.argtypes = [C.c_void_p, C.c_void_p]
    lib.bko_bytes_touched.argtypes = [C.c_void_p]
    lib.bko_bytes_touched.restype = C.c_longlong
    lib.bko_set_threads.argtypes = [C.c_void_p, C.c_int]
    lib.bko_leaf_sd.argtypes = [C.c_void_p]
    lib.bko_leaf_sd.restype = C.c_float
    lib.bko_predict.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


class OracleChain:
    """One chain of the CPU restatement.  X is column-major [p][N] float32."""

    def __init__(self, settings: SamplerSettings, X_colmajor: np.ndarray, y: np.ndarray, chain: int = 0, group: int = 0):
        self.lib = load()
        self.settings = settings
        self.X = np.ascontiguousarray(X_colmajor, dtype=np.float32)
        self.y = np.ascontiguousarray(np.atleast_2d(np.asarray(y, dtype=np.float32)))   # [n_groups][N]
        assert self.X.shape == (settings.n_cols, settings.n_rows) and self.y.shape[1] == settings.n_rows
        self._cs = settings.to_c()
        h = C.c_void_p()
        rc = self.lib.bko_create(C.byref(self._cs), self.X.ctypes.data, self.y.ctypes.data, int(chain), int(group), C.byref(h))
        if rc != 0:
            raise RuntimeError(f"bko_create failed: {rc}")
        self.h = h
        self.N, self.p, self.m = settings.n_rows, settings.n_cols, settings.n_trees
        self.K = max(1, int(getattr(settings, "n_outputs", 1)))

    def set_threads(self, n: int) -> int:
        """Host threads for the particle loops of this chain (single-output step); results do not depend on it."""
        return int(self.lib.bko_set_threads(self.h, int(n)))

    def close(self):
        if getattr(self, "h", None):
            self.lib.bko_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, tune: bool, sigma: float = 1.0):
        vi = np.zeros(self.p, dtype=np.int32)
        st = _cabi.BkStepStats()
        rc = self.lib.bko_step(self.h, int(bool(tune)), float(sigma), vi.ctypes.data, C.byref(st))
        if rc != 0:
            raise RuntimeError(f"bko_step failed: {rc}")
        return vi, st

    def sum_trees(self) -> np.ndarray:
        """[N], or [K][N] for shared-tree multi-output."""
        out = np.empty(self.N if self.K == 1 else (self.K, self.N), dtype=np.float32)
        self.lib.bko_sum_trees(self.h, out.ctypes.data)
        return out

    def leaf_values(self) -> np.ndarray:
        """[m][255][K] leaf values of every output (0 for split nodes / unused slots)."""
        out = np.zeros((self.m, _cabi.BK_MAX_NODES, self.K), dtype=np.float32)
        self.lib.bko_export_leaf_values(self.h, out.ctypes.data)
        return out

    def trace(self) -> np.ndarray:
        cap = max(1, self.settings.trace_capacity)
        buf = np.zeros(cap, dtype=_cabi.TRACE_DTYPE)
        n = self.lib.bko_read_trace(self.h, buf.ctypes.data, cap)
        return buf[:n]

    def forest(self):
        nodes = np.zeros((self.m, _cabi.BK_MAX_NODES), dtype=_cabi.NODE_DTYPE)
        nn = np.zeros(self.m, dtype=np.int32)
        self.lib.bko_export_forest(self.h, nodes.ctypes.data, nn.ctypes.data)
        return nodes, nn

    def leaf_ids(self) -> np.ndarray:
        ids = np.zeros((self.m, self.N), dtype=np.uint8)
        self.lib.bko_export_leaf_ids(self.h, ids.ctypes.data)
        return ids

    def bytes_touched(self) -> int:
        return int(self.lib.bko_bytes_touched(self.h))


def predict(forests: np.ndarray, X_rowmajor: np.ndarray, draw_idx, excluded_mask=None, rules=None) -> np.ndarray:
    """forests: [n_draws][m][255] NODE_DTYPE; returns [n_idx][n] float32."""
    lib = load()
    forests = np.ascontiguousarray(forests, dtype=_cabi.NODE_DTYPE)
    X = np.ascontiguousarray(X_rowmajor, dtype=np.float32)
    di = np.ascontiguousarray(draw_idx, dtype=np.int32)
    n, p = X.shape
    m = forests.shape[1]
    out = np.empty((di.size, n), dtype=np.float32)
    ex = None if excluded_mask is None else np.ascontiguousarray(excluded_mask, dtype=np.uint8)
    ru = None if rules is None else np.ascontiguousarray(rules, dtype=np.int32)
    lib.bko_predict(forests.ctypes.data, None, m, X.ctypes.data, n, p, di.ctypes.data, di.size,
                    None if ex is None else ex.ctypes.data, None if ru is None else ru.ctypes.data, out.ctypes.data)
    return out
