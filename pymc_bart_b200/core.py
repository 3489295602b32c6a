"""Device-resident sampler core: torch tensors own HBM, the C ABI drives the kernels.

Host side of SURVEY.md §8b's boundary: this object is what the PGBART step class
holds per GPU.  X (column-major, padded), y, the sum-of-trees matrix and one
workspace blob are torch CUDA tensors; libpgbart_b200.so borrows their pointers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi
from .settings import SamplerSettings


class DeviceSampler:
    def __init__(self, settings: SamplerSettings, X: np.ndarray, Y: np.ndarray):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("pymc_bart_b200 needs a CUDA device (no CPU fallback)")
        self.lib = _cabi.load()
        self.torch = torch
        self.settings = settings
        self.device = torch.device("cuda", settings.device)
        G = max(1, settings.n_groups)
        K = max(1, getattr(settings, "n_outputs", 1))                          # leaf values per leaf (shared trees)
        N, p, Cn = settings.n_rows, settings.n_cols, settings.n_chains * G   # Cn: chains x output groups
        self.N, self.p, self.m, self.C, self.G, self.K = N, p, settings.n_trees, Cn, G, K
        self.rows = Cn * K                                                     # rows of the sum-of-trees matrix
        self.ld = int(self.lib.bk_padded_rows(N))
        # host staging in pinned memory, column-major fp32 (bart.py:209-210 hands over f64 row-major)
        Xh = torch.zeros((p, self.ld), dtype=torch.float32).pin_memory()
        Xh[:, :N] = torch.from_numpy(np.ascontiguousarray(np.asarray(X, dtype=np.float32).T))
        yh = torch.zeros((G, self.ld), dtype=torch.float32).pin_memory()
        Y2 = np.atleast_2d(np.asarray(Y, dtype=np.float32))
        if Y2.shape != (G, N):
            raise ValueError(f"Y must have shape ({G}, {N}) for {G} output groups")
        yh[:, :N] = torch.from_numpy(np.ascontiguousarray(Y2))
        self.h2d_bytes = Xh.numel() * 4 + yh.numel() * 4
        with torch.cuda.device(self.device):
            self.X_dev = Xh.to(self.device, non_blocking=True)
            self.y_dev = yh.to(self.device, non_blocking=True)
            self.sum_trees_dev = torch.empty((Cn * K, self.ld), dtype=torch.float32, device=self.device)
            self._cs = settings.to_c()
            nbytes = C.c_size_t()
            _cabi.check(self.lib.bk_query_bytes(C.byref(self._cs), C.byref(nbytes)), "bk_query_bytes")
            self.workspace_bytes = int(nbytes.value)
            self.workspace = torch.empty((self.workspace_bytes,), dtype=torch.uint8, device=self.device)
            torch.cuda.synchronize(self.device)
            h = C.c_void_p()
            _cabi.check(
                self.lib.bk_create(C.byref(self._cs), self.X_dev.data_ptr(), self.y_dev.data_ptr(),
                                   self.sum_trees_dev.data_ptr(), self.workspace.data_ptr(), C.byref(h)),
                "bk_create",
            )
        self.h = h
        self._vi = np.zeros((Cn, p), dtype=np.int32)
        self._stats = (_cabi.BkStepStats * Cn)()
        self._sigma = np.ones(Cn, dtype=np.float32)
        self._host_out = None
        self._host_enabled = False
        self._run_n = []        # steps of the launches in flight (run_launch/run_wait), oldest first

    def set_response(self, Y):
        """Replace the response the steps see (rows [G][N]) — for models in which this BART variable is one term of the
        likelihood's location (several BART variables in one model, tests/test_bart.py:167-241: the step of one
        variable sees `observed - the other terms` at the current point).  One small H2D copy from the library's pinned
        staging buffer on the sampler's stream, ordered before the next launch; the kernels read `y` afresh in every step."""
        Y2 = np.ascontiguousarray(np.atleast_2d(np.asarray(Y, dtype=np.float32)))
        if Y2.shape != (self.G, self.N):
            raise ValueError(f"the response must have shape ({self.G}, {self.N})")
        if self._run_n:
            raise RuntimeError("set_response with launches in flight")
        _cabi.check(self.lib.bk_set_response(self.h, Y2.ctypes.data), "bk_set_response")

    def close(self):
        if getattr(self, "h", None):
            self.lib.bk_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def step(self, tune: bool, sigma=1.0):
        """One PGBART step of every chain; returns (vi_counts [C,p] int32, stats)."""
        self._sigma[...] = np.asarray(sigma, dtype=np.float32)
        _cabi.check(
            self.lib.bk_step(self.h, int(bool(tune)), self._sigma.ctypes.data, self._vi.ctypes.data,
                             C.cast(self._stats, C.c_void_p)),
            "bk_step",
        )
        return self._vi, self._stats

    def step_launch(self, tune: bool, sigma=1.0):
        """Enqueue one step on the sampler's stream (returns immediately)."""
        self._sigma[...] = np.asarray(sigma, dtype=np.float32)
        _cabi.check(self.lib.bk_step_launch(self.h, int(bool(tune)), self._sigma.ctypes.data), "bk_step_launch")

    def step_wait(self):
        _cabi.check(self.lib.bk_step_wait(self.h, self._vi.ctypes.data, C.cast(self._stats, C.c_void_p)), "bk_step_wait")
        return self._vi, self._stats

    # ---- several steps per launch (fixed likelihood parameters between them) -----------------------------------------
    MAX_STEPS_PER_LAUNCH = 16

    def run_launch(self, n_steps: int, tune: bool, sigma=1.0, draws_out=None):
        """Enqueue n_steps steps of every chain as ONE launch.  draws_out: None or a float32 CUDA tensor
        [n_steps, C*K, ld] that receives the sum of trees after every step."""
        self._sigma[...] = np.asarray(sigma, dtype=np.float32)
        ptr = None
        if draws_out is not None:
            if tuple(draws_out.shape) != (int(n_steps), self.rows, self.ld) or draws_out.dtype != self.torch.float32 or not draws_out.is_contiguous():
                raise ValueError(f"draws_out must be a contiguous float32 tensor of shape ({n_steps}, {self.rows}, {self.ld})")
            ptr = draws_out.data_ptr()
        _cabi.check(self.lib.bk_run_launch(self.h, int(n_steps), int(bool(tune)), self._sigma.ctypes.data, ptr), "bk_run_launch")
        self._run_n.append(int(n_steps))      # (up to two launches are in flight: the waits come in launch order)

    def run_wait(self):
        """(vi counts [n_steps][C][p] int32, stats [n_steps][C]) of the oldest launch in flight."""
        if not self._run_n:
            raise RuntimeError("run_wait without a launch in flight")
        n = self._run_n.pop(0)
        vi = np.zeros((n, self.C, self.p), dtype=np.int32)
        stats = (_cabi.BkStepStats * (n * self.C))()
        _cabi.check(self.lib.bk_run_wait(self.h, vi.ctypes.data, C.cast(stats, C.c_void_p)), "bk_run_wait")
        return vi, [stats[i * self.C:(i + 1) * self.C] for i in range(n)]

    def set_draw_peers(self, peer_ptrs, local_ptr):
        """The draws a launch writes into `draws_out` (a view into the buffer at `local_ptr`) also go to the same place of
        the peers' buffers (device pointers as mapped in this process), by P2P stores from the commit sweep."""
        n = len(peer_ptrs)
        arr = (C.c_void_p * max(n, 1))(*[C.c_void_p(int(p)) for p in peer_ptrs])
        _cabi.check(self.lib.bk_set_draw_peers(self.h, n, C.cast(arr, C.c_void_p), C.c_void_p(int(local_ptr))), "bk_set_draw_peers")

    def stream(self):
        """torch view of the sampler's CUDA stream (for events and ordered copies)."""
        return self.torch.cuda.ExternalStream(int(self.lib.bk_stream(self.h)), device=self.device)

    # ---- tree history (pymc_bart/utils.py:117-127) -------------------------------------------------------------------
    def enable_history(self, enable: bool = True, steps_per_launch: int = 1):
        """Post-tuning steps also copy the trees they rewrote to pinned host memory behind the kernel (buffers sized
        for launches of `steps_per_launch` steps)."""
        _cabi.check(self.lib.bk_set_history(self.h, max(1, int(steps_per_launch)) if enable else 0), "bk_set_history")
        T = max(self.settings.batch_tune, self.settings.batch_post)
        self._hist_nn = np.zeros((self.C, T), dtype=np.int32)
        self._hist_nodes = np.zeros(self.C * T * _cabi.BK_MAX_NODES, dtype=_cabi.NODE_DTYPE)
        self._hist_vals = np.zeros((self.C * T * _cabi.BK_MAX_NODES, self.K), dtype=np.float32) if self.K > 1 else None
        self.history_bytes_per_step = self.C * self.settings.batch_post * (_cabi.BK_MAX_NODES * 64 + 4)

    def history_batch(self, step: int = 0):
        """Trees rewritten by step `step` of the last launch waited for: (first, n_nodes [C][T], nodes back to back[,
        leaf values [nodes][K] for shared-tree multi-output]) or None."""
        first, total = C.c_int32(), C.c_int64()
        T = self.lib.bk_history_batch_at(self.h, int(step), C.byref(first), self._hist_nn.ctypes.data, self._hist_nodes.ctypes.data,
                                         C.byref(total))
        if T < 0:
            _cabi.check(T, "bk_history_batch_at")
        if T == 0:
            return None
        nn = self._hist_nn.reshape(-1)[: self.C * T].reshape(self.C, T).copy()
        if self.K > 1:
            rc = self.lib.bk_history_values_at(self.h, int(step), self._hist_vals.ctypes.data)
            if rc < 0:
                _cabi.check(rc, "bk_history_values")
            return int(first.value), nn, self._hist_nodes[: int(total.value)].copy(), self._hist_vals[: int(total.value)].copy()
        return int(first.value), nn, self._hist_nodes[: int(total.value)].copy()

    def baseline(self):
        """Current forest of every (chain, group), compacted: list of (nodes, n_nodes) per virtual chain."""
        from .history import compact_forest

        out = []
        for c in range(self.C):
            nodes, nn = self.forest(c)
            flat, nn = compact_forest(nodes, nn)
            if self.K > 1:      # shared-tree multi-output: the nodes' K leaf values travel beside them
                vals = self.leaf_values(c)
                out.append((flat, nn, np.concatenate([vals[t, : nn[t]] for t in range(self.m)])))
            else:
                out.append((flat, nn))
        return out

    def leaf_values(self, chain: int = 0) -> np.ndarray:
        """[m][255][K] leaf values of every output of a chain's current forest."""
        vals = np.zeros((self.m, _cabi.BK_MAX_NODES, self.K), dtype=np.float32)
        _cabi.check(self.lib.bk_export_leaf_values(self.h, int(chain), vals.ctypes.data), "bk_export_leaf_values")
        return vals

    def trees(self, chain: int, first: int, count: int):
        nodes = np.zeros((count, _cabi.BK_MAX_NODES), dtype=_cabi.NODE_DTYPE)
        nn = np.zeros(count, dtype=np.int32)
        _cabi.check(self.lib.bk_export_trees(self.h, int(chain), int(first), int(count), nodes.ctypes.data, nn.ctypes.data), "bk_export_trees")
        return nodes, nn

    def sum_trees(self):
        """torch view [C*K, N] of the current sum of trees (device); rows are (chain, group or output)."""
        return self.sum_trees_dev[:, : self.N]

    def enable_host_output(self, enable: bool = True):
        """Every step also copies the sum of trees to pinned host memory behind the kernel (the value handed to PyMC)."""
        _cabi.check(self.lib.bk_set_host_output(self.h, int(bool(enable))), "bk_set_host_output")
        self._host_enabled = bool(enable)

    def sum_trees_host(self) -> np.ndarray:
        """Host view [C, N] of the sum of trees after the last step: the library's pinned buffer when host output is
        enabled (no extra copy), otherwise a blocking D2H copy."""
        if getattr(self, "_host_enabled", False):   # (two pinned buffers alternate: ask for the last step's every time)
            return np.ctypeslib.as_array(self.lib.bk_sum_trees_host(self.h), shape=(self.rows, self.N))
        torch = self.torch
        if self._host_out is None:
            self._host_out = torch.empty((self.rows, self.N), dtype=torch.float32).pin_memory()
        self._host_out.copy_(self.sum_trees_dev[:, : self.N], non_blocking=False)
        return self._host_out.numpy()

    def trace(self, chain: int = 0) -> np.ndarray:
        cap = max(1, self.settings.trace_capacity)
        buf = np.zeros(cap, dtype=_cabi.TRACE_DTYPE)
        n = self.lib.bk_read_trace(self.h, int(chain), buf.ctypes.data, cap)
        if n < 0:
            _cabi.check(n, "bk_read_trace")
        return buf[:n]

    def forest(self, chain: int = 0):
        nodes = np.zeros((self.m, _cabi.BK_MAX_NODES), dtype=_cabi.NODE_DTYPE)
        nn = np.zeros(self.m, dtype=np.int32)
        _cabi.check(self.lib.bk_export_forest(self.h, int(chain), nodes.ctypes.data, nn.ctypes.data), "bk_export_forest")
        return nodes, nn

    def leaf_ids(self, chain: int = 0) -> np.ndarray:
        ids = np.zeros((self.m, self.N), dtype=np.uint8)
        _cabi.check(self.lib.bk_export_leaf_ids(self.h, int(chain), ids.ctypes.data), "bk_export_leaf_ids")
        return ids
