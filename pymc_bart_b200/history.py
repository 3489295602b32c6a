"""Forest history: the ``(baseline_forest, batches)`` format published in ``op.all_trees`` and its device store.

Reference: bartrs' native ``PosteriorSampler`` as the shell uses it — ``from_history(batches, baseline_forest, m,
n_outputs)`` (pymc_bart/utils.py:124-127), ``n_draws`` / ``n_outputs`` (:63,83,91), ``sample_posterior(X, draw_indices,
excluded)`` (:67,69,103) — and ``_MultiChainSampler`` (:74-107).  "Better tree storage" (CHANGELOG.md:23): a chain's
history is its initial forest plus, per draw, only the trees that draw rewrote.

Formats (plain numpy, picklable through the ``multiprocessing.Manager`` proxy of pymc_bart/bart.py:134-135):

* ``baseline_forest = (nodes, n_nodes)``: ``n_nodes`` int32 ``[G*m]`` (output group major), ``nodes`` the trees' nodes
  back to back (NODE_DTYPE, 24 bytes each, ``n_nodes`` of them per tree);
* one batch per post-tuning draw: ``(first, n_nodes, nodes)`` with ``n_nodes`` int32 ``[G, T]`` for the trees
  ``first .. first+T-1`` of every output group, nodes back to back in that order;
* shared-tree multi-output (every leaf carries K values): both tuples end with one more array, the nodes' leaf
  values ``[n_nodes_total][K]`` float32 (G = 1 then; the K outputs come from one tree walk).

Device store (``DeviceForests``): every tree version once, plus a table ``[forest][tree] -> version``; a draw costs
``m`` ints (C5: 800 bytes instead of a 1.2 MB dense forest).  Forest rows are ordered (chain, draw, group): the
global draw index of the multi-chain sampler addresses the table directly.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _cabi


def compact_forest(nodes: np.ndarray, n_nodes: np.ndarray):
    """Dense ``[k][255]`` nodes + counts -> (flat nodes, counts)."""
    n_nodes = np.ascontiguousarray(n_nodes, dtype=np.int32).reshape(-1)
    nodes = np.asarray(nodes, dtype=_cabi.NODE_DTYPE).reshape(n_nodes.size, -1)
    flat = np.concatenate([nodes[t, : n_nodes[t]] for t in range(n_nodes.size)]) if n_nodes.size else np.zeros(0, _cabi.NODE_DTYPE)
    return np.ascontiguousarray(flat), n_nodes


class ChainHistory:
    """Host description of one chain's history: tree versions and the per-draw version table."""

    def __init__(self, batches, baseline_forest, m: int, n_outputs: int):
        base_nodes, base_nn = baseline_forest[0], baseline_forest[1]
        base_vals = baseline_forest[2] if len(baseline_forest) > 2 else None
        self.K = 1 if base_vals is None else int(np.asarray(base_vals).shape[1])    # shared-tree outputs per leaf
        G = 1 if self.K > 1 else int(n_outputs)
        if self.K > 1 and self.K != int(n_outputs):
            raise ValueError("baseline leaf values do not hold n_outputs columns")
        base_nn = np.ascontiguousarray(base_nn, dtype=np.int32).reshape(-1)
        if base_nn.size != G * m:
            raise ValueError("baseline forest does not hold n_outputs * m trees")
        batches = list(batches)
        self.m, self.G, self.n_draws = int(m), G, len(batches)
        nn_parts = [base_nn] + [np.ascontiguousarray(b[1], dtype=np.int32).reshape(-1) for b in batches]
        node_parts = [np.asarray(base_nodes, dtype=_cabi.NODE_DTYPE)] + [np.asarray(b[2], dtype=_cabi.NODE_DTYPE) for b in batches]
        self.ver_nn = np.concatenate(nn_parts)
        self.nodes = np.concatenate(node_parts) if node_parts else np.zeros(0, _cabi.NODE_DTYPE)
        if int(self.ver_nn.sum()) != self.nodes.size:
            raise ValueError("history node counts do not add up")
        self.vals = None
        if self.K > 1:
            self.vals = np.ascontiguousarray(np.concatenate([np.asarray(base_vals, dtype=np.float32)] +
                                                            [np.asarray(b[3], dtype=np.float32) for b in batches]))
            if self.vals.shape != (self.nodes.size, self.K):
                raise ValueError("history leaf values do not match the nodes")
        # version table: forest row (draw, group) -> version of every tree
        tbl = np.empty((self.n_draws, G, m), dtype=np.int32)
        cur = np.arange(G * m, dtype=np.int32).reshape(G, m)
        nxt = G * m
        for d, b in enumerate(batches):
            first, nn = int(b[0]), np.asarray(b[1])
            T = nn.reshape(G, -1).shape[1]
            cur[:, first:first + T] = nxt + np.arange(G * T, dtype=np.int32).reshape(G, T)
            nxt += G * T
            tbl[d] = cur
        self.ver_tbl = tbl.reshape(self.n_draws * G, m)

    def forest_sizes(self) -> np.ndarray:
        return self.ver_nn[self.ver_tbl].sum(axis=1) if self.ver_tbl.size else np.zeros(0, np.int64)

    def dense_forests(self) -> np.ndarray:
        """[n_draws*G][m][255] dense nodes (a test helper: the CPU checker predicts from dense forests; O(draws*m*255) memory)."""
        off = np.concatenate([[0], np.cumsum(self.ver_nn)])
        out = np.zeros((self.ver_tbl.shape[0], self.m, _cabi.BK_MAX_NODES), dtype=_cabi.NODE_DTYPE)
        for f in range(self.ver_tbl.shape[0]):
            for t in range(self.m):
                v = self.ver_tbl[f, t]
                out[f, t, : self.ver_nn[v]] = self.nodes[off[v]: off[v + 1]]
        return out


class DeviceForests:
    """All chains of an op on the device: nodes, version offsets, version table; one launch per prediction call."""

    def __init__(self, chains: list, split_rules=None, device: int = 0, subset_tables=None):
        import torch

        self.subset_tables = subset_tables or {}   # SubsetSplit columns: value -> category code tables of the training data

        if not torch.cuda.is_available():
            raise RuntimeError("posterior prediction needs a CUDA device (no CPU fallback)")
        if not chains:
            raise ValueError("No posterior draws available yet: run pm.sample() first.")
        self.lib = _cabi.load()
        self.torch = torch
        self.device = torch.device("cuda", device)
        self.m, self.G, self.K = chains[0].m, chains[0].G, chains[0].K
        nodes, nn, tbl, vals, voff = [], [], [], [], 0
        for ch in chains:
            if ch.m != self.m or ch.G != self.G or ch.K != self.K:
                raise ValueError("chains disagree on m / n_outputs")
            nodes.append(ch.nodes); nn.append(ch.ver_nn); tbl.append(ch.ver_tbl + voff)
            if self.K > 1:
                vals.append(ch.vals)
            voff += ch.ver_nn.size
        ver_nn = np.concatenate(nn)
        ver_off = np.zeros(ver_nn.size + 1, dtype=np.int64)
        np.cumsum(ver_nn, out=ver_off[1:])
        if ver_off[-1] >= 2**31:
            raise ValueError("forest history exceeds 2^31 nodes")
        ver_tbl = np.ascontiguousarray(np.concatenate(tbl), dtype=np.int32)
        self.n_draws_per_chain = [ch.n_draws for ch in chains]
        self.n_draws = int(sum(self.n_draws_per_chain))
        self.max_forest_nodes = int(ver_nn[ver_tbl].sum(axis=1).max()) if ver_tbl.size else 0
        self.history_bytes = int(ver_off[-1]) * _cabi.NODE_DTYPE.itemsize + ver_tbl.nbytes + ver_off.size * 4
        with torch.cuda.device(self.device):
            self.nodes_dev = torch.from_numpy(np.concatenate(nodes).view(np.uint8).reshape(-1).copy()).to(self.device)
            self.ver_off_dev = torch.from_numpy(ver_off.astype(np.int32)).to(self.device)
            self.ver_tbl_dev = torch.from_numpy(ver_tbl).to(self.device)
            self.vals_dev = torch.from_numpy(np.ascontiguousarray(np.concatenate(vals))).to(self.device) if self.K > 1 else None
            self.rules_dev = None
            if split_rules is not None:
                self.rules_dev = torch.from_numpy(np.ascontiguousarray(split_rules, dtype=np.int32)).to(self.device)
            self.err_dev = torch.zeros(1, dtype=torch.int32, device=self.device)

    def upload(self, X):
        """Row-major float32 copy of X on the device (upload once, predict many times)."""
        torch = self.torch
        if isinstance(X, torch.Tensor):
            return X
        if self.subset_tables:
            from .settings import encode_subset_columns

            X = encode_subset_columns(np.asarray(X), self.subset_tables)
        Xh = np.ascontiguousarray(np.asarray(X, dtype=np.float32))
        if Xh.ndim != 2:
            raise ValueError("X must be two-dimensional")
        return torch.from_numpy(Xh).to(self.device)

    def predict(self, X, draw_indices, masks=None, per_mask=False):
        """Sum of trees of the forests of the global draws ``draw_indices`` at the rows of X.

        masks: None or uint8 ``[n_masks][p]`` (1 = excluded variable); per_mask: ``draw_indices`` is ``[n_masks][S]``,
        one set of draws per mask.  Returns a device tensor ``[n_masks or 1][S][n_outputs][n]`` float32 (n_outputs =
        separate-tree groups, or the K values of shared-tree leaves)."""
        torch = self.torch
        Xd = self.upload(X)
        n, p = int(Xd.shape[0]), int(Xd.shape[1])
        n_masks = 0 if masks is None else int(np.asarray(masks).shape[0])
        di = np.asarray(draw_indices, dtype=np.int64)
        if per_mask:
            if di.ndim != 2 or di.shape[0] != n_masks:
                raise ValueError("per_mask needs draw indices of shape [n_masks][S]")
        else:
            di = di.reshape(1, -1)
        if di.size and (di.min() < 0 or di.max() >= self.n_draws):
            raise IndexError("draw index out of range")
        S = di.shape[1]
        # forest rows: (draw, group) -> draw * G + group, per mask when the draws differ per mask
        sel2 = (di[:, :, None] * self.G + np.arange(self.G)[None, None, :]).reshape(di.shape[0], S * self.G).astype(np.int32)
        sel = sel2.reshape(-1)
        n_sel = S * self.G
        with torch.cuda.device(self.device):
            sel_dev = torch.from_numpy(sel).to(self.device)
            masks_dev = None
            if n_masks:
                mk = np.ascontiguousarray(masks, dtype=np.uint8)
                if mk.shape != (n_masks, p):
                    raise ValueError("masks must have shape [n_masks][p]")
                masks_dev = torch.from_numpy(mk).to(self.device)
            self.err_dev.zero_()
            stream = torch.cuda.current_stream(self.device)
            if n_sel > 65535:
                raise ValueError("at most 65535 (draw, output) forests per prediction call")
            out = torch.empty((max(n_masks, 1), n_sel, self.K, n), dtype=torch.float32, device=self.device)
            if n_sel and n:
                rc = self.lib.bk_predict_history(
                    self.device.index, C.c_void_p(stream.cuda_stream), self.nodes_dev.data_ptr(), self.ver_off_dev.data_ptr(),
                    self.ver_tbl_dev.data_ptr(), self.m, self.max_forest_nodes, Xd.data_ptr(), n, p,
                    sel_dev.data_ptr(), n_sel, int(bool(per_mask)), None if masks_dev is None else masks_dev.data_ptr(), n_masks,
                    None if self.rules_dev is None else self.rules_dev.data_ptr(),
                    None if self.vals_dev is None else self.vals_dev.data_ptr(), self.K, out.data_ptr(), self.err_dev.data_ptr())
                _cabi.check(rc, "bk_predict_history")
            if int(self.err_dev.item()) != 0:
                raise RuntimeError("posterior prediction: device-side consistency flag set (malformed forest history)")
        return out.reshape(max(n_masks, 1), S, self.G * self.K, n)

    def pearson_r2(self, a, b):
        """Squared Pearson correlation per (subset, sample): a ``[S][len]``, b ``[K][S][len]`` device tensors
        (pymc_bart/utils.py:1339-1346) -> numpy ``[K][S]`` float64."""
        torch = self.torch
        a = a.contiguous(); b = b.contiguous()
        S, ln = int(a.shape[0]), int(a[0].numel())
        K = int(b.shape[0])
        with torch.cuda.device(self.device):
            out = torch.empty((K, S), dtype=torch.float64, device=self.device)
            stream = torch.cuda.current_stream(self.device)
            _cabi.check(self.lib.bk_pearson_r2(self.device.index, C.c_void_p(stream.cuda_stream), a.data_ptr(), b.data_ptr(), ln, S, K,
                                               out.data_ptr()), "bk_pearson_r2")
            return out.cpu().numpy()
