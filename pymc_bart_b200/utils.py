"""Posterior-prediction glue and the variable-inclusion codec.

Mirrors pymc_bart/utils.py:26-130 (``_sample_posterior``, ``_MultiChainSampler``,
``_get_posterior_sampler``) and :1368-1398 (``_decode_vi`` / ``_encode_vi``); the native
``PosteriorSampler`` the reference gets from bartrs (pymc_bart/pymc_bart.py:2) is
implemented here on top of the CUDA kernel behind ``bk_predict``.
"""
from __future__ import annotations

import base64
import ctypes as C

import numpy as np

from . import _cabi


def _decode_vi(s: str, length: int) -> list[int]:
    """base64 -> unsigned LEB128 varints -> counts (pymc_bart/utils.py:1368-1384)."""
    data = base64.b64decode(s)
    out: list[int] = []
    i = 0
    while len(out) < length and i < len(data):
        num = shift = 0
        while i < len(data):
            byte = data[i]
            i += 1
            num |= (byte & 0x7F) << shift
            if not (byte & 0x80):
                break
            shift += 7
        out.append(num)
    return out


def _encode_vi(vec) -> str:
    """counts -> unsigned LEB128 varints -> base64 (pymc_bart/utils.py:1387-1398)."""
    buf = bytearray()
    for num in vec:
        n = int(num)
        while n > 127:
            buf.append((n & 0x7F) | 0x80)
            n >>= 7
        buf.append(n & 0x7F)
    return base64.b64encode(bytes(buf)).decode("ascii")


class PosteriorSampler:
    """Device-resident forest history of one chain + batched prediction.

    Same surface as the native class the reference shell calls
    (pymc_bart/utils.py:60-71,91,124-127): ``from_history``, ``n_draws``, ``n_outputs``,
    ``sample_posterior(X, draw_indices, excluded) -> (n_idx, n_outputs, n)``.
    """

    def __init__(self, forests: np.ndarray, n_outputs: int = 1, split_rules=None, device: int = 0):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("PosteriorSampler needs a CUDA device (no CPU fallback)")
        self.lib = _cabi.load()
        self.torch = torch
        self.device = torch.device("cuda", device)
        forests = np.ascontiguousarray(forests, dtype=_cabi.NODE_DTYPE)
        self._n_draws, self.m = forests.shape[0], forests.shape[1]
        self._n_outputs = int(n_outputs)
        self.forests_dev = torch.from_numpy(forests.view(np.uint8).reshape(-1)).to(self.device)
        self.rules_dev = None
        if split_rules is not None:
            self.rules_dev = torch.from_numpy(np.ascontiguousarray(split_rules, dtype=np.int32)).to(self.device)

    @staticmethod
    def rebuild_forests(batches, baseline_forest, m) -> np.ndarray:
        """Initial forest + per-draw deltas -> [n_draws][m][255] nodes (pure numpy)."""
        base_nodes, _ = baseline_forest
        cur = np.array(base_nodes, dtype=_cabi.NODE_DTYPE, copy=True)
        if cur.shape[0] != m:
            raise ValueError("baseline forest does not hold m trees")
        forests = np.zeros((len(batches), m, _cabi.BK_MAX_NODES), dtype=_cabi.NODE_DTYPE)
        for d, (first, nodes, _nn) in enumerate(batches):
            cur[first:first + nodes.shape[0]] = nodes
            forests[d] = cur
        return forests

    @classmethod
    def from_history(cls, batches, baseline_forest, m, n_outputs, split_rules=None, device: int = 0):
        """Rebuild per-draw forests from the initial forest plus per-draw deltas
        (pymc_bart/utils.py:124-127; CHANGELOG.md:23 "Better tree storage").  With n_outputs > 1
        (separate trees) `baseline_forest` and `batches` are lists with one entry per output group."""
        if n_outputs > 1:
            return _MultiOutputSampler([cls(cls.rebuild_forests(batches[g], baseline_forest[g], m), n_outputs=1,
                                            split_rules=split_rules, device=device) for g in range(n_outputs)])
        return cls(cls.rebuild_forests(batches, baseline_forest, m), n_outputs=n_outputs, split_rules=split_rules, device=device)

    @property
    def n_draws(self) -> int:
        return int(self._n_draws)

    @property
    def n_outputs(self) -> int:
        return self._n_outputs

    def sample_posterior(self, X, draw_indices, excluded=None) -> np.ndarray:
        torch = self.torch
        X = np.ascontiguousarray(np.asarray(X, dtype=np.float32))
        n, p = X.shape
        di = np.ascontiguousarray(np.asarray(draw_indices, dtype=np.int32))
        if di.size and (di.min() < 0 or di.max() >= self._n_draws):
            raise IndexError("draw index out of range")
        with torch.cuda.device(self.device):
            Xd = torch.from_numpy(X).to(self.device)
            dd = torch.from_numpy(di).to(self.device)
            out = torch.empty((di.size, n), dtype=torch.float32, device=self.device)
            ex_ptr = None
            if excluded is not None and len(excluded):
                mask = np.zeros(p, dtype=np.uint8)
                mask[np.asarray(list(excluded), dtype=np.int64)] = 1
                ex = torch.from_numpy(mask).to(self.device)
                ex_ptr = ex.data_ptr()
            stream = torch.cuda.current_stream(self.device)
            rc = self.lib.bk_predict(self.device.index, C.c_void_p(stream.cuda_stream), self.forests_dev.data_ptr(), None, self.m,
                                     Xd.data_ptr(), n, p, dd.data_ptr(), int(di.size), ex_ptr,
                                     None if self.rules_dev is None else self.rules_dev.data_ptr(), out.data_ptr())
            _cabi.check(rc, "bk_predict")
            res = out.cpu().numpy()
        return res.reshape(di.size, 1, n).astype(np.float64)


class _MultiOutputSampler:
    """k single-output samplers (separate trees) presented as one sampler with n_outputs = k."""

    def __init__(self, parts):
        self.parts = parts

    @property
    def n_draws(self):
        return self.parts[0].n_draws

    @property
    def n_outputs(self):
        return len(self.parts)

    def sample_posterior(self, X, draw_indices, excluded=None):
        return np.concatenate([p.sample_posterior(X, draw_indices, excluded) for p in self.parts], axis=1)


def _sample_posterior(sampler, X, rng, size=None, excluded=None):
    """pymc_bart/utils.py:26-71."""
    if size is None:
        size_iter = ()
    elif isinstance(size, int):
        size_iter = [size]
    else:
        size_iter = size
    flat = 1
    for s in size_iter:
        flat *= s
    X = np.ascontiguousarray(np.asarray(X, dtype=np.float64))
    excl = list(excluded) if excluded is not None else None
    first = sampler[0] if isinstance(sampler, list) else sampler
    draw_indices = rng.integers(0, first.n_draws, size=flat).tolist()
    if isinstance(sampler, list):
        pred = np.concatenate([s.sample_posterior(X, draw_indices, excl) for s in sampler], axis=1)
    else:
        pred = sampler.sample_posterior(X, draw_indices, excl)
    return pred.transpose((0, 2, 1)).reshape((*size_iter, -1, pred.shape[1]))


class _MultiChainSampler:
    """Routes each requested draw to the sampler of the chain it came from (pymc_bart/utils.py:74-107)."""

    def __init__(self, chain_samplers: list):
        if not chain_samplers:
            raise ValueError("No posterior draws available yet: run pm.sample() first.")
        self._chain_samplers = chain_samplers
        self._offsets = np.cumsum([0] + [s.n_draws for s in chain_samplers])

    @property
    def n_draws(self) -> int:
        return int(self._offsets[-1])

    @property
    def n_outputs(self) -> int:
        return self._chain_samplers[0].n_outputs

    def sample_posterior(self, X, draw_indices, excluded):
        draw_indices = np.asarray(draw_indices)
        chain_of_draw = np.searchsorted(self._offsets, draw_indices, side="right") - 1
        out = None
        for chain_idx, sampler in enumerate(self._chain_samplers):
            mask = chain_of_draw == chain_idx
            if not np.any(mask):
                continue
            local = (draw_indices[mask] - self._offsets[chain_idx]).tolist()
            preds = sampler.sample_posterior(X, local, excluded)
            if out is None:
                out = np.empty((len(draw_indices), *preds.shape[1:]), dtype=preds.dtype)
            out[mask] = preds
        return out


_posterior_sampler_cache: dict = {}


def _get_posterior_sampler(op) -> _MultiChainSampler:
    """pymc_bart/utils.py:113-130."""
    n_chains = len(op.all_trees)
    cached = _posterior_sampler_cache.get(id(op))
    if cached is not None and cached[0] == n_chains:
        return cached[1]
    from .settings import SPLIT_RULE_CODES

    rules = None
    if getattr(op, "split_rules", None) is not None:
        rules = np.array([SPLIT_RULE_CODES[r if (r is None or isinstance(r, str)) else type(r).__name__] for r in op.split_rules], dtype=np.int32)
    chain_samplers = [
        PosteriorSampler.from_history(batches, baseline_forest, op.m, op.n_outputs, split_rules=rules)
        for baseline_forest, batches in op.all_trees
    ]
    sampler = _MultiChainSampler(chain_samplers)
    _posterior_sampler_cache[id(op)] = (n_chains, sampler)
    return sampler


def get_variable_inclusion_counts(stats, n_cols: int) -> np.ndarray:
    """Sum of the decoded per-draw counts (what pymc_bart/utils.py:778-790 computes from idata)."""
    tot = np.zeros(n_cols, dtype=np.int64)
    for s in stats:
        tot += np.asarray(_decode_vi(s["variable_inclusion"], n_cols), dtype=np.int64)
    return tot
