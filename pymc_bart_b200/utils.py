"""Posterior-prediction glue and the variable-inclusion codec.

Same names and call contracts as pymc_bart/utils.py:26-130 (``_sample_posterior``, ``_MultiChainSampler``,
``_get_posterior_sampler``) and :1368-1398 (``_decode_vi`` / ``_encode_vi``), so that ``BARTRV.rng_fn`` and the
analytics call them unchanged; what sits behind them is different.  The reference keeps one native
``PosteriorSampler`` per chain (bartrs, pymc_bart/pymc_bart.py:2) and routes every requested draw to its chain in a
Python loop.  Here ALL chains of an op live in one device store of tree versions (``history.DeviceForests``): a global
draw index addresses the version table directly, so a prediction call is one upload of X and one kernel launch.
"""
from __future__ import annotations

import base64

import numpy as np

from .history import ChainHistory, DeviceForests


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


def _rule_codes(split_rules):
    if split_rules is None:
        return None
    from .settings import SPLIT_RULE_CODES

    return np.array([SPLIT_RULE_CODES[r if (r is None or isinstance(r, str)) else type(r).__name__] for r in split_rules], dtype=np.int32)


class PosteriorSampler:
    """One chain's forest history (the object bartrs hands the shell: pymc_bart/utils.py:60-71,91,124-127).

    ``from_history(batches, baseline_forest, m, n_outputs)``, ``n_draws``, ``n_outputs`` and
    ``sample_posterior(X, draw_indices, excluded) -> (n_idx, n_outputs, n)``.  The chain is described on the host
    (``history.ChainHistory``: tree versions + a draw -> version table); a device store is built on first use, or once
    for all chains by ``_MultiChainSampler``."""

    def __init__(self, history: ChainHistory, split_rules=None, device: int = 0, subset_tables=None):
        self.history = history
        self.split_rules = split_rules
        self.device = device
        self.subset_tables = subset_tables
        self._store = None

    @classmethod
    def from_history(cls, batches, baseline_forest, m, n_outputs, split_rules=None, device: int = 0, subset_tables=None):
        return cls(ChainHistory(batches, baseline_forest, m, n_outputs), split_rules=split_rules, device=device,
                   subset_tables=subset_tables)

    @property
    def n_draws(self) -> int:
        return self.history.n_draws

    @property
    def n_outputs(self) -> int:
        return self.history.G * self.history.K

    def sample_posterior(self, X, draw_indices, excluded=None) -> np.ndarray:
        if self._store is None:
            self._store = _MultiChainSampler([self])
        return self._store.sample_posterior(X, draw_indices, excluded)


class _MultiChainSampler:
    """All chains of an op behind one sampler (pymc_bart/utils.py:74-107): draws are numbered chain after chain.

    The reference looks the chain of every draw up and loops over the chains; here the chains' histories are
    concatenated into one device store whose forest rows are in that same global order, so the lookup disappears."""

    def __init__(self, chain_samplers: list):
        if not chain_samplers:
            raise ValueError("No posterior draws available yet: run pm.sample() first.")
        first = chain_samplers[0]
        self._forests = DeviceForests([s.history for s in chain_samplers], split_rules=first.split_rules, device=first.device,
                                      subset_tables=getattr(first, "subset_tables", None))
        self.n_chains = len(chain_samplers)

    @property
    def n_draws(self) -> int:
        return self._forests.n_draws

    @property
    def n_outputs(self) -> int:
        return self._forests.G * self._forests.K

    def upload(self, X):
        """Device copy of X for repeated calls (the importance search predicts dozens of times on the same rows)."""
        return self._forests.upload(X)

    @staticmethod
    def _mask(excluded, p):
        if excluded is None or len(excluded) == 0:
            return None
        mk = np.zeros((1, p), dtype=np.uint8)
        mk[0, np.asarray(list(excluded), dtype=np.int64)] = 1
        return mk

    def sample_posterior(self, X, draw_indices, excluded=None) -> np.ndarray:
        Xd = self.upload(X)
        out = self._forests.predict(Xd, np.asarray(draw_indices), self._mask(excluded, int(Xd.shape[1])))
        return out[0].cpu().numpy().astype(np.float64)        # (n_idx, n_outputs, n)

    def predict_subsets(self, X, draws_per_subset, masks):
        """ONE launch for K exclusion masks, each with its own draw indices: device tensor [K][S][n_outputs][n]."""
        return self._forests.predict(self.upload(X), np.asarray(draws_per_subset), np.asarray(masks, dtype=np.uint8), per_mask=True)

    def pearson_r2(self, a, b):
        return self._forests.pearson_r2(a, b)


def _sample_posterior(sampler, X, rng, size=None, excluded=None):
    """Random posterior draws of the sum of trees at the rows of X (pymc_bart/utils.py:26-71): `size` draws are picked
    with ``rng.integers(0, n_draws)``; the result has shape ``(*size, n_rows, n_outputs)`` (``(n_rows, n_outputs)``
    for ``size=None``).  A list of samplers (several BART variables) is stacked along the output axis."""
    dims = () if size is None else ((int(size),) if np.isscalar(size) else tuple(int(v) for v in size))
    samplers = list(sampler) if isinstance(sampler, (list, tuple)) else [sampler]
    picks = rng.integers(0, samplers[0].n_draws, size=int(np.prod(dims)) if dims else 1)
    blocks = [s.sample_posterior(X, picks, None if excluded is None else list(excluded)) for s in samplers]
    pred = blocks[0] if len(blocks) == 1 else np.concatenate(blocks, axis=1)          # (n_picks, n_outputs, n_rows)
    return np.moveaxis(pred, 1, 2).reshape(*dims, pred.shape[2], pred.shape[1])


_posterior_sampler_cache: dict = {}


def _history_signature(op):
    """(chains, draws per chain): the history of a running sampler grows, the cached device store must follow."""
    return tuple(len(entry[1]) for entry in op.all_trees)


def _get_posterior_sampler(op) -> _MultiChainSampler:
    """The op's multi-chain sampler, rebuilt when chains or draws were added (pymc_bart/utils.py:110-130 caches on the
    chain count alone because bartrs publishes a chain's history in one piece)."""
    sig = _history_signature(op)
    cached = _posterior_sampler_cache.get(id(op))
    if cached is not None and cached[0] == sig and cached[2] is op:
        return cached[1]
    rules = _rule_codes(getattr(op, "split_rules", None))
    chains = [PosteriorSampler.from_history(list(batches), baseline_forest, op.m, op.n_outputs, split_rules=rules,
                                            device=getattr(op, "device", 0), subset_tables=getattr(op, "subset_tables", None))
              for baseline_forest, batches in op.all_trees]
    sampler = _MultiChainSampler(chains)
    _posterior_sampler_cache[id(op)] = (sig, sampler, op)
    return sampler


def get_variable_inclusion_counts(stats, n_cols: int) -> np.ndarray:
    """Sum of the decoded per-draw counts (what pymc_bart/utils.py:778-790 computes from idata)."""
    tot = np.zeros(n_cols, dtype=np.int64)
    for s in stats:
        tot += np.asarray(_decode_vi(s["variable_inclusion"], n_cols), dtype=np.int64)
    return tot
