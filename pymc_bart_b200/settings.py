"""Host-side derivation of the sampler settings from the BART op attributes.

Mirrors what the reference's step constructor reads from the per-variable op
(pymc_bart/bart.py:141-158): X, Y, m, alpha, beta, split_prior, split_rules.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import _cabi

SPLIT_RULE_CODES = {
    None: _cabi.BK_RULE_CONTINUOUS,
    "ContinuousSplit": _cabi.BK_RULE_CONTINUOUS,   # tests/test_bart.py:143
    "ContinuousSplitRule": _cabi.BK_RULE_CONTINUOUS,
    "OneHotSplit": _cabi.BK_RULE_ONEHOT,           # tests/test_bart.py:144
    "OneHotSplitRule": _cabi.BK_RULE_ONEHOT,
    "SubsetSplit": _cabi.BK_RULE_SUBSET,           # docs/api_reference.rst:16, pymc_bart/bart.py:103
    "SubsetSplitRule": _cabi.BK_RULE_SUBSET,
}


def subset_category_tables(X: np.ndarray, rules: np.ndarray):
    """Category tables of the SubsetSplit columns: {column: sorted unique non-missing values}.  The device works on
    integer category CODES (the position of a value in its column's table, at most 24 of them: the set of categories
    that go left travels as a 24-bit mask); the reference's rule takes arbitrary values (`np.isin(x, subset)`)."""
    tables = {}
    cols = np.nonzero(np.asarray(rules) == _cabi.BK_RULE_SUBSET)[0]
    if cols.size > _cabi.BK_MAX_SUBSET_COLS:
        raise NotImplementedError(f"at most {_cabi.BK_MAX_SUBSET_COLS} columns may use SubsetSplit")
    for v in cols:
        col = np.asarray(X[:, v], dtype=np.float64)
        cats = np.unique(col[~np.isnan(col)])
        if cats.size > _cabi.BK_SUBSET_MAX_CATS:
            raise NotImplementedError(f"SubsetSplit column {int(v)} holds {cats.size} categories; the device handles up to "
                                      f"{_cabi.BK_SUBSET_MAX_CATS}")
        tables[int(v)] = cats
    return tables


def encode_subset_columns(X: np.ndarray, tables: dict) -> np.ndarray:
    """X with every SubsetSplit column replaced by its category codes (float; NaN stays NaN; a value that is not in the
    column's table — possible only in new data — gets the code 31, which belongs to no set and therefore goes right,
    like `np.isin` of an unseen value)."""
    if not tables:
        return X
    X = np.array(X, dtype=np.float64, copy=True)
    for v, cats in tables.items():
        col = X[:, v]
        nan = np.isnan(col)
        pos = np.searchsorted(cats, col)
        pos_c = np.clip(pos, 0, max(0, cats.size - 1))
        known = (~nan) & (cats.size > 0) & (cats[pos_c] == col)
        X[:, v] = np.where(nan, np.nan, np.where(known, pos_c.astype(np.float64), 31.0))
    return X


def depth_prior_table(alpha: float, beta: float, depth_offset: int = 0) -> np.ndarray:
    """P(node at depth d stays a leaf), d = 0..255.

    depth_offset=0: documented prior, a node at depth d is non-terminal with
    probability alpha*(1+d)^-beta (pymc_bart/bart.py:107-109).
    depth_offset=1: the historical pure-Python indexing:
    the root always attempts a split and depth d>=1 uses alpha*d^-beta.
    """
    d = np.arange(256, dtype=np.float64)
    if depth_offset == 0:
        t = 1.0 - alpha * np.power(1.0 + d, -beta)
    elif depth_offset == 1:
        t = np.empty(256)
        t[0] = 0.0
        t[1:] = 1.0 - alpha * np.power(d[1:], -beta)
    else:
        raise ValueError("depth_offset must be 0 or 1")
    return np.ascontiguousarray(np.clip(t, 0.0, 1.0))


def choose_qshift(abs_max: float) -> int:
    """Fixed-point scale 2^qshift with 4x head-room over max|Y| inside 29 bits."""
    b = max(float(abs_max), 1e-30)
    k = int(math.floor(math.log2((2.0**29 - 1.0) / (4.0 * b))))
    return max(-60, min(60, k))


@dataclass
class SamplerSettings:
    n_rows: int
    n_cols: int
    n_trees: int
    n_particles: int
    n_chains: int
    likelihood: int
    qshift: int
    batch_tune: int
    batch_post: int
    seed: int
    chain_base: int
    init_sum: float
    init_leaf: float
    leaf_sd_init: float
    device: int
    trace_capacity: int
    n_groups: int
    n_outputs: int
    p_leaf: np.ndarray
    split_prior: np.ndarray
    split_rules: np.ndarray
    _keep: list = field(default_factory=list, repr=False)

    def to_c(self) -> _cabi.BkSettings:
        s = _cabi.BkSettings()
        s.abi_version = _cabi.BK_ABI_VERSION
        for name in ("n_rows", "n_cols", "n_trees", "n_particles", "n_chains", "likelihood", "qshift", "batch_tune",
                     "batch_post", "seed", "chain_base", "device", "trace_capacity", "n_groups", "n_outputs"):
            setattr(s, name, int(getattr(self, name)))
        s.init_sum = float(self.init_sum)
        s.init_leaf = float(self.init_leaf)
        s.leaf_sd_init = float(self.leaf_sd_init)
        self.p_leaf = np.ascontiguousarray(self.p_leaf, dtype=np.float64)
        self.split_prior = np.ascontiguousarray(self.split_prior, dtype=np.float64)
        self.split_rules = np.ascontiguousarray(self.split_rules, dtype=np.int32)
        s.p_leaf = self.p_leaf.ctypes.data_as(C.POINTER(C.c_double))
        s.split_prior = self.split_prior.ctypes.data_as(C.POINTER(C.c_double))
        s.split_rules = self.split_rules.ctypes.data_as(C.POINTER(C.c_int32))
        return s


def make_settings(
    X: np.ndarray,
    Y: np.ndarray,
    m: int = 50,
    alpha: float = 0.95,
    beta: float = 2.0,
    split_prior=None,
    split_rules=None,
    num_particles: int = 10,
    batch=(0.1, 0.1),
    n_chains: int = 1,
    seed: int = 0,
    chain_base: int = 0,
    likelihood: int = _cabi.BK_LIK_NORMAL,
    depth_offset: int = 0,
    device: int = 0,
    trace_capacity: int = 0,
    n_groups: int = 1,
    n_outputs: int = 1,
    value_range: float = 0.0,
) -> SamplerSettings:
    X = np.asarray(X)
    Y = np.asarray(Y, dtype=np.float64)
    n, p = X.shape
    if not (0.0 < alpha < 1.0):
        raise ValueError("alpha must be in (0, 1)")
    if beta <= 0:
        raise ValueError("beta must be positive")
    if split_prior is None or len(np.atleast_1d(split_prior)) == 0:  # bart.py:139
        sp = np.ones(p, dtype=np.float64)
    else:
        sp = np.asarray(split_prior, dtype=np.float64)
        if sp.shape != (p,) or np.any(sp < 0) or sp.sum() <= 0:
            raise ValueError("split_prior must hold one non-negative weight per column")
    if split_rules is None:
        rules = np.zeros(p, dtype=np.int32)
    else:
        if len(split_rules) != p:
            raise ValueError("split_rules must hold one rule per column")
        try:
            rules = np.array([SPLIT_RULE_CODES[r if (r is None or isinstance(r, str)) else type(r).__name__] for r in split_rules], dtype=np.int32)
        except KeyError as e:
            raise NotImplementedError(f"split rule {e} has no device implementation") from None
    ymean = float(Y.mean())
    uniq = np.unique(Y)
    if uniq.size == 2 and set(uniq.tolist()) == {0.0, 1.0}:
        leaf_sd = 3.0 / math.sqrt(m)
    else:
        leaf_sd = float(Y.std()) / math.sqrt(m)
    if int(n_outputs) > 1:
        if int(likelihood) not in (_cabi.BK_LIK_NORMAL_HETERO, _cabi.BK_LIK_CATEGORICAL):
            raise NotImplementedError("shared-tree multi-output BART needs likelihood 'normal_hetero' or 'categorical' on the device")
        if int(n_outputs) > _cabi.BK_MAX_OUTPUTS or int(n_groups) > 1:
            raise NotImplementedError(f"at most {_cabi.BK_MAX_OUTPUTS} shared-tree outputs, and not together with separate trees")
        if int(likelihood) == _cabi.BK_LIK_NORMAL_HETERO and int(n_outputs) != 2:
            raise ValueError("the heteroscedastic Normal likelihood takes shape=(2, n): mean and scale")
        if int(likelihood) == _cabi.BK_LIK_CATEGORICAL and not np.all((Y == np.floor(Y)) & (Y >= 0) & (Y < int(n_outputs))):
            raise ValueError("the Categorical likelihood needs integer labels 0..k-1")
        if np.isnan(np.asarray(X, dtype=np.float64)).any():
            raise NotImplementedError("missing covariates are not supported together with shared-tree multi-output")
        if np.any(rules == _cabi.BK_RULE_SUBSET):
            raise NotImplementedError("SubsetSplit is not supported together with shared-tree multi-output")
        qshift = choose_qshift(max(16.0, float(np.abs(Y).max()), abs(ymean)))   # linear predictors: fixed-point range of at least +-64
    elif int(likelihood) in (_cabi.BK_LIK_NORMAL_HETERO, _cabi.BK_LIK_CATEGORICAL):
        raise ValueError("likelihoods 'normal_hetero' / 'categorical' need a multi-output BART variable (shape=(k, n))")
    elif int(likelihood) == _cabi.BK_LIK_BERNOULLI_LOGIT:
        if not np.all((Y == 0.0) | (Y == 1.0)):
            raise ValueError("the Bernoulli likelihood needs a 0/1 response")
        qshift = choose_qshift(16.0)    # the sum of trees is a logit: fixed-point range +-64
    else:
        # (value_range: largest magnitude of the likelihood's data when it is not the Y handed to BART — observed= of the step)
        qshift = choose_qshift(max(float(np.abs(Y).max()), abs(ymean), float(value_range)))
    bt = max(1, int(m * batch[0]))
    bp = max(1, int(m * batch[1]))
    init_sum = np.float32(ymean)
    init_leaf = np.float32(ymean / m)
    return SamplerSettings(
        n_rows=n, n_cols=p, n_trees=int(m), n_particles=int(num_particles), n_chains=int(n_chains),
        likelihood=int(likelihood), qshift=qshift,
        batch_tune=bt, batch_post=bp, seed=int(seed) & 0xFFFFFFFF, chain_base=int(chain_base),
        init_sum=float(init_sum), init_leaf=float(init_leaf), leaf_sd_init=float(np.float32(leaf_sd)),
        device=int(device), trace_capacity=int(trace_capacity), n_groups=max(1, int(n_groups)), n_outputs=max(1, int(n_outputs)),
        p_leaf=depth_prior_table(alpha, beta, depth_offset), split_prior=sp, split_rules=rules,
    )
