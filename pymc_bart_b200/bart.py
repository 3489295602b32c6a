"""BART random variable: PyMC-free mirror of pymc_bart/bart.py (reference @4daa2e2).

The reference builds a per-variable ``BART_{name}`` RandomVariable op whose CLASS
ATTRIBUTES carry everything the step method needs (pymc_bart/bart.py:141-158):
``X, Y, m, alpha, beta, response, split_prior, split_rules, initval, all_trees``.
This module reproduces that object without PyMC/PyTensor so that the PGBART step
(pgbart.py) can be constructed from it exactly as ``bartrs.PGBART([rv], ...)`` is
(tests/test_bart.py:231).  When PyMC is importable, ``pymc_bart_b200.pymc_adapter``
wraps the same op into a real ``pm.Distribution``.
"""
from __future__ import annotations

import warnings

import numpy as np

__all__ = ["BART", "BARTRV", "preprocess_xy"]

_manager = None


def _history_manager():
    """One ``multiprocessing.Manager`` server for the process, started on first use.  The reference starts one per BART
    variable (pymc_bart/bart.py:133-135); the lists it serves are what crosses PyMC's per-chain worker processes."""
    global _manager
    if _manager is None:
        import multiprocessing

        # "spawn": the server must not be forked from a process that already runs CUDA / BLAS threads
        _manager = multiprocessing.get_context("spawn").Manager()
    return _manager


def sibling_list(all_trees):
    """A new list of the same kind as `all_trees`: a fresh shared list on the SAME manager server when `all_trees` is
    a Manager proxy (also from a worker process, which only holds the unpickled proxy), else a plain list.  The step
    uses it for a chain's `batches`, so that every draw ships only its own trees through the proxy."""
    from multiprocessing.managers import BaseProxy, SyncManager

    if not isinstance(all_trees, BaseProxy):
        return []
    mgr = getattr(all_trees, "_manager", None)
    if mgr is None:
        mgr = SyncManager(address=all_trees._token.address, authkey=all_trees._authkey)
        mgr.connect()
    return mgr.list()


def preprocess_xy(X, Y):
    """pandas / polars / array -> float64 numpy (pymc_bart/bart.py:193-212)."""
    for mod in ("pandas", "polars"):
        try:
            m = __import__(mod)
        except ImportError:
            continue
        if isinstance(X, (m.Series, m.DataFrame)):
            X = X.to_numpy()
        if isinstance(Y, (m.Series, m.DataFrame)):
            Y = Y.to_numpy()
    return np.asarray(X).astype(float), np.asarray(Y).astype(float)


class BARTRV:
    """Base class of the per-variable op (pymc_bart/bart.py:35-68)."""

    name = "BART"
    signature = "(m,n),(m),(),(),() -> (m)"
    dtype = "floatX"

    @classmethod
    def rng_fn(cls, rng=None, X=None, Y=None, m=None, alpha=None, beta=None, size=None):
        """Prior/posterior draw of the variable (pymc_bart/bart.py:47-68)."""
        from .utils import _get_posterior_sampler, _sample_posterior

        if not size:
            size = None
        if not getattr(cls, "all_trees", None):
            Yv = cls.Y
            if size is not None:
                return np.full((size[0], Yv.shape[0]), Yv.mean())
            return np.full(Yv.shape[0], Yv.mean())
        shape = size[0] if size is not None else 1
        sampler = _get_posterior_sampler(cls)
        pred = _sample_posterior(sampler, cls.X if X is None else X, rng=rng or np.random.default_rng(), size=shape)
        return pred.squeeze().T


class _Owner:
    def __init__(self, op):
        self.op = op


class BARTVariable:
    """What ``pmb.BART(...)`` hands back when no PyMC model is involved."""

    def __init__(self, name, op, shape):
        self.name = name
        self.owner = _Owner(op)
        self.shape = shape

    def __repr__(self):
        return f"BART({self.name}, shape={self.shape})"


def BART(name, X, Y, m=50, alpha=0.95, beta=2.0, response="constant", split_rules=None, split_prior=None,
         shape=None, separate_trees=False, shared_history=True, **kwargs):
    """Same signature as ``pmb.BART`` (pymc_bart/bart.py:115-127) plus ``separate_trees``
    (dropped from the reference at this commit; BASELINE.json config 4 asks for it).

    ``shared_history=True`` (the reference's behaviour, bart.py:133-135) makes ``op.all_trees`` a
    ``multiprocessing.Manager().list()`` proxy, so that step objects running in worker processes publish their tree
    history to the parent; ``False`` keeps a plain list (single process, no manager server)."""
    if response in ("linear", "mix"):
        warnings.warn("Options linear and mix are experimental and still not well tested\nUse with caution.")
    Xn, Yn = preprocess_xy(X, Y)
    sp = np.array([]) if split_prior is None else np.asarray(split_prior)
    mgr = _history_manager() if shared_history else None
    op_type = type(
        f"BART_{name}",
        (BARTRV,),
        {
            "name": "BART",
            "all_trees": mgr.list() if mgr is not None else [],   # bart.py:134-135
            "inplace": False,
            "initval": Yn.mean(),
            "X": Xn,
            "Y": Yn,
            "m": int(m),
            "response": response,
            "alpha": alpha,
            "beta": beta,
            "split_prior": sp,
            "split_rules": split_rules,
            "separate_trees": bool(separate_trees),
        },
    )
    op = op_type()
    n = Xn.shape[0]
    shp = (n,) if shape is None else tuple(np.atleast_1d(shape))
    return BARTVariable(name, op, shp)
