"""PGBART step method: host side of the B200 sampler (reference: ``bartrs.PGBART``,
constructed as ``PGBART([rv], num_particles=...)`` at tests/test_bart.py:231-232 and
driven by ``pm.sample`` one ``step(point)`` per draw).

State lives on the GPU (pymc_bart_b200.core.DeviceSampler); this class keeps the
step protocol: ``astep`` returns the new value of the BART variable (the sum of
trees) and the per-draw stats ``{"variable_inclusion": <base64 varint>, "tune": bool}``
(pymc_bart/utils.py:1387-1398, consumers :778-790), ``stop_tuning()`` ends
adaptation, and after tuning every step's rewritten trees are appended to the
op's history so that ``op.all_trees`` ends up with ONE ``(baseline_forest,
batches)`` entry per chain (pymc_bart/utils.py:117,124-127) and ``op.n_outputs`` is set
(utils.py:125).  Extension: ``chains=C`` batches C independent chains in one
launch (the reference runs one step object per chain/process).
"""
from __future__ import annotations

import numpy as np

from . import _cabi
from .core import DeviceSampler
from .settings import make_settings
from .utils import _encode_vi

LIKELIHOODS = {"normal": _cabi.BK_LIK_NORMAL, "bernoulli": _cabi.BK_LIK_BERNOULLI_LOGIT}


class PGBART:
    name = "pgbart"
    default_blocked = False
    generates_stats = True
    stats_dtypes_shapes = {"variable_inclusion": (object, []), "tune": (bool, [])}

    def __init__(self, vars=None, num_particles=10, batch=(0.1, 0.1), model=None, *, likelihood="normal", sigma=1.0,
                 chains=1, chain_base=0, seed=0, device=0, depth_offset=0, store_history=True, trace_capacity=0,
                 sigma_name=None, sigma_transform=None, **kwargs):
        if vars is None or len(vars) != 1:
            raise ValueError("PGBART takes exactly one BART variable: PGBART([rv], num_particles=...)")
        rv = vars[0]
        op = rv.owner.op if hasattr(rv, "owner") else rv
        if getattr(op, "name", None) != "BART" or not hasattr(op, "all_trees"):
            raise TypeError("PGBART can only sample BART variables")
        if op.response != "constant":
            raise NotImplementedError(f"response={op.response!r} has no device implementation (constant leaves only)")
        if np.isnan(np.asarray(op.X)).any():
            raise NotImplementedError("NaN covariates have no device implementation yet")
        if likelihood not in LIKELIHOODS:
            raise NotImplementedError(f"likelihood {likelihood!r} has no device implementation")
        # multi-output: BART(shape=(k, n), separate_trees=True) = k output groups, each with its own forest and
        # its own response row (independent likelihood per output); shared-tree multi-output has no device form
        shape = tuple(getattr(rv, "shape", (np.asarray(op.X).shape[0],)))
        groups = int(shape[0]) if len(shape) == 2 else 1
        if groups > 1 and not getattr(op, "separate_trees", False):
            raise NotImplementedError("multi-output BART needs separate_trees=True on the device (shared trees are not implemented)")
        Yarr = np.asarray(op.Y, dtype=np.float64)
        if groups > 1 and Yarr.ndim == 1:
            Yarr = np.broadcast_to(Yarr, (groups, Yarr.shape[0]))
        self.groups = groups
        self.op = op
        self.vars = [rv]
        self.num_particles = int(num_particles)
        self.batch = tuple(batch)
        self.chains = int(chains)
        self.tune = True
        self.sigma = sigma
        self.sigma_name = sigma_name
        self.sigma_transform = sigma_transform   # e.g. np.exp when sigma_name is the log-transformed value variable
        self.store_history = bool(store_history)
        self.settings = make_settings(
            op.X, Yarr, m=op.m, alpha=op.alpha, beta=op.beta, split_prior=op.split_prior, split_rules=op.split_rules,
            num_particles=num_particles, batch=batch, n_chains=chains, seed=seed, chain_base=chain_base,
            likelihood=LIKELIHOODS[likelihood], depth_offset=depth_offset, device=device, trace_capacity=trace_capacity,
            n_groups=groups,
        )
        self.core = DeviceSampler(self.settings, op.X, Yarr)
        self.core.enable_host_output(True)   # astep returns a host array every draw (the trace stores it)
        self.n_rows, self.n_cols, self.m = self.core.N, self.core.p, self.core.m
        self._lower = 0
        self._baseline = None   # per chain: (nodes [m,255], n_nodes [m])
        self._batches = [[] for _ in range(self.chains * self.groups)]
        self._published = False
        self.last_stats = None
        # read back through the CLASS by BARTRV.rng_fn -> _get_posterior_sampler(cls) (bart.py:65, utils.py:125)
        (op if isinstance(op, type) else type(op)).n_outputs = groups

    # ---- step-method protocol -------------------------------------------------
    @staticmethod
    def competence(var, has_grad=False):
        op = getattr(getattr(var, "owner", None), "op", None)
        return 3 if getattr(op, "name", None) == "BART" and hasattr(op, "all_trees") else 0  # 3 == Competence.IDEAL

    def stop_tuning(self):
        self.tune = False

    def astep(self, _q=None):
        tune = bool(self.tune)
        T = self.settings.batch_tune if tune else self.settings.batch_post
        lo = self._lower
        hi = min(lo + T, self.m)
        nvc = self.chains * self.groups          # "virtual chains": chain-major, group-minor
        if not tune and self.store_history and self._baseline is None:
            self._baseline = [self.core.forest(c) for c in range(nvc)]
        vi, stats = self.core.step(tune, self.sigma)
        self.last_stats = stats
        self._lower = hi if hi < self.m else 0
        value = self.core.sum_trees_host()
        if not tune and self.store_history:
            for c in range(nvc):
                nodes, nn = self.core.trees(c, lo, hi - lo)
                self._batches[c].append((lo, nodes, nn))
        vic = vi.reshape(self.chains, self.groups, -1).sum(axis=1)       # one inclusion vector per BART variable
        out_stats = [{"variable_inclusion": _encode_vi(vic[c].tolist()), "tune": tune} for c in range(self.chains)]
        value = value.reshape(self.chains, self.groups, -1)
        if self.groups == 1:
            value = value[:, 0]
        if self.chains == 1:
            return value[0].copy(), [out_stats[0]]
        return value.copy(), out_stats

    def step(self, point):
        """PyMC-style: reads the likelihood scale from the point when ``sigma_name`` is set."""
        if self.sigma_name is not None:
            if self.sigma_name not in point:   # a silently kept default would sample the wrong posterior
                raise KeyError(f"sigma_name={self.sigma_name!r} is not in the point (keys: {sorted(point)}); PyMC points are keyed "
                               "by value-variable names (e.g. 'sigma_log__'): pass sigma_transform= to map it back")
            v = point[self.sigma_name]
            self.sigma = self.sigma_transform(v) if self.sigma_transform is not None else v
        value, stats = self.astep(None)
        new_point = dict(point)
        new_point[self.vars[0].name] = value
        return new_point, stats

    # ---- history (pymc_bart/utils.py:117-127) -----------------------------------
    def publish_history(self):
        """Append one (baseline_forest, batches) entry per chain to op.all_trees."""
        if self._published or self._baseline is None:
            return
        for c in range(self.chains):
            if self.groups == 1:
                self.op.all_trees.append((self._baseline[c], list(self._batches[c])))
            else:   # one (baseline, batches) pair per output group inside the chain's entry
                g0 = c * self.groups
                self.op.all_trees.append(([self._baseline[g0 + g] for g in range(self.groups)],
                                          [list(self._batches[g0 + g]) for g in range(self.groups)]))
        self._published = True

    def close(self):
        self.core.close()
