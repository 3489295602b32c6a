"""PyMC binding of the B200 sampler — what ``import bartrs`` does for the reference (pymc_bart/__init__.py:15-18): a
PyMC step class around ``pymc_bart_b200.pgbart.PGBART``, appended to ``pm.STEP_METHODS`` so that ``pm.sample()`` assigns
it to BART variables (tests/test_bart.py:58,187) or takes it as ``step=[PGBART([mu], num_particles=5)]`` (:231-235).

PyMC cannot be installed in this repository's build environment (no network), so the module is exercised by
tests/test_pymc_adapter.py against a minimal stand-in for the four PyMC names it touches (``modelcontext``,
``STEP_METHODS``, ``ArrayStepShared``, ``Competence``); against a real PyMC it is untested.  Step-method protocol
followed:

* ``PGBART(vars, num_particles=..., batch=..., model=..., likelihood=..., sigma=...)``;
* ``astep`` ignores the raveled point (the state lives on the GPU) and returns ``(value.ravel(), [stats])`` with the
  stats of pymc_bart/utils.py:1387-1398;
* ``stop_tuning()`` ends adaptation; from the first posterior draw on the chain's ``(baseline_forest, batches)`` entry is
  in ``op.all_trees`` and grows by one batch per draw (pymc_bart/utils.py:117-127) — nothing has to be called when
  sampling ends;
* PyMC sets ``tune`` back to True when it starts the next chain on the same step object (``cores=1``): the core then
  starts a fresh chain with the next chain index.

The likelihood must be one of the closed families of the device path and must be named: ``likelihood="normal"`` with
``sigma=<the scale random variable, a value-variable name, or a number>``, or ``likelihood="bernoulli"``.  An arbitrary
PyTensor ``datalogp`` has no device form, and there is no CPU fallback.
"""
from __future__ import annotations

try:
    import pymc as pm
    from pymc.step_methods.arraystep import ArrayStepShared
    from pymc.step_methods.compound import Competence
except ImportError as exc:  # the product never falls back: say what is missing
    raise ImportError(
        "pymc_bart_b200.pymc_adapter needs PyMC (pymc>=5); the PyMC-free driver is pymc_bart_b200.sample()"
    ) from exc

import numpy as np

from .pgbart import PGBART as _CorePGBART


def _resolve_sigma(model, sigma):
    """(fixed value, point key, backward transform) for the likelihood scale.

    PyMC points are keyed by VALUE-variable names and hold transformed values (a HalfNormal ``sigma`` lives in the
    point as ``sigma_log__``), so a scale given as a random variable is mapped to its value variable and the backward
    transform of ``model.rvs_to_transforms``; a string is taken as a point key holding the untransformed scale."""
    if sigma is None:
        return None, None, None
    if isinstance(sigma, (int, float, np.floating)):
        return float(sigma), None, None
    if isinstance(sigma, str):
        return None, sigma, None
    value_var = model.rvs_to_values[sigma]
    transform = getattr(model, "rvs_to_transforms", {}).get(sigma)
    back = None
    if transform is not None:
        def back(v, _t=transform):
            out = _t.backward(v)
            return float(out.eval()) if hasattr(out, "eval") else float(out)
    return None, value_var.name, back


class PGBART(ArrayStepShared):
    name = "pgbart"
    default_blocked = False
    generates_stats = True
    stats_dtypes_shapes = {"variable_inclusion": (object, []), "tune": (bool, [])}

    def __init__(self, vars=None, num_particles=10, batch=(0.1, 0.1), model=None, likelihood=None, sigma=None, observed=None,
                 offset=None, **kwargs):
        model = pm.modelcontext(model)
        if vars is None:
            vars = [v for v in model.free_RVs if getattr(getattr(getattr(v, "owner", None), "op", None), "name", None) == "BART"]
        if likelihood is None:
            raise ValueError("name the likelihood of the observed variable: likelihood='normal' (with sigma=...) or 'bernoulli'; "
                             "the device path has no form for an arbitrary datalogp and does not guess")
        if likelihood == "normal" and sigma is None:
            raise ValueError("likelihood='normal' needs sigma= (the scale random variable, a point key, or a number)")
        value_vars = [model.rvs_to_values[v] for v in vars]
        fixed, key, back = _resolve_sigma(model, sigma)
        core_kw = {k: kwargs.pop(k) for k in ("seed", "device", "depth_offset", "chain_base", "store_history", "lookahead", "tune_draws")
                   if k in kwargs}     # (lookahead / tune_draws: draws served ahead, fixed likelihood parameters only — see pgbart.py)
        # several BART variables in one likelihood (tests/test_bart.py:167-241): observed= the data, offset= the OTHER random
        # variables of the location (their current values are read from the point before every step)
        offset_names = None
        if offset is not None:
            offset_names = [o if isinstance(o, str) else model.rvs_to_values[o].name for o in (offset if isinstance(offset, (list, tuple)) else [offset])]
        self._core = _CorePGBART(vars, num_particles=num_particles, batch=batch, likelihood=likelihood,
                                 sigma=1.0 if fixed is None else fixed, sigma_name=key, sigma_transform=back,
                                 observed=observed, offset_names=offset_names, **core_kw)
        self.tune = True
        super().__init__(value_vars, [], **kwargs)

    def step(self, point):
        core = self._core
        if core.sigma_name is not None:
            if core.sigma_name not in point:
                raise KeyError(f"the likelihood scale {core.sigma_name!r} is not in the point (keys: {sorted(point)})")
            v = point[core.sigma_name]
            core.sigma = core.sigma_transform(v) if core.sigma_transform is not None else float(v)
        if core.offset_names:
            missing = [n for n in core.offset_names if n not in point]
            if missing:
                raise KeyError(f"offset variables {missing} are not in the point (keys: {sorted(point)})")
            core.set_offset(sum(np.asarray(point[n], dtype=np.float64) for n in core.offset_names))
        return super().step(point)

    def astep(self, _):
        self._core.tune = self.tune     # (a flip back to True after posterior draws starts the next chain in the core)
        value, stats = self._core.astep()
        return np.asarray(value).ravel(), stats

    def stop_tuning(self):
        self.tune = False
        self._core.stop_tuning()

    @staticmethod
    def competence(var, has_grad):
        op = getattr(getattr(var, "owner", None), "op", None)
        return Competence.IDEAL if getattr(op, "name", None) == "BART" and hasattr(op, "all_trees") else Competence.INCOMPATIBLE


if PGBART not in pm.STEP_METHODS:
    pm.STEP_METHODS = list(pm.STEP_METHODS) + [PGBART]
