"""PyMC binding of the B200 sampler (what `import bartrs` does for the reference,
pymc_bart/__init__.py:15-18): a PyMC step class around ``pymc_bart_b200.pgbart.PGBART`` registered in
``pm.STEP_METHODS`` so that ``pm.sample()`` assigns it to BART variables.

UNTESTED IN THIS REPOSITORY: PyMC cannot be installed offline, so this module is exercised only
up to its import guard.  It follows the step-method protocol of SURVEY.md App. C:

* ``PGBART(vars, num_particles=..., batch=..., model=...)`` (tests/test_bart.py:231-235);
* ``astep`` ignores the raveled point (state lives on the GPU) and returns ``(value.ravel(), [stats])`` with the
  stats of pymc_bart/utils.py:1387-1398;
* ``stop_tuning()`` ends adaptation; when the chain is done PyMC never calls back, so the history is published
  (one ``(baseline_forest, batches)`` entry per chain, pymc_bart/utils.py:117-127) as soon as tuning stops and
  refreshed by ``publish_history()``.

The likelihood must be one of the closed families of the device path: pass ``likelihood="normal"`` with
``sigma_name=<name of the scale variable in the point>`` or ``likelihood="bernoulli"``; an arbitrary PyTensor
``datalogp`` has no device form (no CPU fallback).
"""
from __future__ import annotations

try:  # pragma: no cover - PyMC is absent in this environment
    import pymc as pm
    from pymc.step_methods.arraystep import ArrayStepShared
    from pymc.step_methods.compound import Competence
except ImportError as exc:  # the product never falls back: say what is missing
    raise ImportError(
        "pymc_bart_b200.pymc_adapter needs PyMC (pymc>=5); the PyMC-free driver is pymc_bart_b200.sample()"
    ) from exc

from .pgbart import PGBART as _CorePGBART


class PGBART(ArrayStepShared):  # pragma: no cover
    name = "pgbart"
    default_blocked = False
    generates_stats = True
    stats_dtypes_shapes = {"variable_inclusion": (object, []), "tune": (bool, [])}

    def __init__(self, vars=None, num_particles=10, batch=(0.1, 0.1), model=None, likelihood="normal", sigma_name=None,
                 **kwargs):
        model = pm.modelcontext(model)
        if vars is None:
            vars = [v for v in model.free_RVs if getattr(v.owner.op, "name", None) == "BART"]
        value_vars = [model.rvs_to_values[v] for v in vars]
        core_kw = {k: kwargs.pop(k) for k in ("seed", "device", "depth_offset", "chain_base") if k in kwargs}
        self._core = _CorePGBART(vars, num_particles=num_particles, batch=batch, likelihood=likelihood,
                                 sigma_name=sigma_name, **core_kw)
        self._sigma_name = sigma_name
        self.tune = True
        super().__init__(value_vars, [], **kwargs)

    def step(self, point):
        if self._sigma_name is not None and self._sigma_name in point:
            self._core.sigma = point[self._sigma_name]
        return super().step(point)

    def astep(self, _):
        self._core.tune = self.tune
        value, stats = self._core.astep()
        return value.ravel(), stats

    def stop_tuning(self):
        self.tune = False
        self._core.stop_tuning()

    def publish_history(self):
        self._core.publish_history()

    @staticmethod
    def competence(var, has_grad):
        op = getattr(getattr(var, "owner", None), "op", None)
        return Competence.IDEAL if getattr(op, "name", None) == "BART" and hasattr(op, "all_trees") else Competence.INCOMPATIBLE


pm.STEP_METHODS = list(pm.STEP_METHODS) + [PGBART]  # pragma: no cover
