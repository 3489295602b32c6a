"""The PyMC binding (pymc_bart_b200/pymc_adapter.py) driven through a minimal stand-in for PyMC.

PyMC is not installable in the build environment, so the four names the adapter touches — ``pm.modelcontext``,
``pm.STEP_METHODS``, ``pymc.step_methods.arraystep.ArrayStepShared``, ``pymc.step_methods.compound.Competence`` — are
provided by a fake package that behaves as SURVEY.md App. C describes: ``step(point)`` ravels nothing, calls
``astep`` and stores the returned value under the value variable's name; the sampler loop sets ``tune`` at the start of
every chain and calls ``stop_tuning()`` after the tuning draws.  With a real PyMC importable the test is skipped
(the stand-in must not shadow it)."""
import enum
import sys
import types

import numpy as np
import pytest

from test_cabi_and_host import _FakeCore


class _Var:
    def __init__(self, name, op=None):
        self.name = name
        self.owner = types.SimpleNamespace(op=op) if op is not None else None


class _LogTransform:
    @staticmethod
    def backward(v):
        return np.exp(v)


@pytest.fixture
def fake_pymc(monkeypatch):
    try:
        import pymc  # noqa: F401
        pytest.skip("a real PyMC is installed")
    except ImportError:
        pass
    pm = types.ModuleType("pymc")
    pm.STEP_METHODS = ["NUTS", "Metropolis"]
    pm.modelcontext = lambda model: model
    arraystep = types.ModuleType("pymc.step_methods.arraystep")
    compound = types.ModuleType("pymc.step_methods.compound")

    class Competence(enum.IntEnum):
        INCOMPATIBLE = 0
        COMPATIBLE = 1
        PREFERRED = 2
        IDEAL = 3

    class ArrayStepShared:
        def __init__(self, vars, shared, blocked=True, rng=None):
            self.vars = vars
            self.shared = shared

        def step(self, point):
            value, stats = self.astep(None)
            new = dict(point)
            new[self.vars[0].name] = value
            return new, stats

    arraystep.ArrayStepShared = ArrayStepShared
    compound.Competence = Competence
    sm = types.ModuleType("pymc.step_methods")
    for name, mod in (("pymc", pm), ("pymc.step_methods", sm), ("pymc.step_methods.arraystep", arraystep),
                      ("pymc.step_methods.compound", compound)):
        monkeypatch.setitem(sys.modules, name, mod)
    monkeypatch.delitem(sys.modules, "pymc_bart_b200.pymc_adapter", raising=False)
    return pm


def test_registration_and_two_chains_through_one_step_object(fake_pymc, monkeypatch):
    import pymc_bart_b200 as pmb
    import pymc_bart_b200.pgbart as pg

    monkeypatch.setattr(pg, "DeviceSampler", _FakeCore)
    import pymc_bart_b200.pymc_adapter as ad                      # what `import bartrs` does (pymc_bart/__init__.py:15-18)

    assert fake_pymc.STEP_METHODS[-1] is ad.PGBART and fake_pymc.STEP_METHODS[:2] == ["NUTS", "Metropolis"]
    rng = np.random.default_rng(1)
    X = rng.normal(size=(40, 3)); Y = rng.normal(size=40)
    mu = pmb.BART("mu", X, Y, m=10)
    sigma_rv = _Var("sigma")
    model = types.SimpleNamespace(
        free_RVs=[mu, sigma_rv], rvs_to_values={mu: _Var("mu"), sigma_rv: _Var("sigma_log__")},
        rvs_to_transforms={sigma_rv: _LogTransform(), mu: None})
    assert ad.PGBART.competence(mu, False) == 3 and ad.PGBART.competence(sigma_rv, False) == 0
    with pytest.raises(ValueError, match="name the likelihood"):
        ad.PGBART(model=model)                                     # no silent Normal(sigma=1)
    with pytest.raises(ValueError, match="needs sigma"):
        ad.PGBART(model=model, likelihood="normal")
    step = ad.PGBART(model=model, likelihood="normal", sigma=sigma_rv, num_particles=5)     # vars found among model.free_RVs
    assert step.vars[0].name == "mu" and step._core.sigma_name == "sigma_log__"
    op = mu.owner.op
    point = {"mu": np.zeros(40), "sigma_log__": np.log(0.7)}
    for chain in range(2):                                          # pm.sample(chains=2, cores=1): one step object, two chains
        step.tune = True
        for d in range(6):
            if d == 3:
                step.stop_tuning()
            point, stats = step.step(point)
            assert point["mu"].shape == (40,) and stats[0]["tune"] == (d < 3)
            assert step._core.core.last_sigma == pytest.approx(0.7)   # backward-transformed value of the point, not log(sigma)
        step._core.flush_history()
        assert len(op.all_trees) == chain + 1 and len(op.all_trees[chain][1]) == 3    # published while sampling, one entry per chain
    assert [c.chain_base for c in _FakeCore.instances[-2:]] == [0, 1]                   # the second chain has its own Philox stream
    with pytest.raises(KeyError):
        step.step({"mu": np.zeros(40)})                             # the scale must be in the point
    fixed = ad.PGBART([mu], model=model, likelihood="normal", sigma=2.0)
    fixed.step({"mu": np.zeros(40)})
    assert fixed._core.core.last_sigma == 2.0


def test_two_bart_variables_through_the_adapter(fake_pymc, monkeypatch):
    """tests/test_bart.py:211-241 (`step=[PGBART([mu1], ...), PGBART([mu2], ...)]`, `pm.Normal("y", mu1 + mu2, sigma,
    observed=Y)`): each step is told the data and the other term of the location as random variables; before it runs it
    reads the other variable's current value from the point."""
    import pymc_bart_b200 as pmb
    import pymc_bart_b200.pgbart as pg

    monkeypatch.setattr(pg, "DeviceSampler", _FakeCore)
    import pymc_bart_b200.pymc_adapter as ad

    rng = np.random.default_rng(2)
    X1 = rng.normal(size=(30, 2)); X2 = rng.normal(size=(30, 2)); Y = rng.normal(size=30)
    mu1 = pmb.BART("mu1", X1, X1[:, 0], m=3); mu2 = pmb.BART("mu2", X2, X2[:, 1], m=3)
    model = types.SimpleNamespace(free_RVs=[mu1, mu2], rvs_to_values={mu1: _Var("mu1"), mu2: _Var("mu2")}, rvs_to_transforms={})
    s1 = ad.PGBART([mu1], model=model, likelihood="normal", sigma=1.0, num_particles=5, observed=Y, offset=[mu2])
    s2 = ad.PGBART([mu2], model=model, likelihood="normal", sigma=1.0, num_particles=5, observed=Y, offset=mu1)
    assert s1._core.offset_names == ["mu2"] and s2._core.offset_names == ["mu1"]
    point = {"mu1": np.zeros(30), "mu2": np.ones(30)}
    point, _ = s1.step(point)
    np.testing.assert_allclose(s1._core.core.response[0], Y - 1.0)
    point, _ = s2.step(point)
    np.testing.assert_allclose(s2._core.core.response[0], Y - point["mu1"])
    assert mu1.owner.op.all_trees is not mu2.owner.op.all_trees
    with pytest.raises(KeyError):
        s1.step({"mu1": np.zeros(30)})
