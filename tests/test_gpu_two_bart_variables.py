"""Several BART variables in one model (tests/test_bart.py:167-241: `pm.Normal("y", mu1 + mu2, sigma, observed=Y)` with one
PGBART step per variable).  Each step weighs its particles with Normal(observed - other variable | value, sigma) at the
current point; the device reads the response afresh in every step (core.set_response), the oracle reads its `y` array."""
import numpy as np
import pytest

from helpers import assert_trace_equal
from pymc_bart_b200.settings import make_settings

pytestmark = pytest.mark.gpu


def _data(N, seed):
    rng = np.random.default_rng(seed)
    X1 = rng.normal(0, 1, size=(N, 2)).astype(np.float32)
    X2 = rng.normal(0, 1, size=(N, 3)).astype(np.float32)
    Y1 = (X1[:, 0] + rng.normal(0, 0.1, N)).astype(np.float32)
    Y2 = (X2[:, 0] + X2[:, 1] + rng.normal(0, 0.1, N)).astype(np.float32)
    return X1, X2, Y1, Y2


def test_two_samplers_alternate_bit_identical_to_two_oracle_chains():
    from oracle.oracle_py import OracleChain
    from pymc_bart_b200.core import DeviceSampler

    N = 300
    X1, X2, Y1, Y2 = _data(N, 91)
    Yobs = (Y1 + Y2).astype(np.float32)
    rng_kw = dict(m=5, num_particles=8, trace_capacity=20000, value_range=float(np.abs(Yobs).max()))
    sa = make_settings(X1, Y1, seed=91, **rng_kw)
    sb = make_settings(X2, Y2, seed=92, **rng_kw)
    da, db = DeviceSampler(sa, X1, Y1), DeviceSampler(sb, X2, Y2)
    oa, ob = OracleChain(sa, X1.T.copy(), Y1), OracleChain(sb, X2.T.copy(), Y2)
    va_d = np.full(N, np.float32(Y1.mean())); vb_d = np.full(N, np.float32(Y2.mean()))
    va_o, vb_o = va_d.copy(), vb_d.copy()
    for d in range(40):
        tune = d < 20
        # variable A given B
        da.set_response(Yobs - vb_d); oa.y[0, :] = Yobs - vb_o
        _, st = da.step(tune, 0.5); oa.step(tune, 0.5)
        assert st[0].error_flags == 0
        assert_trace_equal(da.trace(0), oa.trace(), f"A draw {d}")
        va_d = da.sum_trees().cpu().numpy()[0].copy(); va_o = oa.sum_trees().copy()
        assert np.array_equal(va_d.view(np.uint32), va_o.view(np.uint32)), f"A draw {d}"
        # variable B given the new A
        db.set_response(Yobs - va_d); ob.y[0, :] = Yobs - va_o
        _, st = db.step(tune, 0.5); ob.step(tune, 0.5)
        assert st[0].error_flags == 0
        assert_trace_equal(db.trace(0), ob.trace(), f"B draw {d}")
        vb_d = db.sum_trees().cpu().numpy()[0].copy(); vb_o = ob.sum_trees().copy()
        assert np.array_equal(vb_d.view(np.uint32), vb_o.view(np.uint32)), f"B draw {d}"
    assert np.corrcoef(va_d + vb_d, Yobs)[0, 1] > 0.9           # the two variables share the signal between them
    da.close(); db.close()


def test_multiple_bart_variables_manual_steps():
    """The reference's test (tests/test_bart.py:208-241) through the step protocol: two BART variables, one manually
    built PGBART each, driven point by point as pm.sample's compound step does."""
    import pymc_bart_b200 as pmb
    from pymc_bart_b200.utils import _decode_vi, _get_posterior_sampler

    N = 50
    X1, X2, Y1, Y2 = _data(N, 93)
    Yobs = Y1.astype(np.float64) + Y2
    mu1 = pmb.BART("mu1", X1, Y1, m=5)
    mu2 = pmb.BART("mu2", X2, Y2, m=5)
    step1 = pmb.PGBART([mu1], num_particles=5, sigma=0.3, observed=Yobs, offset_names=["mu2"], seed=1)
    step2 = pmb.PGBART([mu2], num_particles=5, sigma=0.3, observed=Yobs, offset_names=["mu1"], seed=2)
    point = {"mu1": np.full(N, Y1.mean()), "mu2": np.full(N, Y2.mean())}
    post1, post2, vi1, vi2 = [], [], [], []
    for d in range(100):
        if d == 50:
            step1.stop_tuning(); step2.stop_tuning()
        point, s1 = step1.step(point)
        point, s2 = step2.step(point)
        if d >= 50:
            post1.append(point["mu1"].copy()); post2.append(point["mu2"].copy())
            vi1.append(s1[0]["variable_inclusion"]); vi2.append(s2[0]["variable_inclusion"])
    step1.flush_history(); step2.flush_history()
    post1, post2 = np.stack(post1), np.stack(post2)
    assert post1.shape == (50, N) and post2.shape == (50, N)                       # idata.posterior["mu1"].shape == (1, 50, 50)
    op1, op2 = mu1.owner.op, mu2.owner.op
    assert op1.all_trees is not op2.all_trees and len(op1.all_trees) == 1 and len(op2.all_trees) == 1
    fit = (post1 + post2).mean(axis=0)
    assert np.corrcoef(fit, Yobs)[0, 1] > 0.9
    # one inclusion vector per variable, of that variable's width (utils.py:779-787 stacks them along variable_inclusion_dim_0)
    c1 = np.sum([_decode_vi(s, 2) for s in vi1], axis=0); c2 = np.sum([_decode_vi(s, 3) for s in vi2], axis=0)
    assert c1.shape == (2,) and c2.shape == (3,) and c1.sum() > 0 and c2.sum() > 0
    # each variable predicts from its own history (compute_variable_importance(idata, mu1, X1) in the reference test)
    p1 = _get_posterior_sampler(op1).sample_posterior(X1, [0, 49], None)
    np.testing.assert_allclose(p1[:, 0, :], post1[[0, 49]], atol=3e-4, rtol=0)
    p2 = _get_posterior_sampler(op2).sample_posterior(X2, [0, 49], None)
    np.testing.assert_allclose(p2[:, 0, :], post2[[0, 49]], atol=3e-4, rtol=0)
    step1.close(); step2.close()
