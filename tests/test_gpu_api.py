"""GPU tests of the reference-facing API: PGBART step protocol, history -> PosteriorSampler, prediction
kernel vs the oracle's restatement, and the reference's statistical tests re-expressed without PyMC."""
import numpy as np
import pytest

from helpers import friedman
from pymc_bart_b200 import _cabi
from pymc_bart_b200.utils import _decode_vi

pytestmark = pytest.mark.gpu


def _sample(X, Y, m, P, tune, draws, chains=1, seed=0, **kw):
    import pymc_bart_b200 as pmb

    rv = pmb.BART("mu", X, Y, m=m, **{k: v for k, v in kw.items() if k in ("split_rules", "split_prior")})
    out = pmb.sample(rv, tune=tune, draws=draws, chains=chains, num_particles=P, seed=seed, sigma=kw.get("sigma", 1.0))
    return rv, out


def test_step_protocol_and_vi_dominance():
    """tests/test_bart.py:44-64: X[:,0] ~ Y => variable 0 dominates; stats are base64 varints; value shape (N,)."""
    import pymc_bart_b200 as pmb

    rng = np.random.default_rng(3415)
    X = rng.normal(0, 1, size=(250, 3)); Y = rng.normal(0, 1, size=250)
    X[:, 0] = rng.normal(Y, 0.1)
    mu = pmb.BART("mu", X, Y, m=10)
    step = pmb.PGBART([mu], num_particles=10, seed=3415)
    assert step.tune is True and pmb.PGBART.competence(mu) == 3
    tot = np.zeros(3, dtype=np.int64)
    for d in range(400):
        if d == 200:
            step.stop_tuning()
        value, stats = step.astep()
        assert value.shape == (250,) and value.dtype == np.float32
        assert set(stats[0]) == {"variable_inclusion", "tune"} and stats[0]["tune"] == (d < 200)
        vi = np.asarray(_decode_vi(stats[0]["variable_inclusion"], 3))
        if d < 200:
            assert vi.sum() == 0          # inclusion is counted after tuning only (App. A.8)
        tot += vi
    frac = tot / tot.sum()
    assert frac[0] > frac[1:].sum()
    step.publish_history()
    op = mu.owner.op
    assert len(op.all_trees) == 1 and op.n_outputs == 1          # one (baseline, batches) entry per chain (utils.py:117)
    baseline, batches = op.all_trees[0]
    assert len(batches) == 200 and baseline[0].shape == (10, _cabi.BK_MAX_NODES)
    step.close()


def test_posterior_sampler_matches_in_sample_draws_and_oracle_predict():
    from oracle import oracle_py
    from pymc_bart_b200.utils import PosteriorSampler, _get_posterior_sampler, _sample_posterior

    X, Y, _ = friedman(400, 6, 17)
    rv, out = _sample(X, Y, m=12, P=8, tune=30, draws=25, chains=2, seed=17)
    op = rv.owner.op
    assert len(op.all_trees) == 2
    sampler = _get_posterior_sampler(op)
    assert sampler.n_draws == 50 and sampler.n_outputs == 1
    post = out["posterior"]                       # (chains, draws, N) values the sampler itself produced
    for chain in range(2):
        baseline, batches = op.all_trees[chain]
        forests = PosteriorSampler.rebuild_forests(batches, baseline, op.m)
        ps = PosteriorSampler(forests)
        idx = [0, 7, 24]
        pred = ps.sample_posterior(X, idx, None)                       # (3, 1, N)
        assert pred.shape == (3, 1, 400)
        np.testing.assert_allclose(pred[:, 0, :], post[chain, idx, :], atol=3e-4, rtol=0)
        ref = oracle_py.predict(forests, X.astype(np.float32), idx)
        assert np.array_equal(pred[:, 0, :].astype(np.float32), ref)   # kernel vs oracle restatement, bit for bit
        ex = [1, 3]
        mask = np.zeros(6, np.uint8); mask[ex] = 1
        pe = ps.sample_posterior(X[:50], idx, ex)
        re = oracle_py.predict(forests, X[:50].astype(np.float32), idx, excluded_mask=mask)
        assert np.array_equal(pe[:, 0, :].astype(np.float32), re)
    # tests/test_utils.py:24-32 — prediction self-consistency and shapes
    pa = _sample_posterior(sampler, X=X, rng=np.random.default_rng(3), size=2)
    pf = _sample_posterior(sampler, X=X[:10], rng=np.random.default_rng(3))
    np.testing.assert_almost_equal(pf, pa[0, :10], decimal=4)
    assert pa.shape == (2, 400, 1) and pf.shape == (10, 1)
    # rng_fn on new data after sampling (tests/test_bart.py:84-104 shapes)
    assert type(op).rng_fn(rng=np.random.default_rng(0), X=X[:3]).shape == (3,)
    out["step"].close()


def test_fit_quality_friedman():
    X, Y, f = friedman(3000, 10, 5)
    rv, out = _sample(X, Y, m=40, P=16, tune=150, draws=50, chains=2, seed=5)
    mean = out["posterior"].mean(axis=(0, 1))
    rmse = float(np.sqrt(np.mean((mean - f) ** 2)))
    rmse0 = float(np.sqrt(np.mean((Y.mean() - f) ** 2)))     # the initial constant fit
    assert rmse < 0.3 * rmse0, (rmse, rmse0)   # 200 draws x 4 trees/draw: every tree rewritten ~20 times
    out["step"].close()


def test_unsupported_options_raise_not_fallback():
    import pymc_bart_b200 as pmb

    X = np.random.default_rng(0).normal(size=(20, 2)); Y = np.zeros(20)
    with pytest.warns(UserWarning):
        mu = pmb.BART("mu", X, Y, m=3, response="linear")
    with pytest.raises(NotImplementedError):
        pmb.PGBART([mu])
    Xn = X.copy(); Xn[3, 0] = np.nan
    with pytest.raises(NotImplementedError):
        pmb.PGBART([pmb.BART("b", Xn, Y, m=3)])
    with pytest.raises(NotImplementedError):
        pmb.PGBART([pmb.BART("c", X, Y, m=3)], likelihood="poisson")


def test_multi_output_api_shapes_and_prediction():
    """pmb.BART(..., shape=(3, n), separate_trees=True): value (3, n), one VI vector per variable, prediction (.., 3, n)."""
    import pymc_bart_b200 as pmb
    from pymc_bart_b200.utils import _get_posterior_sampler, _sample_posterior

    X, y, _ = friedman(300, 5, 8)
    Y = np.stack([y, -y, 0.5 * y])
    mu = pmb.BART("w", X, Y, m=6, shape=(3, 300), separate_trees=True)
    out = pmb.sample(mu, tune=60, draws=10, chains=1, num_particles=8, seed=8)
    assert out["posterior"].shape == (1, 10, 3, 300)
    op = mu.owner.op
    assert type(op).n_outputs == 3 and len(op.all_trees) == 1
    sampler = _get_posterior_sampler(op)
    pred = _sample_posterior(sampler, X[:7], rng=np.random.default_rng(0), size=4)
    assert pred.shape == (4, 7, 3)
    # outputs 0 and 1 were fitted to y and -y: their posterior means must be strongly anti-correlated
    m0, m1 = out["posterior"][0, :, 0].mean(0), out["posterior"][0, :, 1].mean(0)
    assert np.corrcoef(m0, m1)[0, 1] < -0.25 and np.corrcoef(m0, y)[0, 1] > 0.4 and np.corrcoef(m1, -y)[0, 1] > 0.4
    with pytest.raises(NotImplementedError):
        pmb.PGBART([pmb.BART("s", X, y, m=3, shape=(2, 300))])       # shared-tree multi-output
    out["step"].close()


def test_bernoulli_api_recovers_the_logit():
    """BASELINE.json configs[2] through the public API: 0/1 response, PGBART(likelihood="bernoulli"); the posterior
    mean of the BART variable is the logit (tests/test_bart.py:150-164 checks recovery the same way)."""
    import pymc_bart_b200 as pmb

    X, y, f = friedman(4000, 8, 3, kind="bernoulli")
    mu = pmb.BART("mu", X, y, m=30)
    out = pmb.sample(mu, tune=80, draws=40, chains=2, num_particles=12, seed=3, likelihood="bernoulli")
    logit = out["posterior"].mean(axis=(0, 1))
    assert logit.shape == (4000,) and np.all(np.isfinite(logit))
    pr = 1 / (1 + np.exp(-(f - 14.4) / 4.9))
    ll = float(np.mean(y * logit - np.logaddexp(0, logit)))
    ll_true = float(np.mean(y * np.log(pr) + (1 - y) * np.log(1 - pr)))
    ll0 = float(np.mean(y * 0.5 - np.logaddexp(0, 0.5)))             # the initial constant fit (Y.mean() as a logit)
    assert ll > ll0 + 0.08 and abs(ll - ll_true) < 0.03
    assert np.corrcoef(logit, f)[0, 1] > 0.8
    with pytest.raises(ValueError):
        pmb.PGBART([pmb.BART("bad", X, y + 0.25, m=5)], likelihood="bernoulli")
    out["step"].close()
