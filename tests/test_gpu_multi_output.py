"""Shared-tree multi-output BART (the reference's own multi-output mode at this commit: ``BART(shape=(k, n))`` without
``separate_trees``; tests/test_bart.py:107-123 heteroscedastic Normal, :140-164 Categorical-softmax) on the GPU vs the CPU
oracle, bit for bit: every leaf carries k values, the particle weight is the likelihood of the whole (k, n) value."""
import numpy as np
import pytest

from helpers import assert_trace_equal, friedman
from pymc_bart_b200 import _cabi
from pymc_bart_b200.settings import make_settings

pytestmark = pytest.mark.gpu


def run_multi(X, y, K, lik, m, P, draws, seed, chains=1, depth_offset=0, split_rules=None):
    from oracle.oracle_py import OracleChain
    from pymc_bart_b200.core import DeviceSampler

    s = make_settings(X, y, m=m, num_particles=P, seed=seed, n_chains=chains, depth_offset=depth_offset, trace_capacity=40000,
                      likelihood=lik, n_outputs=K, split_rules=split_rules)
    dev = DeviceSampler(s, X, y)
    oracles = [OracleChain(s, np.ascontiguousarray(np.asarray(X, dtype=np.float32).T), y, chain=c) for c in range(chains)]
    N = X.shape[0]
    for d in range(draws):
        tune = d < draws // 2
        vi, stats = dev.step(tune, 1.0)
        st_dev = dev.sum_trees().cpu().numpy().reshape(chains, K, N)
        for c, o in enumerate(oracles):
            vio, sto = o.step(tune, 1.0)
            ctx = f"draw {d} chain {c}"
            assert stats[c].error_flags == 0, ctx
            assert_trace_equal(dev.trace(c), o.trace(), ctx)
            assert np.array_equal(st_dev[c].view(np.uint32), o.sum_trees().view(np.uint32)), ctx
            assert np.array_equal(vi[c], vio), ctx
            assert stats[c].grow_events == sto.grow_events and stats[c].rounds == sto.rounds, ctx
            assert np.float32(stats[c].leaf_sd).view(np.uint32) == np.float32(sto.leaf_sd).view(np.uint32), ctx
    for c, o in enumerate(oracles):
        nd, nn = dev.forest(c)
        no, nno = o.forest()
        assert np.array_equal(nn, nno) and np.array_equal(nd.view(np.uint8), no.view(np.uint8))
        assert np.array_equal(dev.leaf_ids(c), o.leaf_ids())
        assert np.array_equal(dev.leaf_values(c).view(np.uint32), o.leaf_values().view(np.uint32))
    dev.close()
    return True


def _categorical(N, p, K, seed):
    X, _, f = friedman(N, p, seed)
    y = np.clip(np.floor((f - f.min()) / (np.ptp(f) + 1e-9) * K), 0, K - 1).astype(np.float32)
    return X, y


def test_categorical_reference_shape():
    """tests/test_bart.py:140-164: N=9, p=5, three classes, m=2."""
    rng = np.random.default_rng(0)
    Y = np.repeat([0, 1, 2], 3).astype(np.float32)
    X = np.concatenate([Y[:, None], rng.integers(0, 6, size=(9, 4))], axis=1).astype(np.float32)
    assert run_multi(X, Y, 3, _cabi.BK_LIK_CATEGORICAL, 2, 10, 80, seed=3)
    assert run_multi(X, Y, 3, _cabi.BK_LIK_CATEGORICAL, 2, 10, 40, seed=4, split_rules=["OneHotSplit"] * 5)


def test_categorical_many_tiles_and_outputs():
    X, y = _categorical(3000, 6, 4, 51)
    assert run_multi(X, y, 4, _cabi.BK_LIK_CATEGORICAL, 6, 12, 16, seed=51, chains=2)
    X, y = _categorical(700, 5, 7, 52)                       # the maximum number of outputs
    assert run_multi(X, y, 7, _cabi.BK_LIK_CATEGORICAL, 4, 8, 20, seed=52, depth_offset=1)


def test_heteroscedastic_normal():
    """tests/test_bart.py:107-123: w = BART(shape=(2, n)), y ~ Normal(w[0], |w[1]|)."""
    rng = np.random.default_rng(1)
    X = rng.uniform(-1, 1, size=(900, 3)).astype(np.float32)
    y = (2 * X[:, 0] + (0.3 + np.abs(X[:, 1])) * rng.normal(size=900)).astype(np.float32)
    assert run_multi(X, y, 2, _cabi.BK_LIK_NORMAL_HETERO, 8, 10, 30, seed=53)
    assert run_multi(X[:250, :2], rng.normal(size=250).astype(np.float32), 2, _cabi.BK_LIK_NORMAL_HETERO, 2, 10, 40, seed=54, chains=2)


def test_shared_tree_api_shapes_and_recovery():
    """The reference's two multi-output tests through the public API: value / posterior shapes (tests/test_bart.py:119-123)
    and recovery of the classes by the posterior mean class probabilities (:158-164); prediction returns k outputs per
    tree walk."""
    import pymc_bart_b200 as pmb
    from pymc_bart_b200.utils import _get_posterior_sampler, _sample_posterior

    rng = np.random.default_rng(0)
    Y = np.repeat([0, 1, 2], 3)
    X = np.concatenate([Y[:, None], rng.integers(0, 6, size=(9, 4))], axis=1).astype(float)
    mu = pmb.BART("mu", X, Y, m=2, shape=(3, 9))
    out = pmb.sample(mu, tune=300, draws=300, chains=1, num_particles=10, seed=3, likelihood="categorical")
    post = out["posterior"]                                            # (chains, draws, 3, 9)
    assert post.shape == (1, 300, 3, 9)
    e = np.exp(post - post.max(axis=2, keepdims=True))
    prob = (e / e.sum(axis=2, keepdims=True)).mean(axis=(0, 1))
    assert np.array_equal(prob.argmax(axis=0), Y)
    op = mu.owner.op
    assert type(op).n_outputs == 3 and len(op.all_trees) == 1
    sampler = _get_posterior_sampler(op)
    assert sampler.n_outputs == 3 and sampler.n_draws == 300
    pred = sampler.sample_posterior(X, [0, 299], None)                  # in-sample prediction reproduces the draws
    np.testing.assert_allclose(pred, post[0, [0, 299]], atol=3e-4, rtol=0)
    assert _sample_posterior(sampler, X[:4], rng=np.random.default_rng(0), size=5).shape == (5, 4, 3)
    out["step"].close()
    Xh = rng.normal(0, 1, size=(250, 2)); Yh = rng.normal(0, 1, size=250)
    w = pmb.BART("w", Xh, Yh, m=2, shape=(2, 250))
    outh = pmb.sample(w, tune=50, draws=50, chains=2, num_particles=8, seed=5, likelihood="normal_hetero")
    assert outh["posterior"].shape == (2, 50, 2, 250) and np.all(np.isfinite(outh["posterior"]))
    outh["step"].close()
    with pytest.raises(NotImplementedError):
        pmb.PGBART([pmb.BART("s", Xh, Yh, m=3, shape=(2, 250))])          # shared trees need a multi-output likelihood
    with pytest.raises(ValueError):
        pmb.PGBART([pmb.BART("t", Xh, Yh, m=3)], likelihood="categorical")
