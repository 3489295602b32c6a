"""The C oracle against an independent float64 statement of the same algorithm (oracle/model_float.py) that uses none of
include/bk_spec.h: no fixed-point sums, numpy's exp / log / cos, textbook systematic resampling, its own Philox.

north_star's bar, applied to the oracle itself: identical tree topologies and leaf-index assignments for a fixed RNG stream,
leaf values and log-weights within 1e-5.  Every decision the sampler takes — popped node, split variable, split value,
child sizes, resampling ancestors, the selected particle, inclusion counts, every row's leaf id — must be THE SAME in both;
what may differ is rounding (float32 polynomial kernels and 2^-22 fixed point on one side, float64 libm on the other).
This is the check that a mistake in the shared header would not survive ("what's weak" 3)."""
import numpy as np
import pytest

from helpers import friedman
from oracle.model_float import FloatModelChain, philox4x32_10
from oracle.oracle_py import OracleChain
from pymc_bart_b200.settings import make_settings

INT_FIELDS = ("kind", "tree", "round", "particle", "node", "var", "ancestor")


LIK_NAMES = {0: "normal", 1: "bernoulli", 2: "normal_hetero", 3: "categorical"}


def compare(N, p, m, P, draws, seed, sigma=1.0, depth_offset=0, X=None, y=None, rules=None, split_prior=None, likelihood=0, n_outputs=1):
    if X is None:
        X, y, _ = friedman(N, p, seed, kind="bernoulli" if likelihood else "normal")
    s = make_settings(X, y, m=m, num_particles=P, seed=seed, depth_offset=depth_offset, trace_capacity=60000, split_rules=rules,
                      split_prior=split_prior, likelihood=likelihood, n_outputs=n_outputs)
    orc = OracleChain(s, X.T.copy(), y)
    mod = FloatModelChain(X, y, m, P, s.p_leaf, seed=s.seed, split_prior=s.split_prior, split_rules=s.split_rules,
                          likelihood=LIK_NAMES[likelihood], n_outputs=n_outputs)
    n_grow = 0
    for d in range(draws):
        tune = d < draws // 2
        vi_o, st_o = orc.step(tune, sigma)
        vi_m, grow_m = mod.step(tune, sigma)
        tr = orc.trace()
        assert len(tr) == len(mod.trace), f"draw {d}: {len(tr)} vs {len(mod.trace)} records"
        for a, b in zip(tr, mod.trace):
            ctx = f"draw {d}: oracle {a} model {b}"
            for k in INT_FIELDS:
                assert int(a[k]) == int(b[k]), ctx
            if b["kind"] == 1:
                assert int(a["n_left"]) == b["n_left"] and int(a["n_right"]) == b["n_right"], ctx
                assert float(a["split"]) == b["split"], ctx                                  # a value of X / a set of categories: exact
                assert abs(float(a["val_left"]) - b["val_left"]) < 1e-5 and abs(float(a["val_right"]) - b["val_right"]) < 1e-5, ctx
            else:
                assert abs(float(a["aux"]) - b["aux"]) < 1e-5 * max(1.0, abs(b["aux"])), ctx    # running leaf sd after the commit
            # (heteroscedastic Normal: ((y - f0) / |f1|)^2 in float32 is ill-conditioned for small |f1| — rows near the +-512
            # saturation carry 2e-7 * 500 each — so its log-weights get 1e-4; every decision must still be the same)
            # families whose weights are integer sums of per-row terms rounded to 2^-20: N half-steps of absolute slack
            tol = (1e-4 if likelihood == 2 else 1e-5) * max(1.0, abs(b["log_w"])) + (N * 2.0 ** -20 if likelihood else 0.0)
            assert abs(float(a["log_w"]) - b["log_w"]) < tol, ctx
        assert np.array_equal(vi_o, vi_m) and st_o.grow_events == grow_m
        n_grow += grow_m
    ids = orc.leaf_ids()
    nodes, nn = orc.forest()
    for t in range(m):
        assert np.array_equal(ids[t], mod.forest[t].ids)
        assert nn[t] == len(mod.forest[t].nodes)
        assert [int(v) for v in nodes[t]["var"][: nn[t]]] == [nd.var for nd in mod.forest[t].nodes]
    np.testing.assert_allclose(np.atleast_2d(orc.sum_trees()), mod.st, rtol=0, atol=5e-5)
    if n_outputs > 1:      # every output's leaf values, not only the first one the trace carries
        lv = orc.leaf_values()
        for t in range(m):
            for k, nd in enumerate(mod.forest[t].nodes):
                if nd.var < 0:
                    np.testing.assert_allclose(lv[t, k], nd.value, rtol=0, atol=1e-5)
    return n_grow


def test_philox_known_answers():
    """Random123 known-answer vectors: the model's Philox is its own implementation."""
    assert ["%08x" % v for v in philox4x32_10(0, 0, 0, 0, 0, 0)] == ["6627e8d5", "e169c58d", "bc57ac4c", "9b00dbd8"]
    assert ["%08x" % v for v in philox4x32_10(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)] == [
        "408f276d", "41c83b0e", "a20bc7c6", "6d5451fd"]
    assert ["%08x" % v for v in philox4x32_10(0xA4093822, 0x299F31D0, 0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344)] == [
        "d16cfe09", "94fdcceb", "5001e420", "24126ea1"]


def test_config1_same_decisions_120_draws():
    """BASELINE.json configs[0] (N=200 p=5 m=10 P=20): 120 draws, ~14 000 trace records."""
    assert compare(200, 5, 10, 20, 120, seed=1) > 2000


def test_historical_depth_prior_small_sigma():
    assert compare(150, 4, 6, 12, 40, seed=5, depth_offset=1, sigma=0.5) > 1000


def test_weighted_split_prior_and_tiny():
    assert compare(300, 6, 5, 10, 30, seed=26, split_prior=[5, 1, 0.5, 3, 0.1, 2]) > 300
    compare(3, 2, 3, 4, 20, seed=7)


def test_onehot_and_subset_rules():
    rng = np.random.default_rng(12345)
    Y = np.repeat(np.arange(3), 30).astype(np.float32)
    X = np.concatenate([Y[:, None], rng.integers(0, 6, size=(90, 4))], axis=1).astype(np.float32)
    assert compare(90, 5, 4, 10, 40, seed=13, X=X, y=Y, rules=["OneHotSplit"] * 5) > 100
    cat = rng.integers(0, 6, 400)
    Xs = np.stack([cat, rng.uniform(0, 1, 400), rng.integers(0, 4, 400)], axis=1).astype(np.float32)
    ys = (4.0 * np.isin(cat, [0, 3, 5]) + rng.normal(0, 0.3, 400)).astype(np.float32)
    assert compare(400, 3, 6, 10, 40, seed=61, X=Xs, y=ys, rules=["SubsetSplit", "ContinuousSplit", "SubsetSplit"], sigma=0.3,
                   depth_offset=1) > 300


def test_bernoulli_logit_likelihood():
    """The float32 polynomial softplus / exp kernels and the 2^-20 fixed-point terms of bk_spec.h against numpy's
    logaddexp in float64: same decisions, log-weights within 1e-5."""
    assert compare(300, 5, 6, 10, 40, seed=7, likelihood=1) > 200
    assert compare(120, 3, 4, 8, 30, seed=27, likelihood=1, depth_offset=1) > 100


def _with_missing(N, p, seed, kind="normal", all_nan_col=None):
    X, y, _ = friedman(N, p, seed, kind=kind)
    rng = np.random.default_rng(seed + 1000)
    X = X.copy()
    X[N // 6: N // 3, 0] = np.nan
    X[rng.uniform(size=N) < 0.1, 2] = np.nan
    if all_nan_col is not None:
        X[:, all_nan_col] = np.nan
    return X, y


def test_missing_covariates():
    """App. A.4 as the spec defines it: up to four candidate members per split-value draw, rows without the split covariate
    leave the tree (leaf id 255, predict 0), a column that is missing everywhere cancels the split."""
    X, y = _with_missing(300, 4, 31)
    assert compare(300, 4, 5, 10, 40, seed=31, X=X, y=y) > 200
    X, y = _with_missing(250, 5, 33, all_nan_col=1)
    assert compare(250, 5, 4, 10, 30, seed=33, X=X, y=y, depth_offset=1) > 200
    X, y = _with_missing(300, 5, 34, kind="bernoulli")
    assert compare(300, 5, 5, 8, 30, seed=34, X=X, y=y, likelihood=1) > 100
    rng = np.random.default_rng(62)
    cat = rng.integers(0, 6, 300).astype(np.float32)
    Xs = np.stack([cat, rng.uniform(0, 1, 300).astype(np.float32)], axis=1)
    ys = (4.0 * np.isin(cat, [0, 3, 5]) + rng.normal(0, 0.3, 300)).astype(np.float32)
    Xs[rng.uniform(size=300) < 0.15, 0] = np.nan
    assert compare(300, 2, 5, 10, 30, seed=62, X=Xs, y=ys, rules=["SubsetSplit", None], sigma=0.3, depth_offset=1) > 200


def test_shared_tree_multi_output_likelihoods():
    """K values per leaf, weights from the full (K, N) value: the reference's tested multi-output models
    (tests/test_bart.py:107-123 heteroscedastic Normal, :140-164 Categorical-softmax) — bk_lik_term's float32 kernels against
    numpy in float64."""
    rng = np.random.default_rng(5)
    X = rng.normal(0, 1, size=(250, 3)).astype(np.float32)
    y = (rng.normal(0, 1, size=250) + 2 * X[:, 0]).astype(np.float32)
    assert compare(250, 3, 4, 10, 30, seed=51, X=X, y=y, likelihood=2, n_outputs=2) > 100
    Yc = np.repeat(np.arange(3), 30).astype(np.float32)
    Xc = np.concatenate([Yc[:, None], rng.integers(0, 6, size=(90, 4))], axis=1).astype(np.float32)
    assert compare(90, 5, 5, 8, 40, seed=52, X=Xc, y=Yc, likelihood=3, n_outputs=3, depth_offset=1) > 100


def test_prediction_with_excluded_variables():
    """App. A.10 (the reference's `sample_posterior(X, draw_indices, excluded)`, pymc_bart/utils.py:60-71): the oracle's
    stack-based weighted descent (which the prediction kernel matches bit for bit, tests/test_gpu_api.py) against a
    recursive float64 walk, on new rows, with and without excluded variables, Continuous and Subset rules, NaNs in X."""
    from oracle import oracle_py
    from oracle.model_float import predict_tree_float

    rng = np.random.default_rng(9)
    cat = rng.integers(0, 6, 300)
    X = np.stack([cat, rng.uniform(0, 1, 300), rng.uniform(0, 1, 300)], axis=1).astype(np.float32)
    y = (4.0 * np.isin(cat, [0, 3, 5]) + 3 * X[:, 1] + rng.normal(0, 0.3, 300)).astype(np.float32)
    rules = ["SubsetSplit", "ContinuousSplit", "ContinuousSplit"]
    s = make_settings(X, y, m=8, num_particles=10, seed=9, split_rules=rules, depth_offset=1)
    orc = OracleChain(s, X.T.copy(), y)
    for d in range(40):
        orc.step(d < 20, 0.3)
    nodes, nn = orc.forest()
    Xn = np.stack([rng.integers(0, 8, 40), rng.uniform(-0.2, 1.2, 40), rng.uniform(0, 1, 40)], axis=1).astype(np.float32)
    Xn[3, 0] = np.nan; Xn[5, 1] = np.nan; Xn[7, 0] = 31.0          # missing values and a category code no set contains
    for excl in (None, [1], [0, 2]):
        mask = None
        if excl is not None:
            mask = np.zeros(3, np.uint8); mask[excl] = 1
        got = oracle_py.predict(nodes[None], Xn, [0], excluded_mask=mask, rules=s.split_rules)[0]
        want = [sum(predict_tree_float(nodes[t][: nn[t]], Xn[i], mask, s.split_rules) for t in range(8)) for i in range(40)]
        np.testing.assert_allclose(got, np.array(want), rtol=0, atol=2e-5)


def _same_step(orc, mod, tune, sigma, ctx):
    vi_o, st_o = orc.step(tune, sigma)
    vi_m, grow_m = mod.step(tune, sigma)
    tr = orc.trace()
    assert len(tr) == len(mod.trace), ctx
    for a, b in zip(tr, mod.trace):
        for k in INT_FIELDS:
            assert int(a[k]) == int(b[k]), f"{ctx}: oracle {a} model {b}"
        assert abs(float(a["log_w"]) - b["log_w"]) < 1e-5 * max(1.0, abs(b["log_w"])), f"{ctx}: oracle {a} model {b}"
    assert np.array_equal(vi_o, vi_m) and st_o.grow_events == grow_m, ctx


def test_separate_trees_groups_and_two_variables_in_one_likelihood():
    """An output group of BART(shape=(k, n), separate_trees=True) is a chain with its own Philox group word and response row;
    two BART variables in one Normal likelihood (tests/test_bart.py:167-241) alternate, each seeing the data minus the other's
    current value: same decisions in the oracle and in the float model."""
    X, y, _ = friedman(200, 4, 81)
    Y = np.stack([y, -y, 0.5 * y]).astype(np.float32)
    s = make_settings(X, Y, m=5, num_particles=8, seed=81, trace_capacity=20000, n_groups=3)
    for g in (0, 2):
        orc = OracleChain(s, X.T.copy(), Y, chain=0, group=g)
        mod = FloatModelChain(X, Y[g], 5, 8, s.p_leaf, seed=s.seed, group=g)
        mod.init_leaf[:] = np.float32(s.init_leaf); mod.st[:] = np.float32(s.init_sum)      # (one initial value for all groups: bart.py:148)
        mod.leaf_sd[:] = s.leaf_sd_init
        for nd in (f.nodes[0] for f in mod.forest):
            nd.value[:] = np.float32(s.init_leaf)
        for d in range(20):
            _same_step(orc, mod, d < 10, 1.0, f"group {g} draw {d}")
        np.testing.assert_allclose(orc.sum_trees(), mod.st[0], rtol=0, atol=5e-5)
    # two variables, one likelihood
    rng = np.random.default_rng(91)
    X1 = rng.normal(0, 1, size=(150, 2)).astype(np.float32); X2 = rng.normal(0, 1, size=(150, 3)).astype(np.float32)
    Y1 = (X1[:, 0] + rng.normal(0, 0.1, 150)).astype(np.float32); Y2 = (X2[:, 0] + X2[:, 1] + rng.normal(0, 0.1, 150)).astype(np.float32)
    Yobs = (Y1 + Y2).astype(np.float32)
    kw = dict(m=4, num_particles=8, trace_capacity=20000, value_range=float(np.abs(Yobs).max()))
    sa, sb = make_settings(X1, Y1, seed=91, **kw), make_settings(X2, Y2, seed=92, **kw)
    oa, ob = OracleChain(sa, X1.T.copy(), Y1.copy()), OracleChain(sb, X2.T.copy(), Y2.copy())
    ma, mb = FloatModelChain(X1, Y1, 4, 8, sa.p_leaf, seed=91), FloatModelChain(X2, Y2, 4, 8, sb.p_leaf, seed=92)
    va_o = np.full(150, np.float32(Y1.mean())); vb_o = np.full(150, np.float32(Y2.mean()))
    va_m, vb_m = va_o.copy(), vb_o.copy()
    for d in range(20):
        oa.y[0, :] = Yobs - vb_o; ma.y = (Yobs - vb_m).astype(np.float32)
        _same_step(oa, ma, d < 10, 0.5, f"variable A draw {d}")
        va_o, va_m = oa.sum_trees().copy(), ma.st[0].copy()
        ob.y[0, :] = Yobs - va_o; mb.y = (Yobs - va_m).astype(np.float32)
        _same_step(ob, mb, d < 10, 0.5, f"variable B draw {d}")
        vb_o, vb_m = ob.sum_trees().copy(), mb.st[0].copy()
    np.testing.assert_allclose(va_o + vb_o, va_m + vb_m, rtol=0, atol=1e-4)
