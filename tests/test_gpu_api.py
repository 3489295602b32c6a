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
    step.flush_history()
    op = mu.owner.op
    assert len(op.all_trees) == 1 and op.n_outputs == 1          # one (baseline, batches) entry per chain (utils.py:117)
    baseline, batches = op.all_trees[0]
    assert len(batches) == 200 and baseline[1].shape == (10,) and baseline[0].shape == (int(baseline[1].sum()),)
    first, nn, nodes = batches[0]
    assert first == 0 and nn.shape == (1, 1) and nodes.shape == (int(nn.sum()),) and nodes.dtype == _cabi.NODE_DTYPE
    step.close()


def test_posterior_sampler_matches_in_sample_draws_and_oracle_predict():
    from oracle import oracle_py
    from pymc_bart_b200.history import ChainHistory
    from pymc_bart_b200.utils import PosteriorSampler, _get_posterior_sampler, _sample_posterior

    X, Y, _ = friedman(400, 6, 17)
    rv, out = _sample(X, Y, m=12, P=8, tune=30, draws=25, chains=2, seed=17)
    op = rv.owner.op
    assert len(op.all_trees) == 2
    sampler = _get_posterior_sampler(op)
    assert sampler.n_draws == 50 and sampler.n_outputs == 1
    post = out["posterior"]                       # (chains, draws, N) values the sampler itself produced
    idx = [0, 7, 24]
    ex = [1, 3]
    mask = np.zeros(6, np.uint8); mask[ex] = 1
    for chain in range(2):
        baseline, batches = op.all_trees[chain]
        ps = PosteriorSampler.from_history(list(batches), baseline, op.m, 1)      # one chain on its own (utils.py:124-127)
        assert ps.n_draws == 25 and ps.n_outputs == 1
        pred = ps.sample_posterior(X, idx, None)                       # (3, 1, N)
        assert pred.shape == (3, 1, 400)
        np.testing.assert_allclose(pred[:, 0, :], post[chain, idx, :], atol=3e-4, rtol=0)
        dense = ChainHistory(list(batches), baseline, op.m, 1).dense_forests()     # the CPU restatement predicts from dense forests
        ref = oracle_py.predict(dense, X.astype(np.float32), idx)
        assert np.array_equal(pred[:, 0, :].astype(np.float32), ref)   # kernel vs oracle restatement, bit for bit
        pe = ps.sample_posterior(X[:50], idx, ex)
        re = oracle_py.predict(dense, X[:50].astype(np.float32), idx, excluded_mask=mask)
        assert np.array_equal(pe[:, 0, :].astype(np.float32), re)
        # the multi-chain store numbers the draws chain after chain: same numbers through the global index
        glob = sampler.sample_posterior(X[:50], [25 * chain + i for i in idx], ex)
        assert np.array_equal(glob, pe)
    # one launch for several exclusion masks, each with its own draws (row N4)
    masks = np.zeros((3, 6), np.uint8); masks[1, ex] = 1; masks[2, [0]] = 1
    draws = np.array([[0, 30], [7, 49], [24, 25]])
    multi = sampler.predict_subsets(X[:64], draws, masks).cpu().numpy()
    for k in range(3):
        one = sampler.sample_posterior(X[:64], draws[k], np.nonzero(masks[k])[0].tolist() or None)
        assert np.array_equal(multi[k].astype(np.float64), one)
    # tests/test_utils.py:24-32 — prediction self-consistency and shapes
    pa = _sample_posterior(sampler, X=X, rng=np.random.default_rng(3), size=2)
    pf = _sample_posterior(sampler, X=X[:10], rng=np.random.default_rng(3))
    np.testing.assert_almost_equal(pf, pa[0, :10], decimal=4)
    assert pa.shape == (2, 400, 1) and pf.shape == (10, 1)
    # rng_fn on new data after sampling (tests/test_bart.py:84-104 shapes)
    assert type(op).rng_fn(rng=np.random.default_rng(0), X=X[:3]).shape == (3,)
    out["step"].close()


def test_variable_importance_search_on_device():
    """Row N4 (pymc_bart/utils.py:868-1090): every level of the search is one prediction launch over all candidate
    subsets + one fused correlation launch; the numbers equal the reference's sequential formulation (one
    sample_posterior call per subset, pearsonr2 per sample) for the same random_seed."""
    import pymc_bart_b200 as pmb
    from pymc_bart_b200.importance import compute_variable_importance, generate_sequences, get_variable_inclusion
    from pymc_bart_b200.utils import _get_posterior_sampler, _sample_posterior, get_variable_inclusion_counts

    rng = np.random.default_rng(11)
    X = rng.uniform(0, 1, (300, 5))
    Y = 6 * X[:, 2] + 3 * X[:, 0] + rng.normal(0, 0.3, 300)
    mu = pmb.BART("mu", X, Y, m=20)
    out = pmb.sample(mu, tune=150, draws=60, chains=2, num_particles=10, seed=11, sigma=0.3)
    stats = [{"variable_inclusion": s} for ch in out["variable_inclusion"] for s in ch]
    op = mu.owner.op
    sampler = _get_posterior_sampler(op)

    def pearsonr2(A, B):            # pymc_bart/utils.py:1339-1346
        am, bm = A.ravel() - A.mean(), B.ravel() - B.mean()
        return (am @ bm) ** 2 / (np.sum(am ** 2) * np.sum(bm ** 2))

    # --- "VI": subsets by increasing inclusion, evaluated sequentially like the reference
    res = compute_variable_importance(stats, mu, X, method="VI", samples=12, random_seed=5)
    r = np.random.default_rng(5)
    p_all = _sample_posterior(sampler, X, r, size=12)
    idxs = np.argsort(get_variable_inclusion_counts(stats, 5))
    subsets = [list(idxs[:-i]) for i in range(1, 5)] + [None]
    for k, sub in enumerate(subsets):
        p_sub = _sample_posterior(sampler, X, r, size=12, excluded=sub)
        r2 = np.array([pearsonr2(p_all[j], p_sub[j]) for j in range(12)])
        assert res["r2_mean"][k] == pytest.approx(r2.mean(), rel=1e-5)
        np.testing.assert_allclose(res["preds"][k], p_sub.squeeze(), rtol=0, atol=0)
    assert res["indices"].tolist() == idxs[::-1].tolist() and res["preds_all"].shape == (12, 300)
    assert set(res["indices"][:2].tolist()) == {0, 2}                        # the two informative columns rank first
    assert res["r2_mean"][-1] > 0.9 and res["r2_mean"][0] < res["r2_mean"][-1]   # the full set reproduces itself
    assert res["labels"][0] == str(res["indices"][0]) and res["labels"][1].startswith("+ ")
    # --- "backward": one launch per level, first-maximum choice
    resb = compute_variable_importance(None, mu, X, method="backward", samples=8, random_seed=6)
    r = np.random.default_rng(6)
    p_all = _sample_posterior(sampler, X, r, size=8)
    least = []
    for i_var in range(5):
        best, best_mean = None, -np.inf
        for sub in generate_sequences(5, i_var, least):
            p_sub = _sample_posterior(sampler, X, r, size=8, excluded=sub)
            mean = np.mean([pearsonr2(p_all[j], p_sub[j]) for j in range(8)])
            if mean > best_mean:
                best, best_mean = sub, mean
        assert resb["r2_mean"][::-1][i_var] == pytest.approx(best_mean, rel=1e-5)
        least += [v for v in best if v not in least]
    least += [v for v in range(5) if v not in least]                           # (utils.py:1062-1065: the variables never excluded come last)
    assert resb["indices"].tolist() == least[::-1] and set(resb["indices"][:2].tolist()) == {0, 2}
    resbv = compute_variable_importance(stats, mu, X, method="backward_VI", fixed=2, samples=6, random_seed=7)
    assert sorted(resbv["indices"].tolist()) == [0, 1, 2, 3, 4] and resbv["r2_mean"].shape == (5,)
    vi_norm, labels = get_variable_inclusion(stats, X)
    assert vi_norm.sum() == pytest.approx(1.0) and labels[0] in ("0", "2")
    with pytest.raises(ValueError):
        compute_variable_importance(stats, mu, X, method="forward")
    out["step"].close()


def _spawned_chain(payload, chain, q):
    """Worker process of test_history_crosses_processes: unpickles the step (no device state inside), runs one chain."""
    import pickle

    step = pickle.loads(payload)
    step.chain_base = chain
    vals = []
    for d in range(14):
        if d == 8:
            step.stop_tuning()
        v, _ = step.astep()
        if d >= 8:
            vals.append(v.copy())
    dev = step.core.device.index
    step.close()                                   # (drains the history writer thread before the process reports back)
    q.put((chain, np.stack(vals), dev))


def test_history_crosses_processes():
    """pymc_bart/bart.py:133-146: op.all_trees is a Manager proxy because PyMC runs every chain in its own process with
    a pickled copy of the step.  Two spawned workers sample one chain each; the parent predicts from what they
    published and gets the values the workers saw."""
    import multiprocessing as mp

    import cloudpickle

    import pymc_bart_b200 as pmb
    from pymc_bart_b200.utils import _get_posterior_sampler

    X, Y, _ = friedman(300, 4, 21)
    mu = pmb.BART("mu", X, Y, m=8)
    step = pmb.PGBART([mu], num_particles=6, seed=21)
    payload = cloudpickle.dumps(step)              # what PyMC ships to a worker; the parent never created a device sampler
    assert step.core is None
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_spawned_chain, args=(payload, c, q)) for c in range(2)]
    [p.start() for p in procs]
    got = dict()
    for _ in range(2):
        chain, vals, dev = q.get(timeout=300)
        got[chain] = vals
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    op = mu.owner.op
    assert len(op.all_trees) == 2 and sorted(len(b) for _, b in op.all_trees) == [6, 6]
    sampler = _get_posterior_sampler(op)
    assert sampler.n_draws == 12
    pred = sampler.sample_posterior(X, list(range(12)), None)[:, 0, :]      # entries are in arrival order: match by value
    for chain in range(2):
        hit = [k for k in range(2) if np.allclose(pred[6 * k: 6 * k + 6], got[chain], atol=3e-4, rtol=0)]
        assert len(hit) == 1, chain


def test_fit_quality_friedman():
    X, Y, f = friedman(3000, 10, 5)
    rv, out = _sample(X, Y, m=40, P=16, tune=150, draws=50, chains=2, seed=5)
    mean = out["posterior"].mean(axis=(0, 1))
    rmse = float(np.sqrt(np.mean((mean - f) ** 2)))
    rmse0 = float(np.sqrt(np.mean((Y.mean() - f) ** 2)))     # the initial constant fit
    assert rmse < 0.3 * rmse0, (rmse, rmse0)   # 200 draws x 4 trees/draw: every tree rewritten ~20 times
    out["step"].close()


def test_lookahead_serves_the_same_draws():
    """PGBART(lookahead=n): after tuning, astep is served from launches of n steps (the next launch runs while the
    caller consumes the previous one).  Same draws, same stats, same published history as one step per call."""
    import pymc_bart_b200 as pmb

    X, Y, _ = friedman(700, 5, 71)
    runs = []
    for la, td in ((1, None), (6, None), (6, 10)):   # (6, 10): the 10 tuning calls are served ahead as well
        mu = pmb.BART(f"mu{la}{td}", X, Y, m=20, shared_history=False)
        step = pmb.PGBART([mu], num_particles=8, chains=2, seed=71, sigma=0.8, lookahead=la, tune_draws=td)
        vals, vis = [], []
        for d in range(10 + 17):                      # 17 posterior draws: launches of 6, 6, 6 (one runs ahead)
            if d == 10:
                step.stop_tuning()
            v, st = step.astep()
            if d >= 10:
                vals.append(v.copy()); vis.append([s["variable_inclusion"] for s in st])
        step.flush_history()
        runs.append((np.stack(vals), vis, [(b, list(bt)) for b, bt in mu.owner.op.all_trees]))
        if la > 1:
            step.sigma = 0.9
            with pytest.raises(RuntimeError, match="fixed likelihood parameters"):
                step.astep()                          # a draw computed with the old scale is waiting
        step.close()
    v1, s1, h1 = runs[0]
    for v6, s6, h6 in runs[1:]:
        assert np.array_equal(v1, v6) and s1 == s6
        assert len(h1) == len(h6) == 2
        for (b1, bt1), (b6, bt6) in zip(h1, h6):
            assert all(np.array_equal(x, y) for x, y in zip(b1, b6)) and len(bt1) == len(bt6) == 17
            for x, y in zip(bt1, bt6):
                assert x[0] == y[0] and all(np.array_equal(p, q) for p, q in zip(x[1:], y[1:]))
    # a stop_tuning() that does not come after exactly tune_draws calls is an error, not a wrong chain
    mu = pmb.BART("mu_td", X, Y, m=20, shared_history=False)
    step = pmb.PGBART([mu], num_particles=8, chains=2, seed=71, sigma=0.8, lookahead=6, tune_draws=10)
    for _ in range(4):
        step.astep()
    step.stop_tuning()
    with pytest.raises(RuntimeError, match="tune_draws=10 does not match"):
        step.astep()
    step.close()


def test_missing_data_through_the_api():
    """tests/test_bart.py:67-81: NaNs in X[10:20, 0]; the sampler runs, the fit stays finite and follows the signal, and
    rows whose split covariate is missing are predicted from the other trees only."""
    import pymc_bart_b200 as pmb
    from pymc_bart_b200.utils import _get_posterior_sampler

    rng = np.random.default_rng(0)
    X = rng.normal(0, 1, size=(2, 50)).T.copy()
    Y = rng.normal(0, 1, size=50) + 2 * X[:, 1]
    X[10:20, 0] = np.nan
    mu = pmb.BART("mu", X, Y, m=10)
    out = pmb.sample(mu, tune=100, draws=100, chains=1, num_particles=10, seed=1, sigma=1.0)
    post = out["posterior"]
    assert post.shape == (1, 100, 50) and np.all(np.isfinite(post))
    assert np.corrcoef(post.mean(axis=(0, 1)), Y)[0, 1] > 0.6
    pred = _get_posterior_sampler(mu.owner.op).sample_posterior(X, [0, 50, 99], None)      # NaN compares false: such rows go right
    assert pred.shape == (3, 1, 50) and np.all(np.isfinite(pred))
    out["step"].close()


def test_unsupported_options_raise_not_fallback():
    import pymc_bart_b200 as pmb

    X = np.random.default_rng(0).normal(size=(20, 2)); Y = np.zeros(20)
    with pytest.warns(UserWarning):
        mu = pmb.BART("mu", X, Y, m=3, response="linear")
    with pytest.raises(NotImplementedError):
        pmb.PGBART([mu])
    with pytest.raises(NotImplementedError):
        pmb.PGBART([pmb.BART("c", X, Y, m=3)], likelihood="poisson")
    with pytest.raises(ValueError, match="No posterior draws"):
        from pymc_bart_b200.utils import _get_posterior_sampler

        op = pmb.BART("d", X, Y, m=3).owner.op
        type(op).n_outputs = 1
        _get_posterior_sampler(op)


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
