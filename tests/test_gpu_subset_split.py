"""SubsetSplit on the device against the CPU oracle, bit for bit (same bar as tests/test_gpu_parity.py), and through the
API with arbitrary category values and prediction on new data.

Rule (docs/api_reference.rst:16 `SubsetSplitRule`, pymc_bart/bart.py:103; include/bk_spec.h
bk_subset_*): the left child takes the members whose category is in a uniformly drawn non-empty subset of the categories
present among the node's members, the largest excluded.  On the device the presence masks of a particle's next queue
node are left by the workers beside the per-tile member counts, so the draw costs no extra pass over the rows."""
import numpy as np
import pytest

from pymc_bart_b200.settings import encode_subset_columns, make_settings, subset_category_tables
from test_gpu_parity import run_pair

pytestmark = pytest.mark.gpu


def categorical_data(N, seed, n_cat=6, nan_frac=0.0, kind="normal", single_cat_col=False):
    rng = np.random.default_rng(seed)
    cat = rng.integers(0, n_cat, N)
    other = rng.integers(0, 4, N)
    X = np.stack([cat, rng.uniform(0, 1, N), other, rng.uniform(0, 1, N)], axis=1).astype(np.float32)
    group = np.isin(cat, [0, 3, 5])
    f = 4.0 * group + 2.0 * X[:, 1]
    if kind == "normal":
        y = (f + rng.normal(0, 0.3, N)).astype(np.float32)
    else:
        y = (rng.uniform(size=N) < 1.0 / (1.0 + np.exp(-(f - 3.0)))).astype(np.float32)
    if nan_frac:
        X[rng.uniform(size=N) < nan_frac, 0] = np.nan
        X[rng.uniform(size=N) < nan_frac, 3] = np.nan
    if single_cat_col:
        X[:, 2] = 3.0
    return X, y


RULES = ["SubsetSplit", "ContinuousSplit", "SubsetSplit", "ContinuousSplit"]


def test_subset_gaussian_small_and_two_chains():
    X, y = categorical_data(600, 71)
    assert run_pair(600, 4, 6, 10, 60, seed=71, X=X, y=y, split_rules=RULES, sigma=0.3)
    X, y = categorical_data(777, 72, n_cat=24)                                          # all 24 category codes; ragged last tile
    assert run_pair(777, 4, 8, 16, 30, seed=72, X=X, y=y, split_rules=RULES, chains=2, depth_offset=1, sigma=0.3)


def test_subset_with_missing_values_and_cancelled_draws():
    """NaN categories leave the tree at a subset split like at any other; a column with a single category never
    splits (the draw is cancelled: the partition job turns into a count-only job)."""
    X, y = categorical_data(500, 73, nan_frac=0.12, single_cat_col=True)
    assert run_pair(500, 4, 5, 12, 40, seed=73, X=X, y=y, split_rules=RULES, depth_offset=1, sigma=0.3)


def test_subset_bernoulli():
    X, y = categorical_data(700, 74, kind="bernoulli")
    assert run_pair(700, 4, 6, 10, 30, seed=74, X=X, y=y, split_rules=RULES, likelihood=1)
    X, y = categorical_data(400, 75, kind="bernoulli", nan_frac=0.1)
    assert run_pair(400, 4, 4, 8, 24, seed=75, X=X, y=y, split_rules=RULES, likelihood=1, depth_offset=1)


def test_subset_many_tiles_and_particles():
    """P = 40 on 20 000 rows (79 tiles x 40 jobs of presence masks per epoch), then the bucket-count path (N > 131 072)."""
    X, y = categorical_data(20_000, 76, n_cat=12)
    assert run_pair(20_000, 4, 6, 40, 10, seed=76, X=X, y=y, split_rules=RULES, sigma=0.3, depth_offset=1, trace_capacity=40000)
    X, y = categorical_data(150_000, 77, n_cat=8)
    assert run_pair(150_000, 4, 2, 8, 6, seed=77, X=X, y=y, split_rules=RULES, depth_offset=1, sigma=0.3)


def test_subset_all_columns_and_onehot_mix():
    rng = np.random.default_rng(78)
    N = 300
    X = rng.integers(0, 5, size=(N, 8)).astype(np.float32)
    y = (np.isin(X[:, 0], [1, 4]) * 3.0 + (X[:, 7] == 2) * 2.0 + rng.normal(0, 0.3, N)).astype(np.float32)
    rules = ["SubsetSplit"] * 7 + ["OneHotSplit"]                                        # 7 of the 8 presence-mask slots in use
    assert run_pair(N, 8, 6, 12, 40, seed=78, X=X, y=y, split_rules=rules, sigma=0.3)


def test_subset_rejects_values_that_are_not_codes():
    from pymc_bart_b200.core import DeviceSampler

    X, y = categorical_data(100, 79)
    X[5, 0] = 2.5
    s = make_settings(X, y, m=3, num_particles=4, split_rules=RULES)
    with pytest.raises(RuntimeError, match="category codes"):
        DeviceSampler(s, X, y)


def test_subset_through_the_api_with_arbitrary_category_values():
    """pmb.BART(..., split_rules=["SubsetSplit", ...]) with category VALUES that are not codes (10.5, 20, 99, ...): the
    step encodes them, finds the group structure a single threshold cannot, and prediction encodes new data with the
    training tables (an unseen value belongs to no set and goes right, like np.isin)."""
    import pymc_bart_b200 as pmb
    from pymc_bart_b200.utils import _get_posterior_sampler

    rng = np.random.default_rng(80)
    N = 400
    values = np.array([-3.0, 10.5, 20.0, 99.0, 100.0, 1e6])
    cat = rng.integers(0, 6, N)
    X = np.stack([values[cat], rng.uniform(0, 1, N)], axis=1)
    group = np.isin(cat, [0, 3, 5])
    Y = 4.0 * group + rng.normal(0, 0.3, N)
    mu = pmb.BART("mu", X, Y, m=10, split_rules=["SubsetSplit", "ContinuousSplit"])
    out = pmb.sample(mu, tune=100, draws=50, chains=1, num_particles=10, seed=3, sigma=0.3)
    post = out["posterior"]
    assert post.shape == (1, 50, N) and np.all(np.isfinite(post))
    assert np.corrcoef(post.mean(axis=(0, 1)), 4.0 * group)[0, 1] > 0.97
    op = mu.owner.op
    assert list(op.subset_tables) == [0] and np.array_equal(op.subset_tables[0], np.sort(values))
    sampler = _get_posterior_sampler(op)
    idx = [0, 17, 49]
    pred = sampler.sample_posterior(X, idx, None)
    np.testing.assert_allclose(pred[:, 0, :], post[0, idx, :], atol=3e-4, rtol=0)        # in-sample prediction = the draws
    Xn = X[:8].copy()
    Xn[:4, 0] = 12345.0                                                                  # a category the training data never showed
    pn = sampler.sample_posterior(Xn, idx, None)
    assert pn.shape == (3, 1, 8) and np.all(np.isfinite(pn))
    np.testing.assert_allclose(pn[:, 0, 4:], post[0, idx, 4:8], atol=3e-4, rtol=0)
    # the oracle restatement predicts the same numbers from the same history (codes in, sets in the nodes)
    from oracle import oracle_py
    from pymc_bart_b200.history import ChainHistory

    baseline, batches = op.all_trees[0]
    dense = ChainHistory(list(batches), baseline, op.m, 1).dense_forests()
    enc = encode_subset_columns(Xn, subset_category_tables(X, out["step"].settings.split_rules)).astype(np.float32)
    ref = oracle_py.predict(dense, enc, idx, rules=out["step"].settings.split_rules)
    assert np.array_equal(pn[:, 0, :].astype(np.float32), ref)
    out["step"].close()
