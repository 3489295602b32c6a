"""CUDA step vs the CPU oracle, bit for bit, through the C ABI.

Bar (north_star): identical tree topologies and leaf-index assignments for a
fixed RNG stream; leaf values and log-weights within 1e-5.  The fixed-point /
fixed-order arithmetic of include/bk_spec.h makes them bit-identical, so the
tests assert exact equality of: every trace record (popped node, split variable,
split value, child counts, leaf values, log-weight, resampling ancestor), the
sum of trees, the forest nodes, every row's leaf id in every tree, the
variable-inclusion counts and the running leaf sd.
"""
import numpy as np
import pytest

from helpers import assert_trace_equal, friedman
from pymc_bart_b200.settings import make_settings

pytestmark = pytest.mark.gpu


def run_pair(N, p, m, P, draws, seed, chains=1, depth_offset=0, tune_draws=None, sigma=1.0, split_rules=None, X=None, y=None):
    from oracle.oracle_py import OracleChain
    from pymc_bart_b200.core import DeviceSampler

    if X is None:
        X, y, _ = friedman(N, p, seed)
    s = make_settings(X, y, m=m, num_particles=P, seed=seed, n_chains=chains, depth_offset=depth_offset,
                      trace_capacity=20000, split_rules=split_rules)
    dev = DeviceSampler(s, X, y)
    oracles = [OracleChain(s, X.T.copy(), y, chain=c) for c in range(chains)]
    tune_draws = draws // 2 if tune_draws is None else tune_draws
    for d in range(draws):
        tune = d < tune_draws
        vi, stats = dev.step(tune, sigma)
        st_dev = dev.sum_trees().cpu().numpy()
        for c, o in enumerate(oracles):
            vio, sto = o.step(tune, sigma)
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
        assert np.array_equal(nn, nno)
        assert np.array_equal(nd.view(np.uint8), no.view(np.uint8))
        assert np.array_equal(dev.leaf_ids(c), o.leaf_ids())
    dev.close()
    return True


def test_config1_120_draws():
    """BASELINE.json configs[0]: N=200 p=5 m=10 P=20, 1 chain."""
    assert run_pair(200, 5, 10, 20, 120, seed=1)


def test_ragged_rows_and_two_chains():
    """N not a multiple of the 256-row warp tile; two chains batched in one launch."""
    assert run_pair(777, 7, 12, 9, 40, seed=3, chains=2)


def test_historical_depth_prior():
    assert run_pair(300, 4, 6, 16, 40, seed=5, depth_offset=1)


def test_tiny_and_wide():
    assert run_pair(3, 2, 3, 4, 30, seed=7)
    assert run_pair(64, 40, 20, 5, 30, seed=8)


def test_medium_many_tiles():
    """Several warp tiles and particle groups, deeper trees (sigma small)."""
    assert run_pair(5000, 10, 20, 40, 12, seed=11, sigma=0.5)


def test_onehot_rule_integer_covariates():
    rng = np.random.default_rng(12345)
    Y = np.repeat(np.arange(3), 30).astype(np.float32)
    X = np.concatenate([Y[:, None], rng.integers(0, 6, size=(90, 4))], axis=1).astype(np.float32)
    assert run_pair(90, 5, 4, 10, 60, seed=13, split_rules=["OneHotSplit"] * 5, X=X, y=Y)
