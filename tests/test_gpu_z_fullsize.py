"""BASELINE.json's full sizes on the GPU, through size-independent properties (the oracle needs minutes per draw
there): every tree of every chain is a consistent binary tree whose member counts add up, every row sits in a leaf,
the sum of trees equals the sum of the leaf values the rows sit in, and the device-side consistency flags stay clear.
The checker itself is validated on the oracle (tests/test_oracle.py::test_forest_invariants_hold_for_the_oracle)."""
import numpy as np
import pytest

from helpers import check_forest_invariants
from pymc_bart_b200.settings import make_settings

pytestmark = pytest.mark.gpu


def _run(N, p, m, P, chains, draws, seed, lik=0):
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import friedman as bench_friedman
    from pymc_bart_b200.core import DeviceSampler

    X, y = bench_friedman(N, p, seed, lik)
    s = make_settings(X, y, m=m, num_particles=P, seed=seed, n_chains=chains, likelihood=lik)
    dev = DeviceSampler(s, X, y)
    grow = 0
    for d in range(draws):
        vi, st = dev.step(d < draws // 2, 1.0)
        for c in range(chains):
            assert st[c].error_flags == 0 and st[c].tree_updates == s.batch_tune
            grow += st[c].grow_events
            if d >= draws // 2:   # post-tuning: the inclusion counts are the split nodes of the trees just rewritten
                lo = (d * s.batch_tune) % m
                nodes, nn = dev.trees(c, lo, s.batch_tune)
                assert int(vi[c].sum()) == sum(int((nodes[t]["var"][: nn[t]] >= 0).sum()) for t in range(s.batch_tune))
    assert grow > 0
    st_host = dev.sum_trees().cpu().numpy()
    for c in range(chains):
        nodes, nn = dev.forest(c)
        check_forest_invariants(nodes, nn, dev.leaf_ids(c), st_host[c], N)
    dev.close()


def test_config2_full_size_properties():
    """configs[1]: N=100k p=10 m=50 P=40, 4 chains."""
    _run(100_000, 10, 50, 40, 4, 12, seed=2)


def test_config3_full_size_properties_bernoulli():
    """configs[2]: Bernoulli-logit, N=50k p=20 m=100 P=40."""
    _run(50_000, 20, 100, 40, 2, 8, seed=3, lik=1)


def test_config5_full_size_properties():
    """configs[4] per GPU: N=1M p=50 m=200 P=60, one chain (the two-level member search, HBM-resident working set)."""
    _run(1_000_000, 50, 200, 60, 1, 4, seed=5)
