"""Oracle parity on the code paths only BASELINE.json's large configs reach ("What's weak" 1-2):

* `select_split` above 512 row tiles (N > 131 072: bucket counts, then tile counts), above 16 384 tiles (N > 4.2M:
  two-level search over the bucket counts) and at the true C5 row count / particle count with the forest cut to two trees;
* particles whose trees outgrow the shared-memory node store (`fastF` nodes per particle: 21 at P=60, 10 at P=128)
  and spill into the global overflow array;
* non-uniform `split_prior`, more than BK_CUM_SMEM=1024 columns, depth >= 64 (global depth-prior table) and the
  255-node cap of one-byte leaf ids.

Same bar as tests/test_gpu_parity.py: every trace record, the sum of trees, the forest, every leaf id, the inclusion
counts and the running leaf sd are bit-identical to the CPU oracle.
"""
import hashlib
import os

import numpy as np
import pytest

from helpers import assert_trace_equal, friedman
from pymc_bart_b200.settings import make_settings
from test_gpu_parity import run_pair

pytestmark = pytest.mark.gpu


def test_two_level_member_search_n200k():
    """782 row tiles: every grow below the root goes through the bucket-count level of select_split."""
    assert run_pair(200_000, 6, 4, 8, 6, seed=21)


def test_bucket_counts_n1100k():
    """4297 tiles -> 135 buckets of 32 tiles: bucket level in registers, then the bucket's tile counts."""
    assert run_pair(1_100_000, 3, 1, 6, 8, seed=22)


def test_bucket_counts_two_level_n4300k():
    """16797 tiles -> 525 buckets: more than 16 bucket counts per lane, so the bucket level itself takes the two-level
    branch of find_in_counts (reached only above N = 4.2M rows)."""
    assert run_pair(4_300_000, 2, 1, 4, 6, seed=28)


def test_config5_shape_two_trees():
    """configs[4] (N=1M, p=50, P=60) with the forest cut to m=2: the exact tile count, particle count and
    shared-memory split (fastF=21) of the C5 bench line, three draws against the oracle."""
    assert run_pair(1_000_000, 50, 2, 60, 3, seed=5)


def test_p60_particles_overflow_the_shared_node_store():
    """P=60 -> 21 nodes per particle in shared memory; sigma=0.05 with the historical depth prior grows trees of
    up to ~45 nodes, so node reads/writes/copies cross into the global overflow array."""
    assert run_pair(2000, 5, 4, 60, 16, seed=23, sigma=0.05, depth_offset=1, trace_capacity=40000)


def test_p128_max_particles():
    """The maximum particle count (fastF=10; four resampling slices; 256 pool rows)."""
    assert run_pair(1500, 4, 3, 128, 10, seed=24, sigma=0.05, depth_offset=1, trace_capacity=60000)


def test_non_uniform_split_prior():
    assert run_pair(400, 6, 5, 12, 30, seed=26, split_prior=[5, 1, 0.5, 3, 0.1, 2])


def test_more_columns_than_the_shared_prior_table():
    """p=1100 > BK_CUM_SMEM: the cumulative prior and the split rules of columns >= 1024 come from global memory,
    and the tuned prior rebuild takes the sequential branch."""
    assert run_pair(300, 1100, 4, 10, 20, seed=25)


def test_chain_like_trees_depth_over_64_and_node_cap():
    """OneHot splits on all-distinct columns peel one row per split: depth reaches ~120 (> 64: global depth-prior
    table) and trees hit the 255-node cap of one-byte leaf ids (grow attempts beyond it are rejected)."""
    rng = np.random.default_rng(9)
    N = 120
    X = np.stack([rng.permutation(N).astype(np.float32), rng.permutation(N).astype(np.float32)], 1)
    y = (X[:, 0] * 0.1 + rng.normal(0, 1, N)).astype(np.float32)
    assert run_pair(N, 2, 3, 8, 40, seed=27, sigma=0.02, alpha=0.999, beta=0.001, X=X, y=y, split_rules=["OneHotSplit"] * 2)


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


@pytest.mark.parametrize("name", ["big_n200k_p6_m4_P8", "big_c5shape_n1m_p50_m2_P60", "deep_n2000_p5_m4_P60"])
def test_cuda_matches_committed_big_golden(name):
    """CUDA vs committed digests (tests/golden/make_golden.py BIG_CASES): full trace, sha256 of the sum of trees per
    draw and of the final leaf ids; no oracle involved."""
    from pymc_bart_b200.core import DeviceSampler

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", name + ".npz"))
    N, p, m, P, draws, seed, off = [int(v) for v in g["cfg"][:7]]
    sigma = float(g["sigma"])
    X, y, _ = friedman(N, p, seed)
    s = make_settings(X, y, m=m, num_particles=P, seed=seed, depth_offset=off, trace_capacity=40000)
    dev = DeviceSampler(s, X, y)
    pos = 0
    for d in range(draws):
        vi, st = dev.step(d < draws // 2, sigma)
        n = int(g["trace_len"][d])
        assert_trace_equal(dev.trace(0), g["trace"][pos:pos + n], f"{name} draw {d}")
        pos += n
        assert _sha(dev.sum_trees().cpu().numpy()[0]) == str(g["sum_trees_sha"][d])
        assert np.array_equal(vi[0], g["vi"][d])
    nodes, nn = dev.forest(0)
    assert np.array_equal(nn, g["forest_nn"]) and np.array_equal(nodes.view(np.uint8), g["forest"].view(np.uint8))
    assert _sha(dev.leaf_ids(0)) == str(g["leaf_ids_sha"])
    dev.close()


def _with_missing(N, p, seed, kind="normal", all_nan_col=None):
    X, y, _ = friedman(N, p, seed, kind=kind)
    rng = np.random.default_rng(seed + 1000)
    X = X.copy()
    X[N // 6: N // 3, 0] = np.nan                              # a block of rows without column 0 (tests/test_bart.py:67-81)
    X[rng.uniform(size=N) < 0.1, 2] = np.nan                   # 10 % scattered in column 2
    if all_nan_col is not None:
        X[:, all_nan_col] = np.nan                             # every candidate is missing: the split is cancelled
    return X, y


def test_missing_covariates_gaussian():
    """NaN covariates (reference smoke test tests/test_bart.py:67-81): candidates with a missing
    value are skipped (up to BK_SPLIT_TRIES draws), rows with a missing split covariate leave the tree (leaf id 255,
    predict 0) and count for neither child."""
    X, y = _with_missing(600, 4, 31)
    assert run_pair(600, 4, 6, 10, 40, seed=31, X=X, y=y)
    assert run_pair(600, 4, 6, 10, 20, seed=32, X=X, y=y, chains=2, depth_offset=1, sigma=0.3)


def test_missing_covariates_cancelled_splits_and_bernoulli():
    """A column that is missing everywhere: every draw of it exhausts its tries and the grow is cancelled (the
    partition job turns into the count-only job / nothing).  Bernoulli: the dropped rows' log-likelihood terms."""
    X, y = _with_missing(500, 5, 33, all_nan_col=1)
    assert run_pair(500, 5, 5, 12, 30, seed=33, X=X, y=y, depth_offset=1)
    Xb, yb = _with_missing(700, 6, 34, kind="bernoulli")
    assert run_pair(700, 6, 6, 10, 30, seed=34, X=Xb, y=yb, likelihood=1)
    Xc, yc = _with_missing(400, 5, 35, kind="bernoulli", all_nan_col=3)
    assert run_pair(400, 5, 4, 8, 24, seed=35, X=Xc, y=yc, likelihood=1, depth_offset=1)


def test_missing_covariates_large_n_bucket_counts():
    X, y = _with_missing(150_000, 4, 36)
    assert run_pair(150_000, 4, 2, 8, 6, seed=36, X=X, y=y, depth_offset=1)


def test_several_steps_per_launch_equal_single_steps():
    """bk_run_launch: n steps of every chain inside ONE persistent launch (chains do not wait for each other at step
    boundaries) give bit for bit what n single-step launches give: per-step inclusion counts and stats, the draw written
    by every step's last commit, and the final forest."""
    import torch

    from pymc_bart_b200.core import DeviceSampler

    X, y, _ = friedman(3000, 6, 61)
    s = make_settings(X, y, m=12, num_particles=10, seed=61, n_chains=3)        # 1 tree per step: the batch wraps around m
    a, b = DeviceSampler(s, X, y), DeviceSampler(s, X, y)
    for tune, n in ((True, 5), (True, 16), (False, 7), (False, 1)):
        draws = torch.zeros((n, b.rows, b.ld), dtype=torch.float32, device="cuda")
        b.run_launch(n, tune, 0.7, draws_out=draws)
        vi_b, st_b = b.run_wait()
        torch.cuda.synchronize()
        for k in range(n):
            vi_a, st_a = a.step(tune, 0.7)
            assert np.array_equal(vi_a, vi_b[k])
            for c in range(3):
                assert (st_a[c].grow_events, st_a[c].rounds, st_a[c].tree_updates, st_a[c].iter) == \
                       (st_b[k][c].grow_events, st_b[k][c].rounds, st_b[k][c].tree_updates, st_b[k][c].iter)
                assert np.float32(st_a[c].leaf_sd).view(np.uint32) == np.float32(st_b[k][c].leaf_sd).view(np.uint32)
                assert st_b[k][c].error_flags == 0
            assert torch.equal(draws[k], a.sum_trees_dev)
    for c in range(3):
        na, nna = a.forest(c); nb, nnb = b.forest(c)
        assert np.array_equal(nna, nnb) and np.array_equal(na.view(np.uint8), nb.view(np.uint8))
        assert np.array_equal(a.leaf_ids(c), b.leaf_ids(c))
    with pytest.raises(RuntimeError):
        b.run_launch(17, True, 1.0)
    a.close(); b.close()


def test_draws_also_stored_into_peer_buffers():
    """bk_set_draw_peers: the commit sweep that keeps a draw also stores it into the same place of every peer buffer
    (multi-GPU runs map the other ranks' buffers here; the test uses two more buffers on the same GPU as the peers).
    bench.py checks the real thing against NCCL's all-gather at N > 1 (`config.gather_check`)."""
    import torch

    from pymc_bart_b200.core import DeviceSampler

    X, y, _ = friedman(2500, 5, 62)
    s = make_settings(X, y, m=10, num_particles=8, seed=62, n_chains=2)
    d = DeviceSampler(s, X, y)
    n_keep, world, rank = 6, 3, 1
    bufs = [torch.zeros((world * n_keep, d.rows, d.ld), dtype=torch.float32, device="cuda") for _ in range(world)]
    d.set_draw_peers([bufs[0].data_ptr(), bufs[2].data_ptr()], bufs[rank].data_ptr())
    mine = bufs[rank][rank * n_keep:(rank + 1) * n_keep]
    d.run_launch(4, True, 1.0)                       # tuning steps keep no draw
    d.run_wait()
    d.run_launch(n_keep, False, 1.0, draws_out=mine)
    d.run_wait()
    torch.cuda.synchronize()
    assert torch.equal(mine[-1], d.sum_trees_dev) and float(mine.abs().sum()) > 0
    for b in bufs:
        assert torch.equal(b[rank * n_keep:(rank + 1) * n_keep], mine)
        assert float(b[:rank * n_keep].abs().sum()) == 0 and float(b[(rank + 1) * n_keep:].abs().sum()) == 0
    d.set_draw_peers([], 0)                          # off again: only the local buffer is written
    before = bufs[0].clone()
    d.run_launch(2, False, 1.0, draws_out=mine[:2])
    d.run_wait()
    torch.cuda.synchronize()
    assert torch.equal(bufs[0], before) and torch.equal(mine[1], d.sum_trees_dev)
    with pytest.raises(RuntimeError):
        d.set_draw_peers([1] * 8, bufs[rank].data_ptr())
    d.close()
