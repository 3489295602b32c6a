import numpy as np


def friedman(N, p, seed, kind="normal"):
    """Synthetic Friedman data of SURVEY.md §8(d)."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(0, 1, (N, p)).astype(np.float32)
    Z = np.zeros((N, 5), dtype=np.float64)  # columns missing when p < 5 count as 0
    Z[:, : min(p, 5)] = X[:, : min(p, 5)]
    f = 10 * np.sin(np.pi * Z[:, 0] * Z[:, 1]) + 20 * (Z[:, 2] - 0.5) ** 2 + 10 * Z[:, 3] + 5 * Z[:, 4]
    if kind == "normal":
        y = (f + rng.normal(0, 1, N)).astype(np.float32)
    else:
        pr = 1.0 / (1.0 + np.exp(-(f - 14.4) / 4.9))
        y = (rng.uniform(0, 1, N) < pr).astype(np.float32)
    return X, y, f


def assert_trace_equal(a, b, ctx=""):
    assert len(a) == len(b), f"{ctx}: trace length {len(a)} vs {len(b)}"
    if len(a) == 0:
        return
    ab = a.view(np.uint8).reshape(len(a), -1)
    bb = b.view(np.uint8).reshape(len(b), -1)
    if not np.array_equal(ab, bb):
        bad = np.nonzero((ab != bb).any(axis=1))[0][0]
        raise AssertionError(f"{ctx}: first differing trace record {bad}:\n gpu    {a[bad]}\n oracle {b[bad]}")


def check_forest_invariants(nodes, n_nodes, leaf_ids, sum_trees, n_rows, atol=2e-3):
    """Size-independent properties of a sampler state (one chain): every tree is a consistent binary tree whose node
    counts add up, every row sits in a leaf of every tree, and the sum of trees equals the sum of the leaf values the
    rows sit in.  nodes: [m][255] NODE_DTYPE, n_nodes: [m], leaf_ids: [m][N] uint8, sum_trees: [N] float32."""
    m = nodes.shape[0]
    total = np.zeros(n_rows, dtype=np.float64)
    for t in range(m):
        nn = int(n_nodes[t])
        nd = nodes[t][:nn]
        assert 1 <= nn <= 255 and nn % 2 == 1, (t, nn)
        assert nd["n"][0] == n_rows and nd["depth"][0] == 0
        is_leaf = nd["var"] < 0
        split = np.nonzero(~is_leaf)[0]
        left = nd["left"][split]
        assert np.all(left > split) and np.all(left + 1 < nn)                      # children are created after their parent
        assert np.array_equal(nd["n"][left] + nd["n"][left + 1], nd["n"][split])   # member counts add up
        assert np.array_equal(nd["depth"][left], nd["depth"][split] + 1) and np.array_equal(nd["depth"][left + 1], nd["depth"][split] + 1)
        kids = np.sort(np.concatenate([left, left + 1]))
        assert np.array_equal(kids, np.arange(1, nn))                              # every non-root node has exactly one parent
        ids = leaf_ids[t]
        assert ids.max() < nn and np.all(is_leaf[ids])                             # rows sit in leaves
        counts = np.bincount(ids, minlength=nn)
        assert np.array_equal(counts[is_leaf], nd["n"][is_leaf])                   # leaf membership matches the node counts
        total += nd["value"][ids].astype(np.float64)
    np.testing.assert_allclose(sum_trees.astype(np.float64), total, rtol=0, atol=atol)
