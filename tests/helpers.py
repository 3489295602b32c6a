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
