"""Generates tests/golden/*.npz with the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference itself cannot produce vectors here:
its sampler is the un-vendored `bartrs` (requirements.txt:6) and `import pymc` fails offline,
so these fixtures pin OUR restatement (parity with the reference stays "unpinned").
They let the GPU tests check the CUDA path against stored outputs even without rebuilding the oracle.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402

from helpers import friedman  # noqa: E402
from oracle.oracle_py import OracleChain  # noqa: E402
from pymc_bart_b200.settings import make_settings  # noqa: E402

CASES = {
    # name: (N, p, m, P, draws, seed, depth_offset)
    "c1_n200_p5_m10_P20": (200, 5, 10, 20, 40, 1, 0),
    "ragged_n777_p7_m12_P9": (777, 7, 12, 9, 20, 3, 0),
    "hist_n300_p4_m6_P16": (300, 4, 6, 16, 20, 5, 1),
    # Bernoulli-logit likelihood (8th field = likelihood code)
    "bern_n500_p6_m8_P12": (500, 6, 8, 12, 24, 7, 0, 1),
}


# Large / deep cases (the code paths only the big configs reach).  Arrays of N rows are stored as sha256
# digests: (N, p, m, P, draws, seed, depth_offset, sigma)
BIG_CASES = {
    "big_n200k_p6_m4_P8": (200_000, 6, 4, 8, 6, 21, 0, 1.0),
    "big_c5shape_n1m_p50_m2_P60": (1_000_000, 50, 2, 60, 3, 5, 0, 1.0),
    "deep_n2000_p5_m4_P60": (2000, 5, 4, 60, 16, 23, 1, 0.05),
}


# SubsetSplit with missing categories and a single-category column (cancelled draws): name -> (N, m, P, draws, seed, n_cat)
SUBSET_CASES = {
    "subset_n500_m5_P12": (500, 5, 12, 40, 73, 6),
}
SUBSET_RULES = ["SubsetSplit", "ContinuousSplit", "SubsetSplit", "ContinuousSplit"]


def subset_data(N, seed, n_cat):
    rng = np.random.default_rng(seed)
    cat = rng.integers(0, n_cat, N)
    X = np.stack([cat, rng.uniform(0, 1, N), np.full(N, 3.0), rng.uniform(0, 1, N)], axis=1).astype(np.float32)
    y = (4.0 * np.isin(cat, [0, 3, 5]) + 2.0 * X[:, 1] + rng.normal(0, 0.3, N)).astype(np.float32)
    X[rng.uniform(size=N) < 0.12, 0] = np.nan
    X[rng.uniform(size=N) < 0.12, 3] = np.nan
    return X, y


def run_subset_case(N, m, P, draws, seed, n_cat):
    X, y = subset_data(N, seed, n_cat)
    s = make_settings(X, y, m=m, num_particles=P, seed=seed, depth_offset=1, trace_capacity=20000, split_rules=SUBSET_RULES)
    o = OracleChain(s, X.T.copy(), y)
    traces, sums, vis = [], [], []
    for d in range(draws):
        vi, st = o.step(d < draws // 2, 0.3)
        traces.append(o.trace().copy())
        sums.append(o.sum_trees().copy())
        vis.append(vi.copy())
    nodes, nn = o.forest()
    return dict(trace=np.concatenate(traces), trace_len=np.array([len(t) for t in traces]), sum_trees=np.stack(sums),
                vi=np.stack(vis), forest=nodes, forest_nn=nn, leaf_ids=o.leaf_ids())


def _sha(a):
    import hashlib

    return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


def run_big_case(N, p, m, P, draws, seed, depth_offset, sigma):
    X, y, _ = friedman(N, p, seed)
    s = make_settings(X, y, m=m, num_particles=P, seed=seed, depth_offset=depth_offset, trace_capacity=40000)
    o = OracleChain(s, X.T.copy(), y)
    traces, shas, vis = [], [], []
    for d in range(draws):
        vi, st = o.step(d < draws // 2, sigma)
        traces.append(o.trace().copy())
        shas.append(_sha(o.sum_trees()))
        vis.append(vi.copy())
    nodes, nn = o.forest()
    return dict(trace=np.concatenate(traces), trace_len=np.array([len(t) for t in traces]), sum_trees_sha=np.array(shas),
                vi=np.stack(vis), forest=nodes, forest_nn=nn, leaf_ids_sha=np.array(_sha(o.leaf_ids())), sigma=np.array(sigma))


def run_case(N, p, m, P, draws, seed, depth_offset, lik=0):
    X, y, _ = friedman(N, p, seed, kind="bernoulli" if lik else "normal")
    s = make_settings(X, y, m=m, num_particles=P, seed=seed, depth_offset=depth_offset, trace_capacity=20000, likelihood=lik)
    o = OracleChain(s, X.T.copy(), y)
    traces, sums, vis, sds = [], [], [], []
    for d in range(draws):
        vi, st = o.step(d < draws // 2, 1.0)
        traces.append(o.trace().copy())
        sums.append(o.sum_trees().copy())
        vis.append(vi.copy())
        sds.append(np.float32(st.leaf_sd))
    nodes, nn = o.forest()
    return dict(trace=np.concatenate(traces), trace_len=np.array([len(t) for t in traces]), sum_trees=np.stack(sums),
                vi=np.stack(vis), leaf_sd=np.array(sds, dtype=np.float32), forest=nodes, forest_nn=nn, leaf_ids=o.leaf_ids())


if __name__ == "__main__":
    out = os.path.dirname(os.path.abspath(__file__))
    only = sys.argv[1:]
    for name, cfg in CASES.items():
        if only and name not in only:
            continue
        np.savez_compressed(os.path.join(out, name + ".npz"), cfg=np.array(cfg), **run_case(*cfg))
        print("wrote", name)
    for name, cfg in BIG_CASES.items():
        if only and name not in only:
            continue
        np.savez_compressed(os.path.join(out, name + ".npz"), cfg=np.array(cfg[:7]), **run_big_case(*cfg))
        print("wrote", name)
    for name, cfg in SUBSET_CASES.items():
        if only and name not in only:
            continue
        np.savez_compressed(os.path.join(out, name + ".npz"), cfg=np.array(cfg), **run_subset_case(*cfg))
        print("wrote", name)
