"""SubsetSplit (docs/api_reference.rst:16 `SubsetSplitRule`, pymc_bart/bart.py:103) on the CPU side:
the normative scalar definitions of include/bk_spec.h against a Python model, the oracle's trees, the host's category
tables.  The reference never exercises the rule in its own tests (tests/test_bart.py:143-147 lists "ContinuousSplit"
twice), so what is pinned here is the historical rule: left = members whose value is in a uniformly drawn non-empty
subset of the node's unique values without the largest one."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_py
from oracle.oracle_py import OracleChain
from pymc_bart_b200 import _cabi
from pymc_bart_b200.settings import encode_subset_columns, make_settings, subset_category_tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def subset_draw_model(present: int, r: int) -> int:
    """Python statement of bk_subset_draw."""
    cats = [c for c in range(24) if (present >> c) & 1]
    if len(cats) < 2:
        return 0
    cand = cats[:-1]
    n_sub = (1 << len(cand)) - 1
    pick = ((r * n_sub) >> 32) + 1
    return sum(1 << c for i, c in enumerate(cand) if (pick >> i) & 1)


def _probe(cases):
    src = r'''
#include <stdio.h>
#include <stdlib.h>
#include "bk_spec.h"
int main(int argc, char** argv){
  for (int i = 1; i + 1 < argc; i += 2) {
    uint32_t present = (uint32_t)strtoul(argv[i], 0, 0), r = (uint32_t)strtoul(argv[i + 1], 0, 0);
    printf("%u\n", bk_subset_draw(present, r));
  }
  printf("%d %d %d %d %d %d\n", bk_subset_code(0.0f), bk_subset_code(23.0f), bk_subset_code(24.0f), bk_subset_code(-1.0f), bk_subset_code(2.5f), bk_subset_code(NAN));
  printf("%d %d %d %d\n", bk_subset_left(3.0f, 8.0f), bk_subset_left(3.0f, 7.0f), bk_subset_left(31.0f, 16777215.0f), bk_subset_left(NAN, 16777215.0f));
  return 0; }
'''
    d = os.path.join(ROOT, "oracle", "_probe")
    os.makedirs(d, exist_ok=True)
    cfile, exe = os.path.join(d, "subset_probe.c"), os.path.join(d, "subset_probe")
    open(cfile, "w").write(src)
    subprocess.run(["gcc", "-O2", "-march=x86-64-v3", "-ffp-contract=off", f"-I{ROOT}/include", cfile, "-o", exe, "-lm"], check=True)
    args = [str(v) for pr in cases for v in pr]
    return subprocess.run([exe] + args, check=True, capture_output=True, text=True).stdout.split("\n")


def test_subset_draw_matches_the_model_and_is_uniform():
    rng = np.random.default_rng(5)
    cases = [(int(rng.integers(0, 1 << 24)), int(rng.integers(0, 1 << 32))) for _ in range(300)]
    cases += [(0, 123), (1 << 7, 99), (0b11, 0), (0b11, 0xFFFFFFFF), (0xFFFFFF, 0), (0xFFFFFF, 0xFFFFFFFF), (0b101001, 0x80000000)]
    out = _probe(cases)
    for (present, r), line in zip(cases, out):
        got = int(line)
        assert got == subset_draw_model(present, r), (present, r)
        if bin(present).count("1") >= 2:
            top = present.bit_length() - 1
            assert got != 0 and got & ~present == 0 and not (got >> top) & 1   # non-empty, only present categories, never the largest
        else:
            assert got == 0                                                    # fewer than two categories: no split
    assert out[len(cases)].split() == ["0", "23", "-1", "-1", "-1", "-1"]
    assert out[len(cases) + 1].split() == ["1", "0", "0", "0"]
    # every split of {1, 4, 9, 20} into two non-empty groups is drawn equally often over an even grid of uniforms
    present = (1 << 1) | (1 << 4) | (1 << 9) | (1 << 20)
    counts = {}
    for r in range(0, 1 << 32, 1 << 18):
        s = subset_draw_model(present, r)
        counts[s] = counts.get(s, 0) + 1
    assert len(counts) == 7 and max(counts.values()) - min(counts.values()) <= 1


def test_category_tables_and_encoding():
    X = np.array([[0.5, 10.0], [1.5, 30.0], [2.5, np.nan], [3.5, 10.0], [4.5, 20.0]])
    rules = np.array([_cabi.BK_RULE_CONTINUOUS, _cabi.BK_RULE_SUBSET], dtype=np.int32)
    tables = subset_category_tables(X, rules)
    assert list(tables) == [1] and np.array_equal(tables[1], [10.0, 20.0, 30.0])
    enc = encode_subset_columns(X, tables)
    assert np.array_equal(enc[:, 0], X[:, 0])
    assert np.array_equal(enc[[0, 1, 3, 4], 1], [0.0, 2.0, 0.0, 1.0]) and np.isnan(enc[2, 1])
    new = encode_subset_columns(np.array([[0.0, 20.0], [0.0, 25.0], [0.0, np.nan]]), tables)
    assert new[0, 1] == 1.0 and new[1, 1] == 31.0 and np.isnan(new[2, 1])     # unseen value: a code that is in no set
    assert encode_subset_columns(X, {}) is X
    with pytest.raises(NotImplementedError):
        subset_category_tables(np.arange(60.0).reshape(30, 2), np.array([2, 0]))   # 30 categories > 24
    s = make_settings(np.zeros((4, 2)), np.arange(4.0), split_rules=["ContinuousSplit", "SubsetSplit"])
    assert list(s.split_rules) == [0, 2]
    with pytest.raises(NotImplementedError):
        make_settings(np.zeros((4, 2)), np.arange(4.0) % 2, split_rules=["SubsetSplit", None], n_outputs=2, likelihood=_cabi.BK_LIK_CATEGORICAL)


def _categorical_data(N, seed, n_cat=6, nan_frac=0.0):
    rng = np.random.default_rng(seed)
    cat = rng.integers(0, n_cat, N)
    other = rng.integers(0, 4, N)
    X = np.stack([cat, rng.uniform(0, 1, N), other], axis=1).astype(np.float32)
    group = np.isin(cat, [0, 3, 5])                  # not an interval of the codes: no single x <= s split finds it
    y = (4.0 * group + rng.normal(0, 0.3, N)).astype(np.float32)
    if nan_frac:
        X[rng.uniform(size=N) < nan_frac, 0] = np.nan
    return X, y, group


def test_oracle_subset_trees():
    X, y, group = _categorical_data(500, 61)
    rules = ["SubsetSplit", "ContinuousSplit", "SubsetSplit"]
    s = make_settings(X, y, m=8, num_particles=12, seed=61, split_rules=rules, depth_offset=1)
    o = OracleChain(s, X.T.copy(), y)
    for d in range(60):
        o.step(d < 30, 0.3)
    nodes, nn = o.forest()
    ids = o.leaf_ids()
    n_subset_splits = 0
    for t in range(8):
        nd = nodes[t][: nn[t]]
        for k in np.nonzero(nd["var"] >= 0)[0]:
            lft = nd["left"][k]
            assert nd["n"][lft] + nd["n"][lft + 1] == nd["n"][k]                       # no missing values: nothing is dropped
            if nd["var"][k] in (0, 2):
                n_subset_splits += 1
                mask = int(nd["split"][k])
                assert float(mask) == float(nd["split"][k]) and 0 < mask < (1 << 24)
                assert nd["n"][lft] > 0 and nd["n"][lft + 1] > 0                       # a drawn set splits the members for real
        leaf = nd["var"] < 0
        assert np.array_equal(np.bincount(ids[t], minlength=nn[t])[leaf], nd["n"][leaf])
    assert n_subset_splits > 5
    fit = o.sum_trees()
    assert np.corrcoef(fit, 4.0 * group)[0, 1] > 0.97                                  # the category group is found
    # in-sample prediction from the exported forest walks the same sets
    pred = oracle_py.predict(nodes[None], X, [0], rules=s.split_rules)[0]
    np.testing.assert_allclose(pred, fit, atol=2e-4)
    # leaf membership follows the sets: replay every row through tree 0
    nd = nodes[0][: nn[0]]
    for i in range(0, 500, 7):
        k = 0
        while nd["var"][k] >= 0:
            v, sp = int(nd["var"][k]), nd["split"][k]
            left = (int(sp) >> int(X[i, v])) & 1 if v in (0, 2) else X[i, v] <= sp
            k = nd["left"][k] + (0 if left else 1)
        assert ids[0][i] == k


def test_oracle_subset_with_missing_values_and_single_category():
    X, y, _ = _categorical_data(400, 62, nan_frac=0.15)
    X[:, 2] = 3.0                                     # one category only: a draw of this column never splits
    s = make_settings(X, y, m=5, num_particles=10, seed=62, split_rules=["SubsetSplit", None, "SubsetSplit"], depth_offset=1)
    o = OracleChain(s, X.T.copy(), y)
    for d in range(40):
        o.step(d < 20, 0.3)
    nodes, nn = o.forest()
    ids = o.leaf_ids()
    used = np.concatenate([nodes[t]["var"][: nn[t]] for t in range(5)])
    assert 2 not in used and 0 in used
    limbo = ids == 255
    assert limbo.sum() > 0 and np.all(np.isnan(X[np.nonzero(limbo)[1], 0]))             # rows without a category leave the tree
    assert np.all(np.isfinite(o.sum_trees()))
