"""The CPU oracle against (a) its committed golden fixtures, (b) published Philox known answers,
(c) libm for the hand-written transcendental kernels, (d) the reference's statistical tests
re-expressed without PyMC (tests/test_bart.py:44-64 VI dominance; tests/test_utils.py:24-32
prediction self-consistency)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from helpers import assert_trace_equal, friedman
from oracle import oracle_py
from oracle.oracle_py import OracleChain
from pymc_bart_b200 import _cabi
from pymc_bart_b200.settings import make_settings
from pymc_bart_b200.utils import _decode_vi, _encode_vi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("name", ["c1_n200_p5_m10_P20", "ragged_n777_p7_m12_P9", "hist_n300_p4_m6_P16", "bern_n500_p6_m8_P12"])
def test_oracle_matches_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    N, p, m, P, draws, seed, off = [int(v) for v in g["cfg"][:7]]
    lik = int(g["cfg"][7]) if len(g["cfg"]) > 7 else 0
    X, y, _ = friedman(N, p, seed, kind="bernoulli" if lik else "normal")
    s = make_settings(X, y, m=m, num_particles=P, seed=seed, depth_offset=off, trace_capacity=20000, likelihood=lik)
    o = OracleChain(s, X.T.copy(), y)
    pos = 0
    for d in range(draws):
        vi, st = o.step(d < draws // 2, 1.0)
        n = int(g["trace_len"][d])
        assert_trace_equal(o.trace(), g["trace"][pos:pos + n], f"{name} draw {d}")
        pos += n
        assert np.array_equal(o.sum_trees().view(np.uint32), g["sum_trees"][d].view(np.uint32))
        assert np.array_equal(vi, g["vi"][d])
    nodes, nn = o.forest()
    assert np.array_equal(nn, g["forest_nn"]) and np.array_equal(nodes.view(np.uint8), g["forest"].view(np.uint8))
    assert np.array_equal(o.leaf_ids(), g["leaf_ids"])


@pytest.mark.parametrize("name", ["big_n200k_p6_m4_P8", "big_c5shape_n1m_p50_m2_P60", "deep_n2000_p5_m4_P60"])
def test_oracle_matches_big_golden(name):
    """The digests of the large / deep cases (tests/golden/make_golden.py BIG_CASES) that pin the CUDA paths only
    BASELINE.json's big configs reach."""
    import hashlib

    def sha(a):
        return hashlib.sha256(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()

    g = np.load(os.path.join(GOLD, name + ".npz"))
    N, p, m, P, draws, seed, off = [int(v) for v in g["cfg"][:7]]
    X, y, _ = friedman(N, p, seed)
    s = make_settings(X, y, m=m, num_particles=P, seed=seed, depth_offset=off, trace_capacity=40000)
    o = OracleChain(s, X.T.copy(), y)
    pos = 0
    for d in range(draws):
        vi, st = o.step(d < draws // 2, float(g["sigma"]))
        n = int(g["trace_len"][d])
        assert_trace_equal(o.trace(), g["trace"][pos:pos + n], f"{name} draw {d}")
        pos += n
        assert sha(o.sum_trees()) == str(g["sum_trees_sha"][d])
    assert sha(o.leaf_ids()) == str(g["leaf_ids_sha"])


def test_oracle_matches_subset_golden():
    """SubsetSplit with missing categories and a single-category column (tests/golden/make_golden.py SUBSET_CASES)."""
    sys.path.insert(0, GOLD)
    from make_golden import SUBSET_RULES, subset_data

    name = "subset_n500_m5_P12"
    g = np.load(os.path.join(GOLD, name + ".npz"))
    N, m, P, draws, seed, n_cat = [int(v) for v in g["cfg"]]
    X, y = subset_data(N, seed, n_cat)
    s = make_settings(X, y, m=m, num_particles=P, seed=seed, depth_offset=1, trace_capacity=20000, split_rules=SUBSET_RULES)
    o = OracleChain(s, X.T.copy(), y)
    pos = 0
    for d in range(draws):
        vi, st = o.step(d < draws // 2, 0.3)
        n = int(g["trace_len"][d])
        assert_trace_equal(o.trace(), g["trace"][pos:pos + n], f"{name} draw {d}")
        pos += n
        assert np.array_equal(o.sum_trees().view(np.uint32), g["sum_trees"][d].view(np.uint32))
        assert np.array_equal(vi, g["vi"][d])
    nodes, nn = o.forest()
    assert np.array_equal(nn, g["forest_nn"]) and np.array_equal(nodes.view(np.uint8), g["forest"].view(np.uint8))
    assert np.array_equal(o.leaf_ids(), g["leaf_ids"])
    used = np.concatenate([nodes[t]["var"][: nn[t]] for t in range(m)])
    assert 0 in used and 2 not in used and (o.leaf_ids() == 255).any()      # sets drawn, the one-category column never, NaN rows dropped


def test_oracle_missing_covariates():
    """SURVEY.md App. A.4 in the restatement: NaN candidates are skipped, rows with a missing split covariate go to
    limbo (id 255) and count for neither child, a column that is missing everywhere is never split on."""
    X, _, _ = friedman(400, 5, 41)
    X = X.copy()
    y = (6 * X[:, 0] + np.random.default_rng(41).normal(0, 0.5, 400)).astype(np.float32)   # column 0 carries the signal
    X[50:150, 0] = np.nan
    X[:, 3] = np.nan
    s = make_settings(X, y, m=6, num_particles=10, seed=41, depth_offset=1)
    o = OracleChain(s, X.T.copy(), y)
    for d in range(40):
        o.step(d < 20, 1.0)
    nodes, nn = o.forest()
    ids = o.leaf_ids()
    used = np.unique(np.concatenate([nodes[t]["var"][: nn[t]] for t in range(6)]))
    assert 3 not in used and 0 in used and np.all(np.isfinite(o.sum_trees()))
    assert (ids == 255).sum() > 0
    for t in range(6):
        nd = nodes[t][: nn[t]]
        limbo = ids[t] == 255
        assert np.all(np.isnan(X[limbo, 0]))                                   # only rows without the split covariate are dropped
        split = np.nonzero(nd["var"] >= 0)[0]
        kids = nd["n"][nd["left"][split]] + nd["n"][nd["left"][split] + 1]
        assert np.all(kids <= nd["n"][split]) and nd["n"][0] == 400
        leaf = nd["var"] < 0
        assert np.array_equal(np.bincount(ids[t][~limbo], minlength=nn[t])[leaf], nd["n"][leaf])


def _spec_probe():
    """Small C program over include/bk_spec.h: Philox KATs + max error of the math kernels vs libm."""
    src = r'''
#include <stdio.h>
#include <stdlib.h>
#include "bk_spec.h"
int main(void){
  bk_u32x4 a=bk_philox(0,0,0,0,0,0), b=bk_philox(0xffffffffu,0xffffffffu,0xffffffffu,0xffffffffu,0xffffffffu,0xffffffffu);
  bk_u32x4 c=bk_philox(0xa4093822u,0x299f31d0u,0x243f6a88u,0x85a308d3u,0x13198a2eu,0x03707344u);
  printf("%08x %08x %08x %08x\n%08x %08x %08x %08x\n%08x %08x %08x %08x\n",a.v[0],a.v[1],a.v[2],a.v[3],b.v[0],b.v[1],b.v[2],b.v[3],c.v[0],c.v[1],c.v[2],c.v[3]);
  double me=0,ml=0,mc=0; srand(7);
  for(int i=0;i<400000;i++){ double x=-700.0*rand()/RAND_MAX, e=fabs(bk_exp(x)-exp(x))/exp(x); if(e>me)me=e;
    double y=exp(-40+80.0*rand()/RAND_MAX), l=fabs(bk_log(y)-log(y)); if(l>ml)ml=l;
    uint32_t u=(uint32_t)rand()*2u+(rand()&1u); double d=fabs(bk_cos2pi(u)-cos(2*M_PI*(u/4294967296.0))); if(d>mc)mc=d; }
  printf("%.3e %.3e %.3e\n",me,ml,mc);
  double s=0,s2=0; int n=400000; for(int i=0;i<n;i++){ double z=bk_normal(bk_rng(1,2,i,0,0,0,0,3)); s+=z; s2+=z*z; }
  printf("%.5f %.5f\n", s/n, s2/n);
  printf("%d %d\n", bk_quant(1.0f, 1024.0f), bk_quant(-1e30f, 1024.0f));
  double mb=0; for(int i=0;i<400000;i++){ float f=(float)(-60.0+120.0*rand()/RAND_MAX); float yy=(float)(i&1);
    double ex=(double)yy*f-(f>0?f+log1p(exp(-(double)f)):log1p(exp((double)f))); double d=fabs((double)bk_bernoulli_logit_term(yy,f)-ex)/(1.0+fabs(ex)); if(d>mb)mb=d; }
  printf("%.3e %d\n", mb, bk_bern_q(1.0f, 0.25f, -0.25f));
  return 0; }
'''
    d = os.path.join(ROOT, "oracle", "_probe")
    os.makedirs(d, exist_ok=True)
    cfile, exe = os.path.join(d, "probe.c"), os.path.join(d, "probe")
    open(cfile, "w").write(src)
    subprocess.run(["gcc", "-O2", "-march=x86-64-v3", "-ffp-contract=off", f"-I{ROOT}/include", cfile, "-o", exe, "-lm"], check=True)
    return subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split("\n")


def test_spec_header_known_answers():
    out = _spec_probe()
    # Random123 known-answer vectors for Philox4x32-10
    assert out[0] == "6627e8d5 e169c58d bc57ac4c 9b00dbd8"
    assert out[1] == "408f276d 41c83b0e a20bc7c6 6d5451fd"
    assert out[2] == "d16cfe09 94fdcceb 5001e420 24126ea1"
    me, ml, mc = [float(v) for v in out[3].split()]
    assert me < 1e-15 and ml < 1e-14 and mc < 2e-15
    mean, var = [float(v) for v in out[4].split()]
    assert abs(mean) < 0.01 and abs(var - 1) < 0.01
    assert out[5].split() == ["1024", str(-(2**29 - 1))]
    mb, q0 = out[6].split()
    assert float(mb) < 1e-6                       # Bernoulli-logit term (float kernel) vs libm in double
    assert abs(int(q0) - round(-np.log(2.0) * 2**20)) <= 1  # y=1, f=0: -log 2 in units of 2^-20


def test_vi_dominance_statistical():
    """tests/test_bart.py:44-64 without PyMC: X[:,0] ~ Y => variable 0 dominates the inclusion counts."""
    rng = np.random.default_rng(3415)
    X = rng.normal(0, 1, size=(250, 3))
    Y = rng.normal(0, 1, size=250)
    X[:, 0] = rng.normal(Y, 0.1)
    s = make_settings(X, Y, m=10, num_particles=10, seed=3415)
    o = OracleChain(s, np.ascontiguousarray(X.T, dtype=np.float32), Y.astype(np.float32))
    tot = np.zeros(3, dtype=np.int64)
    for d in range(400):
        vi, _ = o.step(d < 200, 1.0)
        if d >= 200:
            tot += np.asarray(_decode_vi(_encode_vi(vi.tolist()), 3))
    frac = tot / tot.sum()
    assert frac[0] > frac[1:].sum()


def test_fit_improves_and_is_deterministic():
    X, y, f = friedman(1500, 8, 21)
    s = make_settings(X, y, m=30, num_particles=12, seed=21)
    a = OracleChain(s, X.T.copy(), y)
    b = OracleChain(s, X.T.copy(), y)
    r0 = float(np.sqrt(np.mean((a.sum_trees() - f) ** 2)))
    for d in range(120):
        a.step(d < 60, 1.0); b.step(d < 60, 1.0)
    assert np.array_equal(a.sum_trees(), b.sum_trees())
    r1 = float(np.sqrt(np.mean((a.sum_trees() - f) ** 2)))
    assert r1 < 0.45 * r0
    # chains differ only through the Philox key
    c = OracleChain(s, X.T.copy(), y, chain=1)
    c.step(True, 1.0)
    b2 = OracleChain(s, X.T.copy(), y, chain=0)
    b2.step(True, 1.0)
    assert not np.array_equal(c.sum_trees(), b2.sum_trees())


def test_prediction_self_consistency_and_in_sample():
    """tests/test_utils.py:24-32: predicting X[:10] equals the first 10 rows of predicting X;
    additionally the stored forest reproduces the sampler's own in-sample sum of trees."""
    X, y, _ = friedman(300, 5, 9)
    s = make_settings(X, y, m=8, num_particles=8, seed=9)
    o = OracleChain(s, X.T.copy(), y)
    for d in range(30):
        o.step(d < 15, 1.0)
    nodes, _ = o.forest()
    forests = nodes[None]
    all_ = oracle_py.predict(forests, X, [0])
    first = oracle_py.predict(forests, X[:10], [0])
    assert np.array_equal(all_[0, :10], first[0])
    np.testing.assert_allclose(all_[0], o.sum_trees(), rtol=0, atol=2e-4)
    # excluding every variable collapses each tree to its training-weighted mean leaf
    ex = oracle_py.predict(forests, X[:5], [0], excluded_mask=np.ones(5, np.uint8))
    assert np.allclose(ex[0], ex[0][0])


def test_bernoulli_oracle_recovers_the_logit():
    """tests/test_bart.py:150-164 style check for the non-Gaussian path: the fitted logit tracks the truth and the
    in-sample Bernoulli log-likelihood approaches the data-generating one."""
    X, y, f = friedman(4000, 8, 3, kind="bernoulli")
    s = make_settings(X, y, m=30, num_particles=12, seed=3, likelihood=_cabi.BK_LIK_BERNOULLI_LOGIT)
    o = OracleChain(s, X.T.copy(), y)
    ll = lambda st: float(np.mean(y * st - np.logaddexp(0, st)))
    l0 = ll(o.sum_trees())
    for d in range(120):
        o.step(d < 60, 1.0)
    pr = 1 / (1 + np.exp(-(f - 14.4) / 4.9))
    l_true = float(np.mean(y * np.log(pr) + (1 - y) * np.log(1 - pr)))
    l1 = ll(o.sum_trees())
    assert l1 > l0 + 0.08 and abs(l1 - l_true) < 0.03
    assert np.corrcoef(o.sum_trees(), f)[0, 1] > 0.8
    with pytest.raises(ValueError):
        make_settings(X, y + 0.5, m=5, likelihood=_cabi.BK_LIK_BERNOULLI_LOGIT)


def test_edge_cases():
    # N smaller than a tile, constant response, two rows
    for N, p, m, P in [(2, 1, 2, 3), (5, 3, 1, 2)]:
        X = np.random.default_rng(N).uniform(size=(N, p)).astype(np.float32)
        y = np.zeros(N, dtype=np.float32) + 3.0
        s = make_settings(X, y, m=m, num_particles=P, seed=1)
        o = OracleChain(s, X.T.copy(), y)
        for d in range(10):
            vi, st = o.step(d < 5, 1.0)
            assert st.error_flags == 0
        assert np.all(np.isfinite(o.sum_trees()))


def test_forest_invariants_hold_for_the_oracle():
    """The size-independent properties the full-size GPU tests rely on (tests/test_gpu_z_fullsize.py), checked on the oracle."""
    from helpers import check_forest_invariants

    for lik in (0, 1):
        X, y, _ = friedman(3000, 8, 31, kind="bernoulli" if lik else "normal")
        s = make_settings(X, y, m=20, num_particles=16, seed=31, likelihood=lik)
        o = OracleChain(s, X.T.copy(), y)
        for d in range(40):
            o.step(d < 20, 1.0)
        nodes, nn = o.forest()
        check_forest_invariants(nodes, nn, o.leaf_ids(), o.sum_trees(), 3000)


def test_particle_threads_do_not_change_the_results():
    """bko_set_threads: the particles of a round (and the deep copies of a resampling) spread over host threads — every
    trace record, the sum of trees, the forest and the leaf ids stay bit-identical (Gaussian and Bernoulli, missing values)."""
    for lik in (0, 1):
        X, y, _ = friedman(600, 5, 88, kind="bernoulli" if lik else "normal")
        X = X.copy(); X[50:120, 1] = np.nan
        s = make_settings(X, y, m=6, num_particles=16, seed=88, trace_capacity=30000, likelihood=lik, depth_offset=1)
        a, b = OracleChain(s, X.T.copy(), y), OracleChain(s, X.T.copy(), y)
        assert b.set_threads(5) == 5 and a.set_threads(0) == 1
        for d in range(24):
            via, sta = a.step(d < 12, 0.7)
            vib, stb = b.step(d < 12, 0.7)
            assert_trace_equal(a.trace(), b.trace(), f"lik {lik} draw {d}")
            assert np.array_equal(via, vib) and sta.grow_events == stb.grow_events and sta.grow_root == stb.grow_root
            assert np.array_equal(a.sum_trees().view(np.uint32), b.sum_trees().view(np.uint32))
        assert np.array_equal(a.leaf_ids(), b.leaf_ids()) and a.bytes_touched() == b.bytes_touched()
        na, nb = a.forest(), b.forest()
        assert np.array_equal(na[0].view(np.uint8), nb[0].view(np.uint8)) and np.array_equal(na[1], nb[1])
