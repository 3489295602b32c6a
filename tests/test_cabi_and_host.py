"""No-GPU checks: the C-ABI library loads and exports every symbol include/pgbart_b200.h declares,
pure-host entry points work, and the host logic (settings, BART op mirror, history rebuild)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from pymc_bart_b200 import BART, _cabi
from pymc_bart_b200.settings import choose_qshift, depth_prior_table, make_settings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    lib = _cabi.load()
    hdr = open(os.path.join(ROOT, "include", "pgbart_b200.h")).read()
    declared = set(re.findall(r"\b(bk_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.bk_abi_version() == _cabi.BK_ABI_VERSION


def test_padded_rows_and_query_bytes_are_pure_host():
    lib = _cabi.load()
    assert [lib.bk_padded_rows(n) for n in (0, 1, 256, 257, 100_000)] == [0, 256, 256, 512, 100_096]
    X = np.zeros((1000, 4)); Y = np.arange(1000.0)
    s = make_settings(X, Y, m=20, num_particles=16, n_chains=2)
    cs = s.to_c()
    nbytes = C.c_size_t()
    assert lib.bk_query_bytes(C.byref(cs), C.byref(nbytes)) == 0
    assert nbytes.value > 2 * 20 * 1024        # at least the per-tree leaf-id rows
    cs.n_particles = 1
    assert lib.bk_query_bytes(C.byref(cs), C.byref(nbytes)) == -1
    assert b"n_particles" in lib.bk_last_error()
    cs.n_particles = 16
    cs.likelihood = 7
    assert lib.bk_query_bytes(C.byref(cs), C.byref(nbytes)) == -5     # unsupported family: error, never a CPU fallback


def test_struct_layouts_match_the_header():
    assert C.sizeof(_cabi.BkTraceRec) == 64 and C.sizeof(_cabi.BkNode) == 24
    assert C.sizeof(_cabi.BkStepStats) == 64
    assert _cabi.BkSettings.p_leaf.offset % 8 == 0


def test_depth_prior_tables():
    t = depth_prior_table(0.95, 2.0)                        # pymc_bart/bart.py:107-109
    assert t[0] == pytest.approx(0.05) and t[1] == pytest.approx(1 - 0.95 / 4) and t[2] == pytest.approx(1 - 0.95 / 9)
    h = depth_prior_table(0.95, 2.0, depth_offset=1)        # historical indexing
    assert h[0] == 0.0 and h[1] == pytest.approx(0.05) and h[2] == pytest.approx(1 - 0.95 / 4)
    assert np.all(np.diff(t) >= 0) and t.shape == (256,)


def test_settings_from_op_attributes():
    rng = np.random.default_rng(0)
    X = rng.normal(size=(50, 3)); Y = rng.normal(size=50)
    s = make_settings(X, Y, m=10, num_particles=5, batch=(0.1, 0.3))
    assert (s.batch_tune, s.batch_post) == (1, 3)
    assert s.init_leaf == pytest.approx(Y.mean() / 10, rel=1e-6)
    assert s.leaf_sd_init == pytest.approx(Y.std() / np.sqrt(10), rel=1e-6)
    assert 2.0 ** s.qshift * 4 * np.abs(Y).max() <= 2 ** 29
    sb = make_settings(X, (Y > 0).astype(float), m=9)
    assert sb.leaf_sd_init == pytest.approx(1.0)           # 3/sqrt(m) for 0/1 data
    assert choose_qshift(30.0) == 22
    with pytest.raises(NotImplementedError):
        make_settings(X, Y, split_rules=["NoSuchRule"] * 3)
    r = make_settings(X, Y, split_rules=["ContinuousSplit", "OneHotSplit", "SubsetSplit"]).split_rules
    assert r.tolist() == [0, 1, 2]                          # tests/test_bart.py:143-145; docs/api_reference.rst:16
    with pytest.raises(ValueError):
        make_settings(X, Y, alpha=1.5)


def test_bart_op_mirror_attributes():
    X = np.zeros((50, 2)); Y = np.zeros(50)
    mu = BART("x", X=X, Y=Y)                                # tests/test_bart.py:126-137
    op = mu.owner.op
    assert type(op).__name__ == "BART_x" and op.name == "BART"
    for attr in ("X", "Y", "m", "alpha", "beta", "response", "split_prior", "split_rules", "initval", "all_trees"):
        assert hasattr(op, attr), attr                      # pymc_bart/bart.py:141-158
    assert op.m == 50 and op.alpha == 0.95 and op.beta == 2.0 and op.initval == 0.0
    assert op.split_prior.size == 0 and op.X.dtype == np.float64
    assert np.array_equal(op.rng_fn(), np.zeros(50))        # no trees yet -> Y.mean() (bart.py:54-63)
    assert BART("a", X, Y).owner.op.all_trees is not BART("b", X, Y).owner.op.all_trees   # tests/test_bart.py:193
    import pandas as pd

    mu2 = BART("p", pd.DataFrame(X, columns=["u", "v"]), pd.Series(Y))
    assert isinstance(mu2.owner.op.X, np.ndarray)


def _leaf_tree(value):
    nd = np.zeros(1, dtype=_cabi.NODE_DTYPE)
    nd["var"] = -1; nd["value"] = value
    return nd


def test_history_is_baseline_plus_deltas():
    """ChainHistory: every tree version once, a (draw, group) -> version table; a draw costs m ints."""
    from pymc_bart_b200.history import ChainHistory, compact_forest

    m = 4
    base = (np.concatenate([_leaf_tree(t) for t in range(m)]), np.ones(m, dtype=np.int32))
    b1 = (0, np.ones((1, 2), dtype=np.int32), np.concatenate([_leaf_tree(10), _leaf_tree(11)]))     # rewrites trees 0, 1
    b2 = (3, np.ones((1, 1), dtype=np.int32), _leaf_tree(30))                                        # rewrites tree 3
    h = ChainHistory([b1, b2], base, m, 1)
    assert h.n_draws == 2 and h.ver_tbl.tolist() == [[4, 5, 2, 3], [4, 5, 2, 6]]
    d = h.dense_forests()
    assert d.shape == (2, m, _cabi.BK_MAX_NODES)
    assert d["value"][0, :, 0].tolist() == [10, 11, 2, 3] and d["value"][1, :, 0].tolist() == [10, 11, 2, 30]
    assert h.forest_sizes().tolist() == [4, 4]
    # two output groups: group-major baseline, batches carry [G][T] counts
    base2 = (np.concatenate([_leaf_tree(t) for t in range(2 * m)]), np.ones(2 * m, dtype=np.int32))
    b = (1, np.ones((2, 1), dtype=np.int32), np.concatenate([_leaf_tree(100), _leaf_tree(200)]))
    h2 = ChainHistory([b], base2, m, 2)
    assert h2.ver_tbl.tolist() == [[0, 8, 2, 3], [4, 9, 6, 7]]
    with pytest.raises(ValueError):
        ChainHistory([], (base[0], np.ones(3, dtype=np.int32)), m, 1)
    dense = np.zeros((2, _cabi.BK_MAX_NODES), dtype=_cabi.NODE_DTYPE); dense["value"][:, 0] = [7, 8]; dense["value"][1, 1] = 9
    flat, nn = compact_forest(dense, np.array([1, 2]))
    assert flat["value"].tolist() == [7, 8, 9] and nn.tolist() == [1, 2]


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pymc_bart_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("# oracle", ""), fn


def test_device_sampler_refuses_to_run_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from pymc_bart_b200.core import DeviceSampler

    X = np.zeros((10, 2)); Y = np.arange(10.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DeviceSampler(make_settings(X, Y, m=2, num_particles=3), X, Y)


class _FakeCore:
    """Stands in for core.DeviceSampler (no GPU): deterministic counters instead of a sampler, same surface."""

    instances = []

    def __init__(self, settings, X, Y):
        G = max(1, settings.n_groups)
        self.settings = settings
        self.N, self.p, self.m, self.G = settings.n_rows, settings.n_cols, settings.n_trees, G
        self.C = settings.n_chains * G
        self.calls = 0
        self.h2d_bytes = 0
        self.lower = 0
        self.closed = False
        self.chain_base = settings.chain_base
        _FakeCore.instances.append(self)

    def enable_host_output(self, enable=True):
        self.host_output = enable

    def enable_history(self, enable=True, steps_per_launch=1):
        self.history = enable

    def step(self, tune, sigma):
        self.calls += 1
        self.last_sigma = sigma
        T = self.settings.batch_tune if tune else self.settings.batch_post
        self.last = (self.lower, min(self.lower + T, self.m), tune)
        self.lower = self.last[1] if self.last[1] < self.m else 0
        vi = np.zeros((self.C, self.p), dtype=np.int32)
        if not tune:
            vi[:, 0] = np.arange(self.C) + 1          # virtual chain vc used variable 0 (vc + 1) times
        return vi, [None] * self.C

    def sum_trees_host(self):
        return np.arange(self.C * self.N, dtype=np.float32).reshape(self.C, self.N) + 1000 * self.calls

    def baseline(self):
        out = []
        for c in range(self.C):
            nodes = np.zeros(self.m, dtype=_cabi.NODE_DTYPE); nodes["var"] = -1; nodes["value"] = c
            out.append((nodes, np.ones(self.m, dtype=np.int32)))
        return out

    def history_batch(self):
        lo, hi, tune = self.last
        if tune:
            return None
        T = hi - lo
        nodes = np.zeros(self.C * T, dtype=_cabi.NODE_DTYPE); nodes["var"] = -1
        nodes["value"] = 100 * self.calls + np.repeat(np.arange(self.C), T)
        return lo, np.ones((self.C, T), dtype=np.int32), nodes

    def set_response(self, Y):
        self.response = np.atleast_2d(np.array(Y, dtype=np.float64, copy=True))

    def close(self):
        self.closed = True


def test_two_bart_variables_see_each_other_as_offsets(monkeypatch):
    """tests/test_bart.py:167-241 (two BART variables in one Normal likelihood) on the host side: each step is handed
    `observed - the other variable's current value` before it runs, from the point (step(point)) or by set_offset."""
    import pymc_bart_b200.pgbart as pg

    monkeypatch.setattr(pg, "DeviceSampler", _FakeCore)
    _FakeCore.instances.clear()
    rng = np.random.default_rng(1)
    X1 = rng.normal(size=(30, 2)); X2 = rng.normal(size=(30, 3)); Yobs = rng.normal(size=30) * 7
    mu1 = BART("mu1", X1, X1[:, 0], m=5); mu2 = BART("mu2", X2, X2[:, 1], m=5)
    s1 = pg.PGBART([mu1], num_particles=5, observed=Yobs, offset_names=["mu2"])
    s2 = pg.PGBART([mu2], num_particles=5, observed=Yobs, offset_names=["mu1"])
    assert mu1.owner.op.all_trees is not mu2.owner.op.all_trees
    point = {"mu1": np.full(30, X1[:, 0].mean()), "mu2": np.full(30, X2[:, 1].mean())}
    point, st1 = s1.step(point)
    np.testing.assert_allclose(s1.core.response[0], Yobs - X2[:, 1].mean())     # the response of step 1 = observed - mu2
    point, st2 = s2.step(point)
    np.testing.assert_allclose(s2.core.response[0], Yobs - point["mu1"])         # step 2 already sees the new mu1
    assert point["mu1"].shape == (30,) and point["mu2"].shape == (30,) and st1[0]["tune"] and st2[0]["tune"]
    # the fixed-point range covers the likelihood's data, not only the Y handed to BART
    assert 2.0 ** s1.settings.qshift * 4 * np.abs(Yobs).max() <= 2 ** 29
    s1.set_offset(np.ones(30)); np.testing.assert_allclose(s1.core.response[0], Yobs - 1.0)
    with pytest.raises(KeyError):
        s1.step({"mu1": point["mu1"]})
    with pytest.raises(ValueError):
        pg.PGBART([mu1], observed=Yobs, offset_names=["mu2"], lookahead=8)
    with pytest.raises(NotImplementedError):
        pg.PGBART([BART("b", X1, (Yobs > 0).astype(float), m=5)], likelihood="bernoulli", offset_names=["mu2"])
    with pytest.raises(RuntimeError):
        pg.PGBART([mu1]).set_offset(0.0)
    s1.close(); s2.close()


def test_sample_joint_driver_with_a_fake_core(monkeypatch):
    """pmb.sample_joint([mu1, mu2], observed=Y): the PyMC-free driver of the two-variable model (tests/test_bart.py:167-206) —
    shapes, one inclusion string per variable and draw, separate histories, every step told the other variable's value."""
    import pymc_bart_b200 as pmb
    import pymc_bart_b200.pgbart as pg

    monkeypatch.setattr(pg, "DeviceSampler", _FakeCore)
    _FakeCore.instances.clear()
    rng = np.random.default_rng(2)
    X1 = rng.normal(size=(50, 2)); X2 = rng.normal(size=(50, 3)); Y = rng.normal(size=50)
    mu1 = BART("mu1", X1, X1[:, 0], m=5); mu2 = BART("mu2", X2, X2[:, 0], m=5)
    out = pmb.sample_joint([mu1, mu2], Y, tune=4, draws=6, num_particles=5)
    assert out["posterior"]["mu1"].shape == (6, 50) and out["posterior"]["mu2"].shape == (6, 50)      # idata.posterior["mu1"]: (1, draws, 50)
    assert len(out["variable_inclusion"]["mu1"]) == 6 and all(isinstance(s, str) for s in out["variable_inclusion"]["mu2"])
    s1, s2 = out["steps"]["mu1"], out["steps"]["mu2"]
    assert s1.offset_names == ["mu2"] and s2.offset_names == ["mu1"] and not s1.tune
    assert len(mu1.owner.op.all_trees) == 1 and len(mu2.owner.op.all_trees) == 1 and mu1.owner.op.all_trees is not mu2.owner.op.all_trees
    np.testing.assert_allclose(s2.core.response[0], Y - out["posterior"]["mu1"][-1], atol=1e-4)   # the last step of mu2 saw the last mu1
    with pytest.raises(ValueError):
        pmb.sample_joint([mu1, mu1], Y)
    s1.close(); s2.close()


def test_pgbart_step_protocol_with_a_fake_core(monkeypatch):
    """Host logic of PGBART.astep without a GPU: value shapes for (chains, output groups), one inclusion string per
    BART variable (groups summed), round-robin tree batches, ONE all_trees entry per chain that grows by one batch per
    posterior draw, pickling without device state, a fresh chain when tuning starts again."""
    import cloudpickle
    import pickle

    import pymc_bart_b200.pgbart as pg
    from pymc_bart_b200.history import ChainHistory
    from pymc_bart_b200.utils import _decode_vi

    monkeypatch.setattr(pg, "DeviceSampler", _FakeCore)
    _FakeCore.instances.clear()
    rng = np.random.default_rng(0)
    X = rng.normal(size=(30, 3)); Y = rng.normal(size=(2, 30))
    mu = BART("w", X, Y, m=20, shape=(2, 30), separate_trees=True)
    step = pg.PGBART([mu], num_particles=4, chains=3, batch=(0.1, 0.25))     # 2 trees per tuning draw, 5 after
    assert step.tune and step.core is None and type(mu.owner.op).n_outputs == 2      # no device state before the first step
    v, st = step.astep()
    assert step.core.host_output and step.core.history
    assert v.shape == (3, 2, 30) and len(st) == 3 and st[0] == {"variable_inclusion": "AAAA", "tune": True}
    assert v[1, 1, 0] == 1000 + (1 * 2 + 1) * 30                           # chain-major, group-minor rows of the core
    v2, _ = step.astep()
    assert v[0, 0, 0] == 1000 and v2[0, 0, 0] == 2000                        # a fresh array every draw
    op = mu.owner.op
    assert len(op.all_trees) == 0                                            # nothing is published while tuning
    step.stop_tuning()
    for d in range(5):
        v, st = step.astep()
        step.flush_history()                                                 # (Manager appends run on a writer thread)
        assert len(op.all_trees) == 3 and len(op.all_trees[0][1]) == d + 1   # one entry per chain (utils.py:117), one batch per draw
    assert [_decode_vi(s["variable_inclusion"], 3)[0] for s in st] == [1 + 2, 3 + 4, 5 + 6]   # groups of a chain summed
    base, batches = op.all_trees[2]
    batches = list(batches)
    assert [b[0] for b in batches] == [4, 9, 14, 19, 0]                      # 2 tuning draws x 2 trees, then 5 per draw ...
    assert [b[1].shape for b in batches] == [(2, 5), (2, 5), (2, 5), (2, 1), (2, 5)]   # ... the batch that reaches m is cut there (B10)
    assert base[1].shape == (2 * 20,) and base[0]["value"][0] == 4 and base[0]["value"][20] == 5   # (chain 2, groups 0 and 1) = virtual chains 4, 5
    assert batches[0][2]["value"].tolist() == [304] * 5 + [305] * 5
    h = ChainHistory(batches, base, 20, 2)                                   # the published entry is a valid history
    assert h.n_draws == 5 and h.ver_tbl.shape == (10, 20)
    # pickling: no device state crosses a process boundary (PyMC pickles the step into its worker processes)
    clone = pickle.loads(cloudpickle.dumps(step))
    assert clone.core is None and clone._batches is None and clone.chains == 3 and clone.op.m == 20
    # PyMC re-using the step object for the next chain (cores=1): tune goes back to True
    first_core = step.core
    step.tune = True
    step.astep()
    assert first_core.closed and step.core is not first_core and step.core.chain_base == 3 and step.chain_base == 3
    step.stop_tuning(); step.astep(); step.flush_history()
    assert len(op.all_trees) == 6                                            # the new chains' entries follow the old ones
    with pytest.raises(KeyError):
        pg.PGBART([mu], sigma_name="sigma").step({"sigma_log__": 0.0})       # a scale that is not in the point must not be ignored
    s2 = pg.PGBART([BART("z", X, Y[0], m=5)], sigma_name="sigma_log__", sigma_transform=np.exp)
    s2.step({"sigma_log__": np.log(2.5)})
    assert s2.core.last_sigma == pytest.approx(2.5)
    with pytest.raises(ValueError):
        pg.PGBART([mu, mu])
    with pytest.raises(TypeError):
        pg.PGBART([object()])
    step.close()
