"""No-GPU checks: the C-ABI library loads and exports every symbol include/pgbart_b200.h declares,
pure-host entry points work, and the host logic (settings, BART op mirror, history rebuild)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from pymc_bart_b200 import BART, _cabi
from pymc_bart_b200.settings import choose_qshift, depth_prior_table, make_settings
from pymc_bart_b200.utils import PosteriorSampler

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
        make_settings(X, Y, split_rules=["SubsetSplit"] * 3)
    r = make_settings(X, Y, split_rules=["ContinuousSplit", "OneHotSplit", "ContinuousSplit"]).split_rules
    assert r.tolist() == [0, 1, 0]                          # tests/test_bart.py:143-145
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


def test_history_rebuild_is_baseline_plus_deltas():
    m = 4
    base = np.zeros((m, _cabi.BK_MAX_NODES), dtype=_cabi.NODE_DTYPE)
    base["var"] = -1
    base["value"][:, 0] = np.arange(m)
    b1 = np.zeros((2, _cabi.BK_MAX_NODES), dtype=_cabi.NODE_DTYPE); b1["value"][:, 0] = [10, 11]
    b2 = np.zeros((1, _cabi.BK_MAX_NODES), dtype=_cabi.NODE_DTYPE); b2["value"][:, 0] = [30]
    f = PosteriorSampler.rebuild_forests([(0, b1, None), (3, b2, None)], (base, None), m)
    assert f.shape == (2, m, _cabi.BK_MAX_NODES)
    assert f["value"][0, :, 0].tolist() == [10, 11, 2, 3]
    assert f["value"][1, :, 0].tolist() == [10, 11, 2, 30]


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

    def __init__(self, settings, X, Y):
        G = max(1, settings.n_groups)
        self.N, self.p, self.m, self.G = settings.n_rows, settings.n_cols, settings.n_trees, G
        self.C = settings.n_chains * G
        self.calls = 0
        self.h2d_bytes = 0

    def enable_host_output(self, enable=True):
        self.host_output = enable

    def step(self, tune, sigma):
        self.calls += 1
        vi = np.zeros((self.C, self.p), dtype=np.int32)
        if not tune:
            vi[:, 0] = np.arange(self.C) + 1          # virtual chain vc used variable 0 (vc + 1) times
        return vi, [None] * self.C

    def sum_trees_host(self):
        return np.arange(self.C * self.N, dtype=np.float32).reshape(self.C, self.N) + 1000 * self.calls

    def forest(self, c):
        nodes = np.zeros((self.m, _cabi.BK_MAX_NODES), dtype=_cabi.NODE_DTYPE)
        nodes["value"][:, 0] = c
        return nodes, np.ones(self.m, dtype=np.int32)

    def trees(self, c, first, count):
        nodes = np.zeros((count, _cabi.BK_MAX_NODES), dtype=_cabi.NODE_DTYPE)
        nodes["value"][:, 0] = 100 * self.calls + c
        return nodes, np.ones(count, dtype=np.int32)

    def close(self):
        pass


def test_pgbart_step_protocol_with_a_fake_core(monkeypatch):
    """Host logic of PGBART.astep without a GPU: value shapes for (chains, output groups), one inclusion string per
    BART variable (groups summed), round-robin tree batches in the history, one all_trees entry per chain."""
    import pymc_bart_b200.pgbart as pg
    from pymc_bart_b200.utils import _decode_vi

    monkeypatch.setattr(pg, "DeviceSampler", _FakeCore)
    rng = np.random.default_rng(0)
    X = rng.normal(size=(30, 3)); Y = rng.normal(size=(2, 30))
    mu = BART("w", X, Y, m=20, shape=(2, 30), separate_trees=True)
    step = pg.PGBART([mu], num_particles=4, chains=3, batch=(0.1, 0.25))     # 2 trees per tuning draw, 5 after
    assert step.tune and step.core.host_output and type(mu.owner.op).n_outputs == 2
    v, st = step.astep()
    assert v.shape == (3, 2, 30) and len(st) == 3 and st[0] == {"variable_inclusion": "AAAA", "tune": True}
    assert v[1, 1, 0] == 1000 + (1 * 2 + 1) * 30                           # chain-major, group-minor rows of the core
    v2, _ = step.astep()
    assert v[0, 0, 0] == 1000 and v2[0, 0, 0] == 2000                        # a fresh array every draw
    step.stop_tuning()
    for d in range(5):
        v, st = step.astep()
    assert [_decode_vi(s["variable_inclusion"], 3)[0] for s in st] == [1 + 2, 3 + 4, 5 + 6]   # groups of a chain summed
    firsts = [b[0] for b in step._batches[0]]
    assert firsts == [4, 9, 14, 19, 0]                                       # 2 tuning draws x 2 trees, then 5 per draw ...
    assert [b[1].shape[0] for b in step._batches[0]] == [5, 5, 5, 1, 5]      # ... the batch that reaches m is cut there (B10)
    step.publish_history(); step.publish_history()
    op = mu.owner.op
    assert len(op.all_trees) == 3                                            # one entry per chain (utils.py:117), published once
    base, batches = op.all_trees[2]
    assert len(base) == 2 and len(batches) == 2 and len(batches[1]) == 5    # per output group inside the chain's entry
    assert base[1][0]["value"][0, 0] == 2 * 2 + 1                            # baseline of (chain 2, group 1) = virtual chain 5
    with pytest.raises(ValueError):
        pg.PGBART([mu, mu])
    with pytest.raises(TypeError):
        pg.PGBART([object()])
    step.close()
