"""CPU checks of the bench contract: the reference arm's JSON line, one workload string for both arms, the
algorithmic byte model of SURVEY.md §8(d), host-side argument checks of the C ABI, and the guarded PyMC adapter."""
import ctypes as C
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C1", "--steps", "10",
                          "--warmup", "2"], capture_output=True, text=True, check=True, cwd=ROOT).stdout.strip().split("\n")[-1]
    d = json.loads(out)
    assert d["impl"] == "reference" and d["metric"] == "PGBART draws/sec" and d["unit"] == "draws/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["value"] > 0 and d["steps"] == 10 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "draws/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    import bench

    assert d["config"]["workload"] == bench.workload_name("C1", bench.CONFIGS["C1"], 1)   # same string as our arm builds


def test_algorithmic_bytes_and_configs():
    import bench

    # 14N per grow event (+9N non-Gaussian), 23N per tree update, +16N while tuning
    assert bench.algorithmic_bytes(1000, 10, 2, 1) == 14e3 * 10 + 23e3 * 2 + 16e3 * 1
    assert bench.algorithmic_bytes(1000, 10, 2, 0, lik=1) == 23e3 * 10 + 23e3 * 2
    assert set(bench.CONFIGS) == {"C1", "C2", "C3", "C4", "C5"}
    X, y = bench.friedman(500, 20, 3, lik=1)
    assert set(np.unique(y).tolist()) == {0.0, 1.0} and 0.3 < y.mean() < 0.7          # balanced classes (§8d)
    X, Y = bench.friedman(300, 15, 4, groups=3)
    assert Y.shape == (3, 300) and X.shape == (300, 15)


def test_abi_rejects_what_the_device_cannot_do():
    from pymc_bart_b200 import _cabi
    from pymc_bart_b200.settings import choose_qshift, make_settings

    lib = _cabi.load()
    X = np.zeros((64, 3)); Y = (np.arange(64) % 2).astype(float)
    s = make_settings(X, Y, m=10, num_particles=8, likelihood=_cabi.BK_LIK_BERNOULLI_LOGIT)
    assert s.qshift == choose_qshift(16.0) and s.leaf_sd_init == pytest.approx(3 / np.sqrt(10), rel=1e-6)   # logit range, 0/1 response
    nbytes = C.c_size_t()
    cs = s.to_c()
    assert lib.bk_query_bytes(C.byref(cs), C.byref(nbytes)) == 0          # Bernoulli is a device family
    cs.n_rows = (1 << 25) + 1
    assert lib.bk_query_bytes(C.byref(cs), C.byref(nbytes)) == -1 and b"2^25" in lib.bk_last_error()
    cs.n_rows = 64
    cs.n_chains = 65
    assert lib.bk_query_bytes(C.byref(cs), C.byref(nbytes)) == -1
    assert lib.bk_set_host_output(None, 1) == -1 and not lib.bk_sum_trees_host(None)


def test_pymc_adapter_is_guarded():
    try:
        import pymc  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="needs PyMC"):
            import pymc_bart_b200.pymc_adapter  # noqa: F401
