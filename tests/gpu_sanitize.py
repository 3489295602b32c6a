"""Small steps of every path (Gaussian, Bernoulli, multi-output, several chains) for compute-sanitizer:
   compute-sanitizer --tool memcheck python tests/gpu_sanitize.py
   compute-sanitizer --tool synccheck python tests/gpu_sanitize.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import friedman
from pymc_bart_b200.core import DeviceSampler
from pymc_bart_b200.settings import make_settings

for (N, p, m, P, chains, lik, groups) in [(777, 6, 6, 9, 2, 0, 1), (500, 5, 5, 12, 1, 1, 1), (300, 4, 4, 6, 2, 0, 3), (3000, 8, 10, 40, 5, 0, 1)]:
    X, y, f = friedman(N, p, 3, kind="bernoulli" if lik else "normal")
    Y = np.stack([y, -y, 0.5 * y]) if groups > 1 else y
    s = make_settings(X, Y, m=m, num_particles=P, seed=1, n_chains=chains, likelihood=lik, n_groups=groups)
    dev = DeviceSampler(s, X, Y)
    for d in range(4):
        vi, st = dev.step(d < 2, 1.0)
        assert all(st[c].error_flags == 0 for c in range(chains * groups))
    dev.close()
    print("ok", N, p, m, P, chains, lik, groups, flush=True)
