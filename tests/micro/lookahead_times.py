"""Per-call wall times of PGBART.astep with lookahead (diagnostic, not a test)."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
import pymc_bart_b200 as pb
from pymc_bart_b200.pgbart import PGBART

def run(N, p, m, chains, la, steps=60):
    rng = np.random.default_rng(0)
    X = rng.standard_normal((N, p)).astype(np.float32)
    Y = (np.sin(X[:, 0]) + 0.1 * rng.standard_normal(N)).astype(np.float32)
    rv = pb.BART("mu", X, Y, m=m)
    stp = PGBART([rv], num_particles=10, chains=chains, seed=1, lookahead=la)
    t = []
    for i in range(steps):
        if i == steps // 2:
            stp.stop_tuning()
        t0 = time.perf_counter()
        stp.astep()
        t.append((time.perf_counter() - t0) * 1e3)
    t0 = time.perf_counter(); stp.flush_history(); fl = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize()
    print(f"N={N} m={m} chains={chains} lookahead={la}: tune median {np.median(t[5:steps//2]):.2f} ms; post calls:",
          " ".join(f"{x:.2f}" for x in t[steps // 2:]), f"| flush {fl:.1f} ms")
    stp.close()

if __name__ == "__main__":      # (the history Manager starts a spawn-context server process)
    for la in (1, 16):
        run(100000, 10, 50, 4, la)
    for la in (1, 16):
        run(1000000, 50, 200, 1, la, steps=40)
