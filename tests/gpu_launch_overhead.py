"""Where the time between the CUDA events of a bench step goes that the kernel's own timers do not see
(run on the GPU box: python tests/gpu_launch_overhead.py [C2])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from bench import CONFIGS, friedman
from pymc_bart_b200.core import DeviceSampler
from pymc_bart_b200.settings import make_settings

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
N, p, m, P, chains, seed, lik, groups = CONFIGS[cfg]
X, y = friedman(N, p, seed, lik, groups)
s = make_settings(X, y, m=m, num_particles=P, seed=seed, n_chains=chains, likelihood=lik, n_groups=groups)
dev = DeviceSampler(s, X, y)
stream = dev.stream()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for i in range(10):
    dev.step(True, 1.0)


def run(label, steps=60, use_flush=True, pre_op=None):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    us = 0
    for i in range(steps):
        with torch.cuda.stream(stream):
            if use_flush:
                flush.fill_(i & 0xFF)
            elif pre_op is not None:
                pre_op()
        ev[i][0].record(stream)
        dev.step_launch(i < steps // 2, 1.0)
        ev[i][1].record(stream)
        if os.environ.get("BK_EXPERIMENT_NO_D2H"):
            stream.synchronize()
            st = None
        else:
            _, st = dev.step_wait()
            us += max(st[c].us_total for c in range(chains * groups))
    torch.cuda.synchronize()
    ms = np.mean([a.elapsed_time(b) for a, b in ev])
    print(f"{label}: event {1e3 * ms:.1f} us/step, in-kernel (slowest chain) {us / steps:.1f} us/step", flush=True)


small = torch.empty(64 * 1024 * 1024, dtype=torch.uint8, device="cuda")
run("flush 256 MB before the step")
run("no flush, stream idle at launch", use_flush=False)
run("no flush, 64 MB fill queued before the step", use_flush=False, pre_op=lambda: small.fill_(1))
os.environ["BK_EXPERIMENT_NO_D2H"] = "1"
run("flush, no D2H of the outputs")
