"""Prints the control/data/sync split of the step kernel (control CTA's view) for a bench config."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from bench import CONFIGS, friedman
from pymc_bart_b200.core import DeviceSampler
from pymc_bart_b200.settings import make_settings
cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
N, p, m, P, chains, seed, lik, groups = CONFIGS[cfg]
if len(sys.argv) > 3: chains = int(sys.argv[3])
X, y = friedman(N, p, seed, lik, groups)
s = make_settings(X, y, m=m, num_particles=P, seed=seed, n_chains=chains, likelihood=lik, n_groups=groups)
dev = DeviceSampler(s, X, y)
for i in range(5): dev.step(True, 1.0)
import ctypes
acc = np.zeros(9)
sub = np.zeros(8)
for i in range(steps):
    _, st = dev.step(i < steps // 2, 1.0)
    buf = (ctypes.c_ulonglong * 8)()
    dev.lib.bk_debug_timers.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
    dev.lib.bk_debug_timers(dev.h, 0, buf)
    sub += np.array(list(buf), dtype=np.float64) / 1e3
    acc += [st[0].us_control, st[0].us_data, st[0].us_sync, st[0].us_total, st[0].phases, st[0].rounds, st[0].grow_events, st[0].count_passes, st[0].reserved[0]]
acc /= steps
print(f"{cfg} chains={chains}: per step us control={acc[0]:.0f} data={acc[1]:.0f} sync={acc[2]:.0f} total={acc[3]:.0f} phases={acc[4]:.1f} "
      f"rounds={acc[5]:.1f} grow={acc[6]:.1f} count_passes={acc[7]:.1f} sweep_wait={acc[8]:.0f}us; per phase us control={acc[0]/acc[4]:.1f} data={acc[1]/acc[4]:.1f} sync={acc[2]/acc[4]:.1f}")

sub /= steps
print("control sub-steps us/step: finalize=%.0f weights=%.0f resample=%.0f copy=%.0f pop=%.0f select=%.0f jobs=%.0f finish+init=%.0f" % tuple(sub))

wd = (ctypes.c_ulonglong * (148 * 16 + 32))()
dev.lib.bk_debug_worker_timers.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
if dev.lib.bk_debug_worker_timers(dev.h, wd, 148) == 0:
    cd = np.array(list(wd)[148 * 16:], dtype=np.float64)
    w = np.array(list(wd)[:148 * 16], dtype=np.float64).reshape(148, 16)
    w = w[w[:, 4] > 0]
    n = w[:, 4].sum()
    print("worker ROUND claims: %d per step; mean ns after publish: claimed=%.0f staged=%.0f units_done=%.0f done_added=%.0f; "
          "slowest CTA mean done_added=%.0f" % (n / (steps + 5), w[:, 0].sum() / n, w[:, 1].sum() / n, w[:, 2].sum() / n, w[:, 3].sum() / n,
                                             (w[:, 3] / w[:, 4]).max()))
    print("control cycle split per step (chain 0, thread 0):", " ".join("%d:%.0f" % (i, v / (steps + 5)) for i, v in enumerate(cd) if v > 0))
