"""Manual GPU bring-up script (not collected by pytest): prints progress with timestamps."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
t0 = time.time()


def log(*a):
    print(f"[{time.time() - t0:7.2f}s]", *a, flush=True)


import numpy as np  # noqa: E402

log("importing torch")
import torch  # noqa: E402

log("torch", torch.__version__, torch.cuda.is_available())
from helpers import friedman  # noqa: E402
from oracle.oracle_py import OracleChain  # noqa: E402
from pymc_bart_b200.core import DeviceSampler  # noqa: E402
from pymc_bart_b200.settings import make_settings  # noqa: E402

N, p, m, P, draws = [int(a) for a in sys.argv[1:6]]
chains = int(sys.argv[6]) if len(sys.argv) > 6 else 1
X, y, _ = friedman(N, p, 1)
s = make_settings(X, y, m=m, num_particles=P, seed=1, n_chains=chains, trace_capacity=20000)
log("creating device sampler")
dev = DeviceSampler(s, X, y)
log("created; workspace MB", dev.workspace_bytes / 1e6)
orc = [OracleChain(s, X.T.copy(), y, chain=c) for c in range(chains)]
for d in range(draws):
    tune = d < draws // 2
    t1 = time.time()
    import threading, ctypes
    res = {}
    def _run():
        try:
            res["out"] = dev.step(tune, 1.0)
        except Exception as e:  # noqa
            res["err"] = e
    th = threading.Thread(target=_run, daemon=True)
    th.start()
    th.join(float(os.environ.get("STEP_TIMEOUT", "20")))
    if th.is_alive():
        log("STEP HUNG; markers:")
        lib = dev.lib
        lib.bk_debug_markers.restype = ctypes.POINTER(ctypes.c_int32)
        lib.bk_debug_markers.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
        cnt = ctypes.c_int()
        ptr = lib.bk_debug_markers(dev.h, ctypes.byref(cnt))
        if ptr:
            full = np.ctypeslib.as_array(ptr, shape=(cnt.value,)).copy()
            bar = full[4736:4736 + 148 * 64].reshape(148, 32, 2)
            log("block0 lane0 (line,nbar):", [(int(v) >> 16, int(v) & 0xffff) for v in bar[0, :, 0]])
            log("block0 lane1 (line,nbar):", [(int(v) >> 16, int(v) & 0xffff) for v in bar[0, :, 1]])
            log("block1 lane0 (line,nbar):", [(int(v) >> 16, int(v) & 0xffff) for v in bar[1, :4, 0]])
            log("abort/barrier words n/a")
            arr = full[:148 * 33].reshape(-1, 33)
            from collections import Counter
            dec = lambda v: (int(v) & 0xfff, int(v) >> 12)
            log("block0 warps (code,phase):", [dec(v) for v in arr[0, :32]])
            log("other blocks (code,phase):", Counter(dec(v) for v in arr[1:, :32].ravel()))
        os._exit(3)
    if "err" in res:
        raise res["err"]
    vi, stats = res["out"]
    dt = time.time() - t1
    a = dev.sum_trees().cpu().numpy()
    ok = True
    for c in range(chains):
        vio, sto = orc[c].step(tune, 1.0)
        b = orc[c].sum_trees()
        ta, tb = dev.trace(c), orc[c].trace()
        same_st = np.array_equal(a[c].view(np.uint32), b.view(np.uint32))
        same_tr = len(ta) == len(tb) and np.array_equal(ta.view(np.uint8), tb.view(np.uint8))
        if d < 3 or not (same_st and same_tr):
            log(f"draw {d} chain {c} step {dt*1e3:.2f} ms phases {stats[c].phases} rounds {stats[c].rounds} grow {stats[c].grow_events}/{sto.grow_events} "
                f"err {stats[c].error_flags} st_equal {same_st} trace_equal {same_tr} len {len(ta)}/{len(tb)} maxdiff {np.abs(a[c]-b).max():.3g}")
        if not (same_st and same_tr):
            ok = False
            k = min(len(ta), len(tb))
            for i in range(k):
                if ta[i].tobytes() != tb[i].tobytes():
                    log("first diff rec", i, "\n gpu", ta[i], "\n orc", tb[i])
                    break
    if not ok:
        sys.exit(1)
log("all draws identical; last step ms", dt * 1e3)
