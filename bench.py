#!/usr/bin/env python
"""bench.py — PGBART draws/sec on synthetic Friedman data (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C2|C1|C3|C4|C5] [--no-c5]

One "step" = one PGBART step (astep) of every chain batched on a GPU = one draw per chain.
The headline workload is BASELINE.json configs[1] (C2: N=100k, p=10, m=50, 40 particles, 4 chains on one B200).
The same line carries a SECOND measured workload under ``config.c5``: BASELINE.json configs[4] (C5: N=1M, p=50,
m=200, 60 particles, 8 chains), the large case, so that the driver's 1/2/4/8-GPU runs report both N=1e5 and N=1e6
(north_star).  Like C2's four chains, the eight chains of C5 are batched on every GPU (weak scaling: 8 chains per GPU at
every N): one chain's control phases run while the workers stream the other chains' epochs, which a single chain per GPU
cannot do (1 -> 8 chains per GPU: 197 -> 337 draws/s on one B200).  ``config.c5.one_chain_per_gpu`` keeps the
device-timed figure of the "8 chains sharded one per GPU" reading beside it.
With --gpus N (launched by torchrun) every rank runs its own chains (weak scaling: chains are independent, no
data-path collective); the run's single collective — one NCCL all-gather of the posterior draws, as
pymc_bart_b200.sampling.gather_posterior does it — is inside the timed region.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

CONFIGS = {
    # BASELINE.json configs[0..4]:  N, p, m, P, chains per GPU, seed, likelihood (0 Normal, 1 Bernoulli-logit), output groups
    "C1": (200, 5, 10, 20, 1, 1, 0, 1),
    "C2": (100_000, 10, 50, 40, 4, 2, 0, 1),
    "C3": (50_000, 20, 100, 40, 4, 3, 1, 1),
    "C4": (50_000, 15, 50, 40, 4, 4, 0, 3),
    "C5": (1_000_000, 50, 200, 60, 8, 5, 0, 1),   # BASELINE's 8 chains, batched on every GPU (see the module docstring)
}


def friedman(N, p, seed, lik=0, groups=1):
    """SURVEY.md §8(d) synthetic inputs (Friedman; Bernoulli: centred/scaled logit; multi-output: column permutations)."""
    rng = np.random.default_rng(seed)
    X = rng.uniform(0, 1, (N, p)).astype(np.float32)

    def fr(Z):
        return 10 * np.sin(np.pi * Z[:, 0] * Z[:, 1]) + 20 * (Z[:, 2] - 0.5) ** 2 + 10 * Z[:, 3] + 5 * Z[:, 4]

    if groups > 1:
        y = np.stack([fr(X[:, np.roll(np.arange(p), j)]) + rng.normal(0, 1, N) for j in range(groups)]).astype(np.float32)
    elif lik == 1:
        pr = 1.0 / (1.0 + np.exp(-(fr(X) - 14.4) / 4.9))
        y = (rng.uniform(0, 1, N) < pr).astype(np.float32)
    else:
        y = (fr(X) + rng.normal(0, 1, N)).astype(np.float32)
    return X, y


def workload_name(name, cfg, trees_per_draw):
    """One string for both arms (ours and --impl reference), so that the driver compares like with like."""
    N, p, m, P, chains, seed, lik, groups = cfg
    return (f"{name}: Friedman N={N} p={p} m={m} trees particles={P} chains/GPU={chains} "
            f"{'Bernoulli-logit' if lik else 'Normal'} likelihood{f', {groups} outputs with separate trees' if groups > 1 else ''}, "
            f"sigma=1 fixed, batch=(0.1,0.1) -> {trees_per_draw} trees/draw, depth prior alpha(1+d)^-beta (bart.py:107-109)")


def algorithmic_bytes(N, grow_events, tree_updates, tune_updates, lik=0):
    """per grow event 14N (+9N second pass for non-Gaussian likelihoods); per tree update 23N
    (+16N while tuning)."""
    return (14.0 + (9.0 if lik else 0.0)) * N * grow_events + 23.0 * N * tree_updates + 16.0 * N * tune_updates


class ClockSampler:
    """nvidia-smi clocks / throttle reasons; started before the warm-up (nvidia-smi needs ~0.5 s to produce its first
    row), rows are stamped on arrival and only those inside [mark_begin, mark_end] count."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_ms=20):
        self.rows = []
        self.proc = None
        self.gpu = str(gpu_index)
        self.windows = []
        self.period_ms = int(period_ms)

    def start(self):
        if self.period_ms <= 0:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", str(self.period_ms),
                                          "-i", self.gpu], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, smax, reasons, n_all = [], [], set(), 0
        for t, r in self.rows:
            try:
                v_sm, v_max = float(r[1]), float(r[2])
            except Exception:
                continue
            n_all += 1
            if not any(a - 0.02 <= t <= b + 0.02 for a, b in self.windows):
                continue
            sm.append(v_sm); smax.append(v_max)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_whole_run": n_all}


# ---------------------------------------------------------------------------------------------- reference arm (CPU)
def oracle_sample(name, cfg, n_chains, steps, warm, threads):
    """`steps` draws (half tuning, half post-tuning, like our arm) of `n_chains` oracle chains on `threads` host
    threads (ctypes releases the GIL).  Returns (draws/s, seconds)."""
    from oracle.oracle_py import OracleChain
    from pymc_bart_b200.settings import make_settings

    N, p, m, P, chains, seed, lik, groups = cfg
    X, y = friedman(N, p, seed, lik, groups)
    s = make_settings(X, y, m=m, num_particles=P, seed=seed, n_chains=1, likelihood=lik, n_groups=groups)
    Xc = np.ascontiguousarray(X.T)
    orcs = [OracleChain(s, Xc, y, chain=c, group=g) for c in range(n_chains) for g in range(groups)]
    # cores left over after one thread per chain go to the chains' particle loops (independent particles; same results)
    # (small problems stay sequential: a round of N = 200 rows is shorter than starting the threads)
    inner = max(1, (os.cpu_count() or 1) // max(1, min(threads, len(orcs)))) if N >= 20000 else 1
    for o in orcs:
        o.set_threads(inner)
    oracle_sample.last_inner = inner

    def run_all(n, tune):
        if n <= 0:
            return
        sem = threading.Semaphore(threads)
        ths = []
        for o in orcs:
            def job(o=o):
                with sem:
                    for _ in range(n):
                        o.step(tune, 1.0)
            t = threading.Thread(target=job); t.start(); ths.append(t)
        for t in ths:
            t.join()

    run_all(warm, True)
    t0 = time.perf_counter()
    run_all(steps - steps // 2, True)
    run_all(steps // 2, False)
    dt = time.perf_counter() - t0
    return n_chains * steps / dt, dt, s.batch_tune


def bounded_cpu_steps(cfg, steps, n_chains, threads):
    """About 10-30 s of CPU work: one chain-draw of the oracle costs ~0.4 us x N x trees/draw x P/40 on the B200
    hosts (C2: 0.2 s, C5: 12 s; about 2.5x that in the build container)."""
    N, p, m, P = cfg[:4]
    per_draw = 400e-9 * N * max(1, int(0.1 * m)) * P / 40.0 * (2.0 if cfg[6] else 1.0)
    waves = max(1, -(-n_chains * cfg[7] // threads))
    return int(max(2, min(steps, 20.0 / (per_draw * waves))))


def run_reference(args, cfg):
    """The reference's CPU path: oracle/ port (bartrs is not installable offline), one chain per host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, p, m, P, chains, seed, lik, groups = cfg
    n_chains = chains * max(1, args.gpus)
    threads = max(1, min(os.cpu_count() or 1, n_chains * groups))
    steps, warm = args.steps, args.warmup
    if args.config != "C1":
        steps, warm = bounded_cpu_steps(cfg, steps, n_chains, threads), min(warm, 1)
    val, dt, tpd = oracle_sample(args.config, cfg, n_chains, steps, warm, threads)
    inner = oracle_sample.last_inner
    config = {"workload": workload_name(args.config, cfg, tpd),
              "note": f"CPU restatement oracle/pgbart_oracle.c (bartrs unavailable offline), {n_chains} chains x {groups} output "
                      f"groups on {threads} host threads (ctypes releases the GIL) x {inner} particle thread(s) per chain"}
    if args.config == "C2" and not args.no_c5:
        c5 = CONFIGS["C5"]
        # bounded sample: one chain per host thread, at most 16 of the workload's chains (a C5 draw takes the oracle ~10 s and a
        # chain ~350 MB of host memory)
        n5 = max(1, min(c5[4] * max(1, args.gpus), os.cpu_count() or 1, 16))
        th5 = n5
        v5, dt5, tpd5 = oracle_sample("C5", c5, n5, 2, 0, th5)
        in5 = oracle_sample.last_inner
        config["c5"] = {"workload": workload_name("C5", c5, tpd5), "value": v5, "unit": "draws/s", "ms_per_step": 1e3 * dt5 / 2,
                        "cpu_baseline": {"value": v5, "unit": "draws/s", "cores": th5 * in5, "kind": "port",
                                         "sample": f"1 tuning + 1 post-tuning draw x {n5} chains ({in5} particle thread(s) per chain), no warm-up"}}
    line = {
        "impl": "reference", "metric": "PGBART draws/sec", "value": val, "unit": "draws/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32+i64", "data": "synthetic", "config": config,
        "cpu_baseline": {"value": val, "unit": "draws/s", "cores": threads * inner, "kind": "port",
                         "sample": f"{steps - steps // 2} tuning + {steps // 2} post-tuning draws x {n_chains} chains after {warm} warm-up, "
                                   f"{threads} chain threads x {inner} particle thread(s)"},
        "e2e": {"value": val, "unit": "draws/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- our arm (B200)
def measure(name, steps, warm, args, rank, world, local, clocks, peak, peak_src, cpu_baseline=True, chains_override=0):
    """One workload on this rank's GPU: device-timed K steps (+ the all-gather of the draws when world > 1), then the
    end-to-end leg through PGBART.astep with host buffers.  Returns the dict of measurements (rank 0: complete)."""
    import torch
    import torch.distributed as dist

    from pymc_bart_b200.core import DeviceSampler
    from pymc_bart_b200.settings import make_settings

    cfg = CONFIGS[name]
    if chains_override:
        cfg = cfg[:4] + (int(chains_override),) + cfg[5:]
    N, p, m, P, chains, seed, lik, groups = cfg
    X, y = friedman(N, p, seed, lik, groups)
    s = make_settings(X, y, m=m, num_particles=P, seed=seed, n_chains=chains, chain_base=rank * chains, device=local,
                      likelihood=lik, n_groups=groups)
    dev = DeviceSampler(s, X, y)
    stream = dev.stream()
    flush = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    n_tune = steps - steps // 2
    n_post = steps - n_tune
    nvc = chains * groups   # (chain, output group) pairs: one forest and one sum-of-trees row each

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the posterior draws stay on the device: ring [draw, chain x group, N]; gathered once at the end
    # (rows keep the sampler's padding so that a draw is ONE contiguous device-to-device copy)
    # N > 1: the run's one collective, the all-gather of the posterior draws (sampling.gather_posterior), is FUSED into the
    # step kernel: the buffer is a symmetric allocation mapped into every peer, the commit sweep that keeps a draw stores it
    # into block `rank` of all of them over NVLink (bk_set_draw_peers); what remains after the last launch is a barrier.
    # --nccl-gather (or symmetric memory unavailable): one NCCL all_gather_into_tensor after the steps instead.
    n_keep = max(n_post, 1)
    gathered = peer_hdl = None
    gather_kind = "none (1 GPU)"
    if world > 1:
        from pymc_bart_b200.sampling import peer_draw_buffer

        why = "--nccl-gather"
        if not args.nccl_gather:
            gathered, peer_hdl, peer_ptrs, why = peer_draw_buffer(n_keep, (nvc, dev.ld), rank, world)
        if gathered is not None:
            draws_dev = gathered[rank * n_keep:(rank + 1) * n_keep]
            dev.set_draw_peers(peer_ptrs, gathered.data_ptr())
            gather_kind = "P2P stores from the step kernel's commit sweep into every peer's buffer (NVLink), then a barrier"
        else:
            gather_kind = f"NCCL all_gather_into_tensor after the steps ({why})"
    if gathered is None:
        draws_dev = torch.empty((n_keep, nvc, dev.ld), dtype=torch.float32, device="cuda")
        if world > 1:
            gathered = torch.empty((world * n_keep, nvc, dev.ld), dtype=torch.float32, device="cuda")
    for i in range(warm):
        dev.step(True, 1.0)
    if world > 1:   # warm the collective of the timed region (channel / NVLS / signal-pad setup is not the run's cost)
        with torch.cuda.stream(stream):
            if peer_hdl is not None:
                peer_hdl.barrier(channel=0)
            else:
                dist.all_gather_into_tensor(gathered, draws_dev)
    sync_all()
    # K steps in launches of `spl` steps (bk_run_launch: the chains of a launch do not wait for each other between steps and
    # no launch gap separates them; sigma is fixed in this benchmark).  A launch never mixes tuning and
    # posterior steps; the posterior steps' draws are written into the ring by each step's last commit.
    spl = max(1, min(int(args.steps_per_launch), dev.MAX_STEPS_PER_LAUNCH))
    plan = []
    i = 0
    while i < steps:
        tune = i < n_tune
        n = min(spl, (n_tune if tune else steps) - i)
        plan.append((i, n, tune))
        i += n
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in plan]
    grow = tupd = tune_upd = rounds = phases = 0
    us_total = us_control = us_data = 0
    us_by_phase = {True: 0.0, False: 0.0}
    t_w0 = time.perf_counter()
    for li, (i0, n, tune) in enumerate(plan):
        if flush is not None:
            with torch.cuda.stream(stream):
                flush.fill_(li & 0xFF)   # untimed: evicts the working set from the 126 MB L2
        ev[li][0].record(stream)
        dev.run_launch(n, tune, 1.0, draws_out=None if tune else draws_dev[i0 - n_tune: i0 - n_tune + n])
        ev[li][1].record(stream)
        _, sts = dev.run_wait()
        for st in sts:
            for c in range(nvc):
                grow += st[c].grow_events; tupd += st[c].tree_updates; rounds += st[c].rounds
                tune_upd += st[c].tree_updates if tune else 0
            phases += st[0].phases
        last = sts[-1]           # (the launch-wide in-kernel timers travel with the last step's record)
        us_total += max(last[c].us_total for c in range(nvc)); us_control += last[0].us_control; us_data += last[0].us_data
        us_by_phase[tune] += max(last[c].us_total for c in range(nvc))
    t_gather0 = torch.cuda.Event(enable_timing=True); t_gather1 = torch.cuda.Event(enable_timing=True)
    t_gather0.record(stream)
    if world > 1:   # the run's single collective: ordered after the steps on their stream
        with torch.cuda.stream(stream):
            if peer_hdl is not None:
                peer_hdl.barrier(channel=0)      # every rank's launches are complete: every block of every buffer is there
            else:
                dist.all_gather_into_tensor(gathered, draws_dev)
    t_gather1.record(stream)
    torch.cuda.synchronize()
    gather_check = None
    if world > 1 and peer_hdl is not None:   # untimed: the fused gather against NCCL's
        ref = torch.empty_like(gathered)
        dist.all_gather_into_tensor(ref, draws_dev.contiguous())
        torch.cuda.synchronize()
        same = torch.tensor([1 if torch.equal(ref, gathered) else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        gather_check = "equal to nccl all_gather_into_tensor on every rank" if int(same.item()) else "MISMATCH against nccl all_gather_into_tensor"
        del ref
    clocks.window(t_w0, time.perf_counter())
    launch_ms = [a.elapsed_time(b) for a, b in ev]
    step_ms = [ms / n for ms, (_, n, _) in zip(launch_ms, plan) for _ in range(n)]      # per step, for the tuning / posterior split
    gather_ms = t_gather0.elapsed_time(t_gather1) if world > 1 else 0.0
    total_ms = float(sum(launch_ms)) + gather_ms
    tm = torch.tensor([total_ms, gather_ms], dtype=torch.float64, device="cuda")
    agg = torch.tensor([float(grow), float(tupd), float(tune_upd)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    total_ms_max, gather_ms_max = [float(v) for v in tm.tolist()]
    g_all, t_all, tu_all = [float(v) for v in agg.tolist()]
    value = world * chains * steps / (total_ms_max / 1e3)
    out = {"value": value, "unit": "draws/s", "ms_per_step": total_ms_max / steps, "steps": steps, "warmup": warm,
           "steps_per_launch": spl, "launches": len(plan), "gather": gather_kind, "gather_check": gather_check,
           "in_kernel_us": {"step": us_total / steps, "control_chain0": us_control / steps, "data_wait_chain0": us_data / steps,
                            "step_tuning": us_by_phase[True] / max(n_tune, 1), "step_post": us_by_phase[False] / max(n_post, 1),
                            "event_tuning": 1e3 * float(np.mean(step_ms[:n_tune])) if n_tune else None,
                            "event_post": 1e3 * float(np.mean(step_ms[n_tune:])) if n_post else None}}
    if args.profile_only:
        dev.close()
        return out

    # ---------------- end to end through the public step API (host buffers, copies inside the timed region)
    from pymc_bart_b200 import BART
    from pymc_bart_b200.pgbart import PGBART

    dev.close()
    del dev, draws_dev, gathered
    torch.cuda.synchronize()
    e2e_steps = max(10, steps // 4) if args.lookahead <= 1 else max(64, steps)
    t0 = time.perf_counter()
    rv = BART("mu", X, y, m=m, shape=(groups, N) if groups > 1 else None, separate_trees=groups > 1)
    # sigma is fixed in this benchmark, so the posterior draws may be served from launches of several steps
    # (PGBART(lookahead=n): the next launch runs while the caller consumes the previous one); every astep() still returns the
    # draw as a host array with its stats and publishes its batch of rewritten trees.  `e2e.one_launch_per_call` below is the
    # same loop with lookahead=1 (what a model whose scale is updated by another step method between draws gets).
    step_kw = dict(num_particles=P, chains=chains, chain_base=rank * chains, seed=seed, device=local, store_history=True,
                   likelihood="bernoulli" if lik else "normal", lookahead=args.lookahead)
    td = e2e_steps // 2 if (args.lookahead > 1 and args.tune_ahead) else None
    stp = PGBART([rv], tune_draws=td, **step_kw).prepare()      # device state, X/Y upload, pinned buffers
    t_build = time.perf_counter() - t0
    # warm-up on a second step object (first launch of the process, allocator, writer thread), so that the timed region
    # starts with nothing computed ahead: every one of its K steps runs inside it
    warm_rv = BART("mu_warm", X, y, m=m, shape=(groups, N) if groups > 1 else None, separate_trees=groups > 1)
    warm_stp = PGBART([warm_rv], tune_draws=2 if td is not None else None, **step_kw)
    for i in range(4):
        if i == 2:
            warm_stp.stop_tuning()
        warm_stp.astep()
    warm_stp.close()
    del warm_stp, warm_rv
    sync_all()
    t_w0 = time.perf_counter()
    call_ms = []
    for i in range(e2e_steps):
        if i == e2e_steps // 2:
            stp.stop_tuning()
        tc = time.perf_counter()
        val_host, stats = stp.astep()      # step kernel, D2H sum-of-trees + VI counts + stats (+ the rewritten trees after tuning)
        call_ms.append((time.perf_counter() - tc) * 1e3)
    stp.flush_history()                    # every batch has reached op.all_trees (the Manager proxy of pymc_bart/bart.py:134)
    e2e_dt = time.perf_counter() - t_w0    # all K draws, stats and tree batches are on the host (a launch that ran ahead is not waited for)
    torch.cuda.synchronize()
    clocks.window(t_w0, time.perf_counter())
    e2e_t = torch.tensor([e2e_dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_val = world * chains * e2e_steps / float(e2e_t.item())
    h2d = nvc * 4
    d2h = nvc * N * 4 + nvc * p * 4 + nvc * 64 + 4 + int(getattr(stp, "history_bytes_per_step", 0))
    h2d_once = stp.core.h2d_bytes
    post_ms = call_ms[e2e_steps // 2:]
    call_summary = {"tuning_median": float(np.median(call_ms[: e2e_steps // 2])), "post_first": float(post_ms[0]),
                    "post_median": float(np.median(post_ms)), "post_max_after_first": float(max(post_ms[1:])),
                    "post_sum": float(sum(post_ms)), "note": "wall ms per astep() call on this rank"}
    e2e_single = None
    if args.lookahead > 1:       # the same loop, one launch per astep() call
        stp.close()
        rv1 = BART("mu1", X, y, m=m, shape=(groups, N) if groups > 1 else None, separate_trees=groups > 1)
        stp1 = PGBART([rv1], num_particles=P, chains=chains, chain_base=rank * chains, seed=seed, device=local, store_history=True,
                      likelihood="bernoulli" if lik else "normal", lookahead=1)
        n1 = max(10, e2e_steps // 2)
        for i in range(3):
            stp1.astep()
        sync_all()
        t1 = time.perf_counter()
        for i in range(n1):
            if i == n1 // 2:
                stp1.stop_tuning()
            stp1.astep()
        stp1.flush_history()
        torch.cuda.synchronize()
        dt1 = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt1, op=dist.ReduceOp.MAX)
        e2e_single = {"value": world * chains * n1 / float(dt1.item()), "unit": "draws/s", "steps": n1}
        stp1.close()
        del stp1
    # ---------------- posterior prediction from the history the e2e leg just published (row N1): all chains, one launch
    predict = None
    try:
        from pymc_bart_b200.utils import _get_posterior_sampler

        sampler = _get_posterior_sampler(rv.owner.op)
        n_d = min(sampler.n_draws, 64)
        picks = np.arange(n_d) * (sampler.n_draws // n_d)
        Xd = sampler.upload(X)
        sampler._forests.predict(Xd, picks)                       # warm-up (module load, first launch)
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        p0.record()
        for _ in range(reps):
            pred = sampler._forests.predict(Xd, picks)
        p1.record()
        torch.cuda.synchronize()
        ms = p0.elapsed_time(p1) / reps
        predict = {"value": N * m * n_d * sampler.n_outputs / (ms / 1e3), "unit": "row-tree walks/s", "ms_per_call": ms,
                   "rows": N, "trees": m, "draws": int(n_d), "outputs": int(sampler.n_outputs), "chains_in_store": int(sampler.n_chains),
                   "history_bytes_on_device": int(sampler._forests.history_bytes),
                   "note": "one bk_predict_history launch over all chains' draws, X resident, output [draws][outputs][rows] float32 on the device"}
        del pred, Xd, sampler
    except Exception as e:  # noqa
        predict = {"value": None, "note": f"failed: {e}"}
    stp.close()
    del stp
    if rank != 0:
        return out

    # per LAUNCH of pgbart_step_kernel (a launch = `spl` steps of every chain): mean launch duration from the CUDA events on
    # the launch stream, algorithmic bytes of the steps it ran
    steps_per_launch_mean = steps / len(plan)
    ms_kernel = float(np.mean(launch_ms))
    bytes_per_launch = algorithmic_bytes(N, grow / steps, tupd / steps, tune_upd / steps, lik) * steps_per_launch_mean
    achieved = bytes_per_launch / (ms_kernel / 1e3) / 1e9
    traffic = traffic_src = None
    try:   # DRAM bytes per launch of the same command under `ncu --set full` (profiles/, committed)
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(name, {})
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")
        if int(tj.get("chains_per_gpu", chains)) != chains:   # captured with another number of chains per GPU: not this workload's traffic
            traffic, traffic_src = None, f"no capture at {chains} chains per GPU ({traffic_src})"
        cap_spl = float(tj.get("steps_per_launch", 1))
        if traffic is not None and abs(cap_spl - steps_per_launch_mean) > 1e-9:   # captured at another launch size: per step x steps
            traffic = traffic / cap_spl * steps_per_launch_mean
            traffic_src = f"{traffic_src}; captured at {cap_spl:g} steps per launch, scaled to {steps_per_launch_mean:g}"
    except Exception:
        pass
    out.update({
        "workload": workload_name(name, cfg, s.batch_tune),
        "draws_timed": f"{n_tune} tuning + {n_post} post-tuning",
        "l2": "256 MB fill between timed steps (L2 flushed)" if flush is not None else "no flush (working set stays in L2)",
        "grow_events_per_tree_update": g_all / max(t_all, 1.0),
        "tree_updates_per_s": t_all / (total_ms_max / 1e3),
        "grow_events_per_s": g_all / (total_ms_max / 1e3),
        "rounds_per_tree_update": rounds / max(tupd, 1),
        "grid_phases_per_step": phases / steps,
        "gather_ms": gather_ms_max,
        "gather_bytes_per_rank": int(n_post * nvc * N * 4) if world > 1 else 0,   # (+0.1 % row padding on the wire)
        "e2e": {"value": e2e_val, "unit": "draws/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "call_ms": call_summary, "h2d_bytes_once_XY": h2d_once, "build_seconds": t_build, "lookahead": int(args.lookahead), "tune_draws_given": bool(args.lookahead > 1 and args.tune_ahead),
                "one_launch_per_call": e2e_single,
                "history": "store_history=True: after tuning every step's rewritten trees are exported (op.all_trees protocol)"},
        "predict": predict,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "frac_dram_traffic": (traffic / (ms_kernel / 1e3) / 1e9 / peak) if traffic else None,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                     "kernel": "pgbart_step_kernel", "kernel_ms": ms_kernel, "steps_per_launch": steps_per_launch_mean,
                     "note": "achieved / frac use SURVEY 8(d)'s ALGORITHMIC bytes (every grow event is charged the column, both "
                             "leaf-id rows and the two residual tiles: 14 N); the kernel loads the residual tiles once per row tile "
                             "and finds much of the rest in L1/L2, so frac can exceed 1 when several chains overlap — "
                             "frac_dram_traffic (ncu DRAM bytes per launch / launch time / peak) is the hardware figure"},
    })
    if cpu_baseline:
        # ---------------- CPU baseline: the oracle on the host cores, bounded sample of the same workload
        try:
            ncpu = max(1, min((os.cpu_count() or 1) // groups, chains))   # chains sampled; each needs `groups` threads
            nd = args.cpu_draws or bounded_cpu_steps(cfg, 40, ncpu, ncpu * groups)
            v, cdt, _ = oracle_sample(name, cfg, ncpu, nd, 0, ncpu * groups)
            inner = oracle_sample.last_inner
            out["cpu_baseline"] = {"value": v, "unit": "draws/s", "cores": ncpu * groups * inner, "kind": "port",
                                   "sample": f"{nd - nd // 2} tuning + {nd // 2} post-tuning draws x {ncpu} chains of {name} in {cdt:.1f} s, "
                                             f"one host thread per chain x {inner} particle thread(s) each (oracle/pgbart_oracle.c, gcc -O3; "
                                             f"bartrs is not installable offline)"}
        except Exception as e:  # noqa
            out["cpu_baseline"] = {"value": None, "unit": "draws/s", "cores": 0, "kind": "port", "sample": f"failed: {e}"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="C2")
    ap.add_argument("--chains", type=int, default=0, help="override the chains per GPU of --config (experiments; 0 = the config's own)")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--no-c5", action="store_true", help="skip the second workload (config.c5)")
    ap.add_argument("--c5-steps", type=int, default=30)
    ap.add_argument("--cpu-draws", type=int, default=None, help="CPU baseline sample size (draws per chain)")
    ap.add_argument("--profile-only", action="store_true", help="device-timed leg only (for ncu runs)")
    ap.add_argument("--clock-ms", type=int, default=20, help="nvidia-smi sampling period in ms (0 = no clock sampling)")
    ap.add_argument("--tune-ahead", type=int, default=1, help="e2e leg: tell the step where tuning ends (PGBART(tune_draws=)), so tuning "
                    "steps are served ahead too (0 = the PyMC protocol: one launch per tuning call)")
    ap.add_argument("--nccl-gather", action="store_true", help="N > 1: gather the draws with one NCCL all_gather after the steps "
                    "instead of the P2P stores fused into the step kernel")
    ap.add_argument("--lookahead", type=int, default=16, help="PGBART(lookahead=) of the e2e leg (1 = one launch per astep call)")
    ap.add_argument("--steps-per-launch", type=int, default=16, help="steps of every chain per kernel launch in the device-timed leg (1..16)")
    args = ap.parse_args()
    if args.chains > 0:
        c = list(CONFIGS[args.config]); c[4] = int(args.chains); CONFIGS[args.config] = tuple(c)
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
        return

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    steps, warm = args.steps, max(3, args.warmup)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    clocks = ClockSampler(local, args.clock_ms)
    clocks.start()
    main_m = measure(args.config, steps, warm, args, rank, world, local, clocks, peak, peak_src)
    c5_m = None
    if args.config == "C2" and not args.no_c5 and not args.profile_only:
        c5_m = measure("C5", max(25, args.c5_steps), 5, args, rank, world, local, clocks, peak, peak_src, cpu_baseline=False)
        # the "8 chains sharded one per GPU" reading of the same config, device-timed only
        one_args = argparse.Namespace(**vars(args)); one_args.profile_only = True
        c5_one = measure("C5", max(25, args.c5_steps), 5, one_args, rank, world, local, clocks, peak, peak_src, cpu_baseline=False,
                         chains_override=1)
        if rank == 0:
            c5_m["one_chain_per_gpu"] = {k: c5_one[k] for k in ("value", "unit", "ms_per_step", "steps", "steps_per_launch", "in_kernel_us")}
    clk = clocks.stop()
    if rank == 0:
        if args.profile_only:
            print(json.dumps({"profile_only": True, "value": main_m["value"], "ms_per_step": main_m["ms_per_step"],
                              "steps_per_launch": main_m["steps_per_launch"], "in_kernel_us": main_m["in_kernel_us"]}), flush=True)
        else:
            config = {k: main_m[k] for k in ("workload", "draws_timed", "l2", "grow_events_per_tree_update", "tree_updates_per_s",
                                             "grow_events_per_s", "rounds_per_tree_update", "grid_phases_per_step", "in_kernel_us",
                                             "steps_per_launch", "launches", "gather", "gather_check", "gather_ms",
                                             "gather_bytes_per_rank", "predict")}
            if c5_m is not None:
                c5_m["n_gpus"] = world
                c5_m["gpu_launches"] = c5_m["launches"]
                config["c5"] = c5_m
            line = {
                "metric": "PGBART draws/sec", "value": main_m["value"], "unit": "draws/s", "n_gpus": world, "steps": steps, "warmup": warm,
                "ms_per_step": main_m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32+i64", "data": "synthetic", "config": config,
                "gpu_launches": main_m["launches"], "e2e": main_m["e2e"], "roofline": main_m["roofline"], "clocks": clk,
                "cpu_baseline": main_m.get("cpu_baseline"),
            }
            print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
