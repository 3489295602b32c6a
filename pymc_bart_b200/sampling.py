"""Minimal PyMC-free driver: tune + draws through PGBART.astep, chains sharded one group per
GPU (torch.distributed / NCCL), ONE all-gather of the posterior draws at the end.

Stands in for ``pm.sample(tune, draws, chains, step=[PGBART(...)])`` (tests/test_bart.py:58,235)
when PyMC is absent; the likelihood scale is held fixed (or supplied per draw by ``sigma_fn``).
"""
from __future__ import annotations

import numpy as np


def chain_base_for_rank(rank: int, chains_per_rank: int) -> int:
    """Global index of a rank's first chain: chain c of rank r draws from Philox key (seed, r*C + c),
    so a chain's stream does not depend on which GPU runs it."""
    return int(rank) * int(chains_per_rank)


def gather_posterior(local, world: int):
    """The single collective of a run: all-gather of per-rank draws [C, ...] -> [world*C, ...].
    `local` is a torch tensor on the backend's device (CUDA for NCCL, CPU for gloo)."""
    import torch
    import torch.distributed as dist

    if world <= 1:
        return local
    local = local.contiguous()
    # one preallocated buffer (rank-major concatenation along dim 0), no list copy-outs
    out = torch.empty((world * local.shape[0], *local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local)
    return out


def peer_draw_buffer(rows_per_rank: int, row_shape, rank: int, world: int):
    """Buffer for the all-gather of device-resident draws FUSED into the step kernel (bk_set_draw_peers): a symmetric
    allocation [world * rows_per_rank, *row_shape] float32 on every rank, mapped into every peer (torch symmetric memory:
    CUDA VMM handles exchanged through the process group's store).  Rank r points bk_run_launch's draws_out into its block
    r; the commit sweep stores each draw there and into the same place of the 7 or fewer peers over NVLink, so after the last
    launch a barrier is all that is left of the collective.
    Returns (buffer, handle, peer_ptrs, why): buffer None (and `why` set) when symmetric memory is unavailable — every
    rank takes the same branch (the outcome is agreed with an all-reduce); the caller then uses gather_posterior."""
    import torch
    import torch.distributed as dist

    buf = hdl = None
    ptrs, why = [], ""
    try:
        if world - 1 > 7:
            raise RuntimeError("more than 7 peers")
        import torch.distributed._symmetric_memory as symm

        dev = torch.device("cuda", torch.cuda.current_device())
        buf = symm.empty((world * int(rows_per_rank), *row_shape), dtype=torch.float32, device=dev)
        hdl = symm.rendezvous(buf, dist.group.WORLD)
        ptrs = [int(hdl.buffer_ptrs[r]) for r in range(world) if r != rank]
    except Exception as e:  # noqa: BLE001 - any failure means "use NCCL"
        why = f"{type(e).__name__}: {e}"[:200]
        buf = hdl = None
    ok = torch.tensor([1 if buf is not None else 0], dtype=torch.int32, device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 0:
        return None, None, [], why or "symmetric memory failed on another rank"
    return buf, hdl, ptrs, ""


def sample(rv, tune=200, draws=200, chains=4, num_particles=10, batch=(0.1, 0.1), sigma=1.0, seed=0,
           likelihood="normal", sigma_fn=None, keep_draws=True, **step_kwargs):
    """Returns a dict: posterior (chains_total, draws, N) float32 [if keep_draws],
    variable_inclusion (chains_total, draws) of base64 strings, and the step object."""
    import torch
    import torch.distributed as dist

    from .pgbart import PGBART

    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if distributed else 0
    world = dist.get_world_size() if distributed else 1
    device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if sigma_fn is None:          # fixed likelihood parameters: draws may be served from launches of several steps,
        step_kwargs.setdefault("lookahead", 16)      # tuning included, because this driver knows where tuning ends
        step_kwargs.setdefault("tune_draws", tune)
    step = PGBART([rv], num_particles=num_particles, batch=batch, likelihood=likelihood, sigma=sigma, chains=chains,
                  chain_base=chain_base_for_rank(rank, chains), seed=seed, device=device, **step_kwargs)
    N = step.n_rows
    k_out = step.groups * step.outputs
    vshape = (N,) if k_out == 1 else (k_out, N)
    post = np.empty((chains, draws, *vshape), dtype=np.float32) if keep_draws else None
    vi = [[None] * draws for _ in range(chains)]
    for d in range(tune + draws):
        if d == tune:
            step.stop_tuning()
        if sigma_fn is not None:
            step.sigma = sigma_fn(d, step)
        value, stats = step.astep()      # (after tuning every draw appends its trees to the chains' op.all_trees entries)
        if d >= tune:
            v = value if chains > 1 else value[None]
            s = stats if chains > 1 else [stats[0]]
            if keep_draws:
                post[:, d - tune] = v
            for c in range(chains):
                vi[c][d - tune] = s[c]["variable_inclusion"]
    step.flush_history()
    out = {"step": step, "variable_inclusion": vi, "posterior": post, "rank": rank, "world": world}
    if distributed and world > 1 and keep_draws:
        # the single collective of the run: all-gather of the posterior draws over NVLink
        out["posterior"] = gather_posterior(torch.from_numpy(post).cuda(), world).cpu().numpy()
    return out


def sample_joint(rvs, observed, tune=200, draws=200, num_particles=10, batch=(0.1, 0.1), sigma=1.0, seed=0, sigma_fn=None,
                 **step_kwargs):
    """Several BART variables in ONE Normal likelihood, ``observed ~ Normal(sum of the variables, sigma)`` — the model of
    tests/test_bart.py:167-241 (``pm.Normal("y", mu1 + mu2, sigma, observed=Y)``, one PGBART step per variable) without PyMC.
    One chain; the steps run one after the other on the same point, as pm.sample's compound step does: each sees the data
    minus the other variables' current values (``PGBART(observed=, offset_names=)``).
    Returns {"posterior": {name: (draws, N) float32}, "variable_inclusion": {name: [str]}, "steps": {name: PGBART}}."""
    from .pgbart import PGBART

    names = [rv.name for rv in rvs]
    if len(set(names)) != len(names):
        raise ValueError("the BART variables need distinct names")
    observed = np.asarray(observed, dtype=np.float64)
    steps, point = {}, {}
    for k, rv in enumerate(rvs):
        steps[rv.name] = PGBART([rv], num_particles=num_particles, batch=batch, likelihood="normal", sigma=sigma, seed=seed + k,
                                observed=observed, offset_names=[n for n in names if n != rv.name], **step_kwargs)
        op = rv.owner.op
        point[rv.name] = np.full(observed.shape, float(np.asarray(op.Y, dtype=np.float64).mean()))    # bart.py:148 initial value
    post = {n: np.empty((draws, *observed.shape), dtype=np.float32) for n in names}
    vi = {n: [None] * draws for n in names}
    for d in range(tune + draws):
        for n in names:
            st = steps[n]
            if d == tune:
                st.stop_tuning()
            if sigma_fn is not None:
                st.sigma = sigma_fn(d, point)
            point, stats = st.step(point)
            if d >= tune:
                post[n][d - tune] = point[n]
                vi[n][d - tune] = stats[0]["variable_inclusion"]
    for st in steps.values():
        st.flush_history()
    return {"posterior": post, "variable_inclusion": vi, "steps": steps}
