"""world_size-2 gloo test of the multi-GPU host logic: chains are sharded by rank through the
Philox key word (chain_base), no data-path collective, ONE all-gather of the draws at the end;
a chain's draws do not depend on which rank ran it."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import friedman
    from oracle.oracle_py import OracleChain
    from pymc_bart_b200.sampling import chain_base_for_rank, gather_posterior
    from pymc_bart_b200.settings import make_settings

    chains = 2
    X, y, _ = friedman(120, 5, 4)
    base = chain_base_for_rank(rank, chains)
    s = make_settings(X, y, m=6, num_particles=6, seed=4, chain_base=base)
    local = []
    for c in range(chains):
        o = OracleChain(s, X.T.copy(), y, chain=c)       # stands in for the GPU chains of this rank
        draws = []
        for d in range(12):
            o.step(d < 6, 1.0)
            draws.append(o.sum_trees().copy())
        local.append(np.stack(draws))
    allp = gather_posterior(torch.from_numpy(np.stack(local)), world).numpy()
    if rank == 0:
        q.put(allp)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_all_gather_matches_single_process():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import friedman
    from oracle.oracle_py import OracleChain
    from pymc_bart_b200.settings import make_settings

    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    allp = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert allp.shape == (4, 12, 120)
    # the same four global chains in one process
    X, y, _ = friedman(120, 5, 4)
    s = make_settings(X, y, m=6, num_particles=6, seed=4, chain_base=0)
    for g in range(4):
        o = OracleChain(s, X.T.copy(), y, chain=g)
        for d in range(12):
            o.step(d < 6, 1.0)
            assert np.array_equal(o.sum_trees(), allp[g, d])
