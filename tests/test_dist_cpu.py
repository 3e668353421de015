"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: chain sharding, the all-gather
layout the DE-MC proposal kernel addresses, max-over-ranks timing."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import philox_ref
from conftest import ROOT


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    from epi_mcmc_b200.dist import Comm, partner_address, shard
    comm = Comm(backend="gloo")
    d, n, ld = 5, 7, 32
    first, cnt = shard(2 * n, rank, world)
    local = torch.zeros(d * ld, dtype=torch.float64)
    for p in range(d):
        for i in range(n):
            local[p * ld + i] = 1000 * (first + i) + p      # value encodes (global chain, parameter)
    pop = comm.all_gather_state(local)
    flat = pop.reshape(-1).numpy()
    ok = True
    for z in range(world * n):
        for p in range(d):
            ok &= flat[partner_address(z, n, d, ld, p)] == 1000 * z + p
    mx = comm.max_over_ranks(1.0 + rank)
    tot = comm.allreduce_sum([10.0 + rank, 100.0])     # the acceptance totals Sampler.tune_scale reduces
    ok &= bool(tot[0] == 21.0 and tot[1] == 200.0)
    comm.barrier()
    q.put((rank, bool(ok), mx, first, cnt))
    comm.close()


def test_allgather_layout_and_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=60) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1]
    assert res[0][2] == 2.0 and res[1][2] == 2.0          # MAX over ranks
    assert (res[0][3], res[0][4]) == (0, 7) and (res[1][3], res[1][4]) == (7, 7)


def test_shard_covers_everything():
    sys.path.insert(0, ROOT)
    from epi_mcmc_b200.dist import shard
    for total in (1, 7, 64, 65536, 65537):
        for world in (1, 2, 3, 8):
            blocks = [shard(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == total
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + c0 == f1


def test_de_partners_distinct_and_uniform():
    tot = 1000
    seen = np.zeros(tot)
    for chain in range(200):
        for it in range(1, 6):
            a, b = philox_ref.de_partners(42, chain, it, tot)
            assert a != b and 0 <= a < tot and 0 <= b < tot
            seen[a] += 1
            seen[b] += 1
    assert seen.max() < 12       # 2000 draws over 1000 slots


def test_ess_estimator_on_known_ar1():
    """coda-style spectral ESS: for AR(1) with coefficient rho, ESS/n -> (1-rho)/(1+rho)"""
    sys.path.insert(0, ROOT)
    from epi_mcmc_b200.sampler import effective_size
    rng = np.random.default_rng(0)
    n = 20000
    for rho in (0.0, 0.5, 0.9):
        e = rng.standard_normal(n)
        x = np.zeros(n)
        for i in range(1, n):
            x[i] = rho * x[i - 1] + e[i]
        assert effective_size(x) / n == pytest.approx((1 - rho) / (1 + rho), rel=0.15)
