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
    parts = comm.allgather_np(np.full((2, 3), float(rank)))   # the per-rank replicate statistics of Sampler.replicate_stats
    ok &= len(parts) == world and all(parts[r].shape == (2, 3) and (parts[r] == r).all() for r in range(world))
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


def test_replicate_statistics_on_iid_draws():
    """combine_replicate_stats (the cross-rank half of Sampler.replicate_stats): for independent N(3, 2^2) draws the ESS
    of the posterior-mean estimator is the number of states (within the sqrt(2 / (B - 1)) noise of a variance
    estimated from B replicate means), split R-hat is 1, the pooled variance is exact across ranks with different
    means; a replicate that spans two ranks (DE-MC) counts once."""
    sys.path.insert(0, ROOT)
    from epi_mcmc_b200.sampler import combine_replicate_stats
    rng = np.random.default_rng(0)

    def part(x):   # x[k, p, B, cpr], what the device reduction produces for one rank
        k = x.shape[0]
        rep = x.mean(axis=(0, 3))
        tot = rep.mean(axis=1)
        cnt = float(x.size / x.shape[1])
        c, h = x[:, :, :, 0], k // 2
        halves = np.stack([c[:h], c[h:2 * h]])
        return dict(rep_mean=rep, hm=halves.mean(axis=1), hv=halves.var(axis=1, ddof=1), cnt=cnt, s1=tot * cnt,
                    s2=((x - tot[None, :, None, None]) ** 2).sum(axis=(0, 2, 3)))

    xs = [3 + 2 * rng.standard_normal((40, 2, 32, 64)) for _ in range(2)]
    n_states = 2 * 40 * 32 * 64
    for pooled, B in ((False, 64), (True, 32)):
        r = combine_replicate_stats([part(x) for x in xs], pooled, 20)
        assert r["n_replicates"] == B
        assert np.all(np.abs(r["ess"] / n_states - 1) < 4 * np.sqrt(2 / (B - 1)))
        assert np.all(np.abs(r["rhat"] - 1) < 0.02) and np.allclose(r["post_sd"], 2.0, rtol=0.01)
    # ranks with different means: the pooled variance is the variance of the union
    ys = [xs[0], xs[1] + 5.0]
    r = combine_replicate_stats([part(y) for y in ys], False, 20)
    union = np.concatenate([y.transpose(1, 0, 2, 3).reshape(2, -1) for y in ys], axis=1)
    assert np.allclose(r["post_sd"], union.std(axis=1, ddof=1), rtol=1e-12)
    # strongly autocorrelated replicates (every chain of a replicate shares one draw): ESS = number of replicates
    z = np.repeat(rng.standard_normal((1, 2, 64, 1)), 40, axis=0).repeat(16, axis=3)
    r = combine_replicate_stats([part(z)], False, 20)
    assert np.all(np.abs(r["ess"] / 64 - 1) < 0.05)
