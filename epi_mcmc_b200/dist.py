"""Multi-GPU plumbing: one process per GPU, chains sharded by rank, and ONE collective on the
whole path — the all-gather of chain states that refreshes the DE-MC population every K
iterations.  Likelihood sweeps need no communication at all.

torch.distributed is used for what it is good at (rendezvous + NCCL over NVLink/NVSwitch);
the buffers it moves are the library's own device buffers, wrapped zero-copy.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


class _DevPtr:
    """zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface v3)"""

    def __init__(self, ptr: int, n_doubles: int):
        self.__cuda_array_interface__ = {"shape": (n_doubles,), "typestr": "<f8", "data": (ptr, False), "version": 3}


def shard(n_total: int, rank: int, world: int):
    """contiguous block of chains owned by `rank`: (first global chain id, count)"""
    base, rem = divmod(n_total, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def partner_address(z: int, chains_per_rank: int, d: int, ld_rank: int, p: int) -> int:
    """offset (in doubles) of parameter p of population member z inside the gathered buffer
    [rank][d][ld_rank] — the addressing k_pop_to_aos reads when it turns the gathered block into the member-major
    snapshot the proposal kernel draws partners from (csrc/epi_sampler.cu)."""
    return ((z // chains_per_rank) * d + p) * ld_rank + (z % chains_per_rank)


class Comm:
    def __init__(self, backend: str | None = None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if self.world > 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            if self.backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                dist.init_process_group("nccl", rank=self.rank, world_size=self.world,
                                        device_id=torch.device("cuda", self.local_rank))
            else:
                dist.init_process_group(self.backend, rank=self.rank, world_size=self.world)
        self._pop = None
        self._state = None

    # -- generic helpers (work on CPU tensors with gloo: used by the world_size-2 CPU tests)
    def all_gather_state(self, local: torch.Tensor) -> torch.Tensor:
        """local: this rank's flat d*ld state block.  Returns the [world, d*ld] population."""
        if self.world == 1:
            return local.reshape(1, -1).clone()
        out = torch.empty(self.world * local.numel(), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous().reshape(-1))
        return out.reshape(self.world, -1)

    def max_over_ranks(self, value: float) -> float:
        if self.world == 1:
            return float(value)
        dev = torch.device("cuda", self.local_rank) if self.backend == "nccl" else torch.device("cpu")
        t = torch.tensor([value], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allreduce_sum(self, values):
        """element-wise sum of a small float64 vector over the ranks (returns a numpy array)"""
        import numpy as np
        v = np.asarray(values, dtype=np.float64)
        if self.world == 1:
            return v
        dev = torch.device("cuda", self.local_rank) if self.backend == "nccl" else torch.device("cpu")
        t = torch.tensor(v, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def allgather_np(self, a):
        """the same-shaped float64 numpy array of every rank, as a list ordered by rank"""
        import numpy as np
        a = np.ascontiguousarray(a, dtype=np.float64)
        if self.world == 1:
            return [a]
        dev = torch.device("cuda", self.local_rank) if self.backend == "nccl" else torch.device("cpu")
        t = torch.from_numpy(a.reshape(-1)).to(dev)
        out = torch.empty(self.world * t.numel(), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(out, t)
        return [o.reshape(a.shape) for o in out.cpu().numpy().reshape(self.world, -1)]

    def barrier(self):
        if self.world > 1:
            if self.backend == "nccl":
                dist.barrier(device_ids=[self.local_rank])
            else:
                dist.barrier()

    # -- DE-MC population exchange on the library's device buffers
    def attach(self, sampler):
        """Two population buffers and a side stream: the all-gather started by exchange() k
        fills buffer k % 2 while the chains already run their next iterations against the snapshot of exchange k - 1."""
        th_ptr, _, _, ld = sampler.state_device()
        d = sampler.model.d
        dev = torch.device("cuda", sampler.model.device)
        self._state = torch.as_tensor(_DevPtr(th_ptr, d * ld), device=dev)
        self._stage = torch.empty(d * ld, dtype=torch.float64, device=dev)
        self._pops = [torch.empty(self.world * d * ld, dtype=torch.float64, device=dev) for _ in range(2)]
        self._pop = self._pops[0]
        self._side = torch.cuda.Stream(device=dev)
        self._done = [None, None]
        self._k = 0
        self._ld = ld
        self.exchange(sampler, wait=True)

    def exchange(self, sampler, wait: bool = False):
        """Start the all-gather of every rank's d x ld state block into [rank][d][ld] on the side stream and hand
        the proposal kernel the newest COMPLETED snapshot.  The chains' state is first copied aside (a device copy
        of d x ld doubles, the only thing the host waits for: the kernels of the next iterations overwrite the state
        while the collective is still reading), so the collective overlaps the next K1 launches; its result is
        picked up by the following call.  wait=True (first call, tests): block until this very snapshot is in."""
        dev = self._stage.device
        k = self._k
        with torch.cuda.stream(self._side):
            self._stage.copy_(self._state)
            staged = torch.cuda.Event()
            staged.record(self._side)
            if self.world == 1:
                self._pops[k % 2].copy_(self._stage)
            else:
                dist.all_gather_into_tensor(self._pops[k % 2], self._stage)
            ev = torch.cuda.Event()
            ev.record(self._side)
        self._done[k % 2] = ev
        self._k += 1
        if wait or k == 0:
            ev.synchronize()
            use = k % 2
        else:
            staged.synchronize()                 # the state may be overwritten from here on
            use = (k - 1) % 2
            self._done[use].synchronize()        # started a whole exchange period ago: normally long complete
        self._pop = self._pops[use]
        sampler.set_population(self._pop.data_ptr(), self.world, sampler.n_chains, self._ld)

    def close(self):
        if self.world > 1 and dist.is_initialized():
            dist.destroy_process_group()
