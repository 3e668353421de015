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
    [rank][d][ld_rank] — the addressing k_accept_propose uses (csrc/epi_sampler.cu)."""
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

    def barrier(self):
        if self.world > 1:
            if self.backend == "nccl":
                dist.barrier(device_ids=[self.local_rank])
            else:
                dist.barrier()

    # -- DE-MC population exchange on the library's device buffers
    def attach(self, sampler):
        th_ptr, _, _, ld = sampler.state_device()
        d = sampler.model.d
        dev = torch.device("cuda", sampler.model.device)
        self._state = torch.as_tensor(_DevPtr(th_ptr, d * ld), device=dev)
        self._pop = torch.empty(self.world * d * ld, dtype=torch.float64, device=dev)
        self._ld = ld
        self.exchange(sampler)

    def exchange(self, sampler):
        """all-gather every rank's d x ld state block into [rank][d][ld] and hand it to the
        proposal kernel as the new population snapshot"""
        if self.world == 1:
            self._pop.copy_(self._state)
        else:
            dist.all_gather_into_tensor(self._pop, self._state)
        torch.cuda.current_stream(self._pop.device).synchronize()
        sampler.set_population(self._pop.data_ptr(), self.world, sampler.n_chains, self._ld)

    def close(self):
        if self.world > 1 and dist.is_initialized():
            dist.destroy_process_group()
