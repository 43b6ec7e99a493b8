"""Data-parallel plumbing of the training step: the global batch is sharded over the ranks of one box,
every rank replays the SAME seeded host RNG (pool indices, step counts, per-step fire rate / gain / offsets are
therefore identical everywhere) and works on its slice; per step there is exactly one small collective on the data
path -- a SUM all-reduce of the flat parameter gradient (~37 KB) -- plus an all-gather of the B per-sample losses
(so the worst-k pool reset is decided on the global batch, bit-exactly) and of the final states for the replicated
pool.  Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


class Shard:
    def __init__(self, global_batch: int, rank: Optional[int] = None, world: Optional[int] = None):
        on = dist.is_available() and dist.is_initialized()
        self.rank = rank if rank is not None else (dist.get_rank() if on else 0)
        self.world = world if world is not None else (dist.get_world_size() if on else 1)
        if global_batch % self.world != 0:
            raise ValueError(f"global batch {global_batch} is not divisible by world size {self.world}")
        self.global_batch = global_batch
        self.local_batch = global_batch // self.world
        self.lo = self.rank * self.local_batch
        self.hi = self.lo + self.local_batch

    def take(self, t: torch.Tensor) -> torch.Tensor:
        """This rank's contiguous slice of a tensor whose dim 0 is the global batch."""
        return t[self.lo:self.hi]

    # -- collectives --------------------------------------------------------------------------------------
    def _host_staged(self, t: torch.Tensor) -> bool:
        """gloo has no CUDA transport here: CUDA tensors go through the host (the single-GPU DP equivalence test runs two
        ranks on one device with gloo; on the box the backend is NCCL and nothing is staged)."""
        return t.is_cuda and dist.get_backend() == "gloo"

    def allreduce_sum_(self, flat: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            if self._host_staged(flat):
                h = flat.cpu()
                dist.all_reduce(h, op=dist.ReduceOp.SUM)
                flat.copy_(h)
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        return flat

    def allgather(self, local: torch.Tensor) -> torch.Tensor:
        """[local_batch, ...] -> [global_batch, ...] in rank order."""
        if self.world == 1:
            return local
        if self._host_staged(local):
            parts = [torch.empty(local.shape, dtype=local.dtype) for _ in range(self.world)]
            dist.all_gather(parts, local.cpu().contiguous())
            return torch.cat(parts, 0).to(local.device)
        out = torch.empty((self.global_batch,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out


def worst_k_indices(per_sample_global: torch.Tensor, frac: float) -> Optional[torch.Tensor]:
    """`torch.topk(per_sample, int(frac * B)).indices` (train...:378-380) on the GLOBAL batch."""
    n_reset = int(float(frac) * per_sample_global.numel())
    return torch.topk(per_sample_global, n_reset).indices if n_reset > 0 else None
