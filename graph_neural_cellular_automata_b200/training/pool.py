"""Sample pool (drop-in for the reference's training/pool.py:5-42).

Same constructor / `sample` / `replace` API and the same RNG draws (`random.sample(range(N), B)`; one `seed_fn`
call per slot at construction), but the pool is ONE device tensor [N,C,H,W] with gather / scatter instead of a
Python list of N tensors with B clones per step.
"""
from __future__ import annotations

import random
from typing import List, Sequence, Tuple

import torch


class SamplePool:
    def __init__(self, pool_size, seed_fn, device="cpu"):
        seeds = [seed_fn(batch_size=1).squeeze(0).to(device) for _ in range(pool_size)]
        self.pool = torch.stack(seeds).contiguous()          # [N,C,H,W]

    def __len__(self) -> int:
        return self.pool.shape[0]

    def _index(self, idx: Sequence[int]) -> torch.Tensor:
        """Slot indices on the pool's device.  On CUDA the list goes through pinned memory with a non-blocking copy: a
        pageable `torch.as_tensor(list, device=cuda)` would make the host wait for everything queued on the stream (one
        hidden synchronisation per training step)."""
        if self.pool.is_cuda:
            host = torch.tensor(list(idx), dtype=torch.int64).pin_memory()
            dev = host.to(self.pool.device, non_blocking=True)
            self._keep = (host, dev)             # the pinned source stays alive until the next call has queued its own copy
            return dev
        return torch.as_tensor(list(idx), device=self.pool.device)

    def sample(self, batch_size) -> Tuple[List[int], torch.Tensor]:
        idx = random.sample(range(len(self)), batch_size)
        self._last = (idx, self._index(idx))
        batch = self.pool[self._last[1]]                                     # gather = fresh copy
        return idx, batch

    def replace(self, idx: Sequence[int], new_samples: torch.Tensor) -> None:
        last = getattr(self, "_last", None)
        dev_idx = last[1] if last is not None and last[0] is idx else self._index(idx)
        self.pool[dev_idx] = new_samples.detach()
