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

    def sample(self, batch_size) -> Tuple[List[int], torch.Tensor]:
        idx = random.sample(range(len(self)), batch_size)
        batch = self.pool[torch.as_tensor(idx, device=self.pool.device)]     # gather = fresh copy
        return idx, batch

    def replace(self, idx: Sequence[int], new_samples: torch.Tensor) -> None:
        self.pool[torch.as_tensor(list(idx), device=self.pool.device)] = new_samples.detach()
