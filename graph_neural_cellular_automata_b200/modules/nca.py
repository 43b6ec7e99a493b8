"""Drop-in for the reference's `modules/nca.py` (NeuralCA, nca.py:7-105): same constructor, attributes,
state-dict keys and RNG side effect (one `torch.rand(B,1,H,W)` iff fire_rate < 1); the step itself is the fused
CUDA path of libgnca.so (perception -> MLP on active cells -> masks -> per-sample GroupNorm -> bounded update ->
alpha post-gate) with a hand-written backward."""
from __future__ import annotations

import torch
import torch.nn as nn

from .perception import FixedSobelPerception
from ._base import FusedStepMixin


class NeuralCA(FusedStepMixin, nn.Module):
    _is_graph = False

    def __init__(self, n_channels: int, update_hidden: int = 128, img_size: int = 40, update_gain: float = 0.1,
                 alpha_thr: float = 0.1, use_groupnorm: bool = True, device: str = "cpu"):
        super().__init__()
        self.n_channels = n_channels
        self.img_size = img_size
        self.update_gain = update_gain
        self.alpha_thr = alpha_thr
        self.device = device
        self.perception = FixedSobelPerception(n_channels)
        self.update_net = nn.Sequential(
            nn.Conv2d(3 * n_channels, update_hidden, kernel_size=1, bias=True),
            nn.ReLU(inplace=False),
            nn.Conv2d(update_hidden, n_channels, kernel_size=1, bias=False),
        )
        nn.init.zeros_(self.update_net[-1].weight)            # nca.py:46
        self.norm = nn.GroupNorm(1, n_channels, eps=1e-3, affine=True) if use_groupnorm else nn.Identity()

    def forward(self, x: torch.Tensor, fire_rate: float = 1.0) -> torch.Tensor:
        fire_u = None
        if fire_rate < 1.0:                                   # nca.py:81-83: the only RNG draw
            fire_u = torch.rand(x.shape[0], 1, x.shape[2], x.shape[3], device=x.device)
        return self._fused_step(x, fire_rate, fire_u, chosen=(), message_gain=0.0)
