"""Drop-in for the reference's `modules/perception.py` (FixedSobelPerception, perception.py:5-26).

Same constructor, same `conv.weight` state-dict entry (frozen [3C,1,3,3] identity/sobel_x/sobel_y stack), same
output channel order [identity(C) | sobel_x(C) | sobel_y(C)].  The forward is one fused CUDA kernel
(gnca_perception_fwd) with an explicit transpose kernel for the backward; the Conv2d is only a parameter
container so shipped checkpoints load unchanged.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import functional as F_gnca

_STENCILS = (
    ((0.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 0.0)),       # identity
    ((1.0, 0.0, -1.0), (2.0, 0.0, -2.0), (1.0, 0.0, -1.0)),    # sobel_x  (perception.py:9)
    ((1.0, 2.0, 1.0), (0.0, 0.0, 0.0), (-1.0, -2.0, -1.0)),    # sobel_y  (perception.py:10)
)


class FixedSobelPerception(nn.Module):
    def __init__(self, n_channels):
        super().__init__()
        self.n_channels = int(n_channels)
        # same layer type/shape as the reference so RNG consumption at construction and state_dict keys match
        self.conv = nn.Conv2d(n_channels, 3 * n_channels, 3, 1, 1, groups=n_channels, bias=False)
        stencil = torch.tensor(_STENCILS, dtype=torch.float32).unsqueeze(1)          # [3,1,3,3]
        with torch.no_grad():
            self.conv.weight.copy_(stencil.repeat(n_channels, 1, 1, 1))
        self.conv.weight.requires_grad_(False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F_gnca.perception(x)
