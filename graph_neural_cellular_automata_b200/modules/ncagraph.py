"""Drop-in for the reference's `modules/ncagraph.py` (NeuralCAGraph, ncagraph.py:10-168).

Constructor signature, attributes (`message_gain` is read at call time -- the trainer mutates it every step,
train_graph_augmented_nca.py:318-324), `graph` sub-module, state-dict keys and RNG side effects are the
reference's: per forward exactly one `random.sample(graph.offsets, k)` (also when message_gain == 0) followed by
one `torch.rand(B,1,H,W)` iff fire_rate < 1.  The arithmetic is the fused CUDA step of libgnca.so.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .perception import FixedSobelPerception
from .graph_augmentation import GraphAugmentation
from ._base import FusedStepMixin


class NeuralCAGraph(FusedStepMixin, nn.Module):
    _is_graph = True

    def __init__(self, n_channels: int, update_hidden: int = 128, img_size: int = 40, update_gain: float = 0.1,
                 alpha_thr: float = 0.1, use_groupnorm: bool = True, *, message_gain: float = 0.5,
                 hidden_only: bool = True, graph_d_model: int = 16, graph_attention_radius: int = 4,
                 graph_num_neighbors: int = 8, graph_gating_hidden: int = 32, graph_alive_to_alive: bool = True,
                 graph_zero_padded_shift: bool = True, device: str = "cpu"):
        super().__init__()
        self.n_channels = n_channels
        self.img_size = img_size
        self.update_gain = float(update_gain)
        self.alpha_thr = float(alpha_thr)
        self.device = device
        self.perception = FixedSobelPerception(n_channels)
        self.update_net = nn.Sequential(
            nn.Conv2d(3 * n_channels, update_hidden, kernel_size=1, bias=True),
            nn.ReLU(inplace=False),
            nn.Conv2d(update_hidden, n_channels, kernel_size=1, bias=False),
        )
        nn.init.zeros_(self.update_net[-1].weight)            # ncagraph.py:65
        self.norm = nn.GroupNorm(1, n_channels, eps=1e-3, affine=True) if use_groupnorm else nn.Identity()
        self.graph = GraphAugmentation(
            n_channels=n_channels, d_model=graph_d_model, attention_radius=graph_attention_radius,
            num_neighbors=graph_num_neighbors, gating_hidden=graph_gating_hidden,
            alive_to_alive=graph_alive_to_alive, zero_padded_shift=graph_zero_padded_shift, alpha_thr=self.alpha_thr)
        self.message_gain = float(message_gain)
        self.hidden_only = bool(hidden_only)

    def forward(self, x: torch.Tensor, fire_rate: float = 1.0, *, return_attention: bool = False):
        chosen = self.graph.draw_offsets()                    # graph_augmentation.py:120-121, always drawn
        fire_u = None
        if fire_rate < 1.0:                                   # ncagraph.py:144-146
            fire_u = torch.rand(x.shape[0], 1, x.shape[2], x.shape[3], device=x.device)
        return self._fused_step(x, fire_rate, fire_u, chosen=chosen, message_gain=float(self.message_gain),
                                want_attn=return_attention)
