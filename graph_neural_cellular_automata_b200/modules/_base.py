"""Shared host logic of NeuralCA / NeuralCAGraph: descriptor construction, canonical parameter order,
packed-weight caching and the call into the fused step."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from .. import functional as F_gnca


class FusedStepMixin:
    _is_graph = False

    # ---- reference API: nca.py:55-62 / ncagraph.py:85-92 -------------------------------------------------
    @torch.no_grad()
    def _alive_mask(self, x: torch.Tensor) -> torch.Tensor:
        return F_gnca.alive_mask(x, float(self.alpha_thr))

    # ---- canonical parameter list (order of gnca_param_layout) ---------------------------------------------
    def canonical_params(self) -> List[torch.Tensor]:
        dev = self.update_net[0].weight.device
        C = self.n_channels
        if isinstance(self.norm, nn.GroupNorm):
            gamma, beta = self.norm.weight, self.norm.bias
        else:                                   # use_groupnorm=False: layout keeps the slots, kernel ignores them
            gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        ps = [self.update_net[0].weight, self.update_net[0].bias, self.update_net[2].weight, gamma, beta]
        if self._is_graph:
            g = self.graph
            ps += [g.msg_proj.weight, g.msg_proj.bias, g.query_proj.weight, g.query_proj.bias,
                   g.key_proj.weight, g.key_proj.bias, g.scaling]
        return ps

    def model_desc(self):
        g = getattr(self, "graph", None)
        return F_gnca.make_model_desc(
            self.n_channels, self.update_net[0].out_channels, g.d_model if g is not None else 0,
            graph=self._is_graph, torus=(g is not None and not g.zero_padded_shift),
            hidden_only=bool(getattr(self, "hidden_only", False)),
            alive_to_alive=(g is not None and g.alive_to_alive),
            groupnorm=isinstance(self.norm, nn.GroupNorm), update_gain=float(self.update_gain),
            alpha_thr=float(self.alpha_thr), graph_alpha_thr=float(g.alpha_thr) if g is not None else float(self.alpha_thr),
            gn_eps=float(self.norm.eps) if isinstance(self.norm, nn.GroupNorm) else 1e-3)

    def flat_params(self) -> torch.Tensor:
        return torch.cat([p.detach().reshape(-1) for p in self.canonical_params()])

    def packed_weights(self) -> torch.Tensor:
        """Kernel-side weight buffer, rebuilt only when a parameter changed (version counters / storage)."""
        ps = self.canonical_params()
        key = tuple((p.data_ptr(), p._version) for p in ps)
        cache = self.__dict__.get("_gnca_packed")
        if cache is None or cache[0] != key:
            packed = F_gnca.pack_weights(self.model_desc(), self.flat_params())
            cache = (key, packed)
            self.__dict__["_gnca_packed"] = cache
        return cache[1]

    def step(self, x: torch.Tensor, fire_rate: float = 1.0, *, fire_u: Optional[torch.Tensor] = None,
             chosen: Optional[Sequence[Tuple[int, int]]] = None, message_gain: Optional[float] = None,
             return_attention: bool = False):
        """`forward` with the randomness made explicit (extension, used by the parity tests and the rollout):
        `fire_u` = the uniforms `torch.rand(B,1,H,W)` would have produced, `chosen` = the `random.sample` result.
        Anything left None is drawn exactly as `forward` draws it."""
        if self._is_graph:
            if chosen is None:
                chosen = self.graph.draw_offsets()
            gain = float(self.message_gain if message_gain is None else message_gain)
        else:
            chosen, gain = (), 0.0
        if fire_rate < 1.0 and fire_u is None:
            fire_u = torch.rand(x.shape[0], 1, x.shape[2], x.shape[3], device=x.device)
        return self._fused_step(x, fire_rate, fire_u, chosen=chosen, message_gain=gain, want_attn=return_attention)

    def invalidate_packed(self) -> None:
        """Drop the cached kernel-side weights (needed after an optimiser writes the parameters through raw
        device pointers, which does not bump the tensors' version counters)."""
        self.__dict__.pop("_gnca_packed", None)

    def _fused_step(self, x, fire_rate, fire_u, *, chosen: Sequence[Tuple[int, int]], message_gain: float,
                    want_attn: bool = False):
        if not x.is_cuda:
            raise RuntimeError(f"{type(self).__name__}: input is on {x.device}; this implementation runs on CUDA "
                               "(sm_100a) only and has no CPU fallback")
        ps = self.canonical_params()
        if ps[0].device != x.device:
            raise RuntimeError(f"model parameters are on {ps[0].device}, input on {x.device}")
        return F_gnca.nca_step(x, ps, self.model_desc(), self.packed_weights(), fire_rate=float(fire_rate),
                               fire_u=fire_u, chosen=chosen, message_gain=message_gain, want_attn=want_attn)
