"""Per-tensor gradient normalisation + Adam as ONE kernel over a flat parameter buffer
(train_graph_augmented_nca.py:370-375 with torch.optim.Adam(lr, weight_decay) semantics: coupled L2)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from .. import _lib
from .. import functional as GF


class FusedNormalizedAdam:
    """Owns a flat fp32 copy of the model's trainable (canonical) parameters; the nn.Parameters are re-pointed
    to views of it, so `state_dict()` / checkpoints keep working.  `step(gflat)` = for every parameter tensor
    `g /= ||g|| + 1e-8` (if `normalize`), then Adam.  Tensors the reference leaves without a gradient
    (`gate_mlp.*`, the frozen perception stencil) are not in the buffer and are never touched -- exactly like
    `optimizer.step()` skipping `p.grad is None`."""

    def __init__(self, model, lr: float = 2e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 normalize: bool = True):
        self.model = model
        self.lr, self.betas, self.eps, self.weight_decay, self.normalize = lr, betas, eps, weight_decay, normalize
        self.desc = model.model_desc()
        params = model.canonical_params()
        self.seg = list(GF.segment_offsets(self.desc))
        self.flat = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
        self.has_grad = []
        for i, p in enumerate(params):
            if isinstance(p, torch.nn.Parameter):
                p.data = self.flat[self.seg[i]:self.seg[i + 1]].view(p.shape)
                self.has_grad.append(1 if p.requires_grad else 0)
            else:                                   # placeholder gamma/beta when GroupNorm is off
                self.has_grad.append(0)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.step_count = 0
        model.invalidate_packed()

    def step(self, gflat: torch.Tensor, lr: Optional[float] = None) -> None:
        if gflat.numel() != self.flat.numel() or not gflat.is_cuda:
            raise RuntimeError("gflat must be the CUDA flat gradient in the canonical layout")
        self.step_count += 1
        n = len(self.seg) - 1
        seg = (C.c_int64 * (n + 1))(*self.seg)
        hg = (C.c_int32 * n)(*self.has_grad)
        _lib.check(_lib.load().gnca_normalize_adam(
            GF._ptr(self.flat), GF._ptr(gflat), GF._ptr(self.exp_avg), GF._ptr(self.exp_avg_sq), seg, hg, n,
            1 if self.normalize else 0, float(self.lr if lr is None else lr), float(self.betas[0]),
            float(self.betas[1]), float(self.eps), float(self.weight_decay), self.step_count, GF._stream()),
            "gnca_normalize_adam")
        self.model.invalidate_packed()

    def state_dict(self):
        return {"flat": self.flat.clone(), "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "step": self.step_count, "lr": self.lr}

    def load_state_dict(self, sd):
        self.flat.copy_(sd["flat"]); self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.step_count = int(sd["step"]); self.lr = float(sd.get("lr", self.lr))
        self.model.invalidate_packed()

    # ---- torch.optim.Adam-format state (what the reference's checkpoints hold: train...:405-416) -------------------
    def torch_state_dict(self, current_lr: Optional[float] = None):
        """State in `torch.optim.Adam(model.parameters()).state_dict()` format (indices = position in
        `model.parameters()`), so a checkpoint written here resumes in the reference's trainer and vice versa.
        `current_lr`: the learning rate the scheduler holds NOW (the trainer passes the decayed rate per step and never
        stores it here).  torch's StepLR is chainable: after a resume it multiplies `group["lr"]`, it does not recompute
        it from `base_lrs`, so the group must carry the decayed value and `initial_lr` the base one."""
        params = list(self.model.parameters())
        index = {id(p): i for i, p in enumerate(params)}
        state = {}
        if self.step_count > 0:
            for ci, p in enumerate(self.model.canonical_params()):
                if not isinstance(p, torch.nn.Parameter) or not self.has_grad[ci]:
                    continue
                sl = slice(self.seg[ci], self.seg[ci + 1])
                state[index[id(p)]] = {"step": torch.tensor(float(self.step_count)),
                                       "exp_avg": self.exp_avg[sl].view(p.shape).clone(),
                                       "exp_avg_sq": self.exp_avg_sq[sl].view(p.shape).clone()}
        group = {"lr": float(self.lr if current_lr is None else current_lr), "initial_lr": float(self.lr),
                 "betas": tuple(self.betas), "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "maximize": False, "foreach": None, "capturable": False, "differentiable": False,
                 "fused": None, "params": list(range(len(params)))}
        return {"state": state, "param_groups": [group]}

    def load_torch_state_dict(self, sd) -> None:
        """Inverse of `torch_state_dict`; states of parameters this optimiser does not own (`gate_mlp.*`) are ignored."""
        params = list(self.model.parameters())
        index = {id(p): i for i, p in enumerate(params)}
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps = float(g.get("initial_lr", g["lr"])), tuple(g["betas"]), float(g["eps"])
        self.weight_decay = float(g["weight_decay"])
        self.exp_avg.zero_(); self.exp_avg_sq.zero_()
        step = 0
        for ci, p in enumerate(self.model.canonical_params()):
            if not isinstance(p, torch.nn.Parameter) or not self.has_grad[ci]:
                continue
            st = sd["state"].get(index[id(p)])
            if st is None:
                continue
            sl = slice(self.seg[ci], self.seg[ci + 1])
            self.exp_avg[sl].copy_(st["exp_avg"].reshape(-1).to(self.exp_avg.device))
            self.exp_avg_sq[sl].copy_(st["exp_avg_sq"].reshape(-1).to(self.exp_avg.device))
            step = max(step, int(float(st["step"])))
        self.step_count = step
