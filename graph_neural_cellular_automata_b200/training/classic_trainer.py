"""One training iteration of the CLASSIC trainer with its live stability phase
(train_intermediate_loss.py:230-296) on the fused CUDA path.

    pool sample -> per-sample-length rollout (fire rate ~ U(0.5, 1) per step) -> target-masked loss
    -> STABILITY PHASE: the samples already close to the target (per-sample loss < 0.01) are rolled K = 24 more steps
       from their terminal states and their drift from the target is penalised (weight 0.5)
    -> joint backward through both rollouts -> clip_grad_norm_(0.5) -> Adam -> worst-k / random reseed -> pool.

Both rollouts are `rollout()` calls (one cluster-resident launch each, resident BPTT); autograd chains the second
through `x_T[close]` into the first.  RNG consumption follows the reference call for call (python `random`:
pool.sample, regime, reseed; torch device generator: step counts, per-step fire rate, per-step fire uniforms of the
ACTIVE samples, reseed index, seed noise).  Loss arithmetic, clipping and Adam are stock PyTorch on ~8k floats --
host-side glue exactly as in the reference; no step / rollout arithmetic runs in PyTorch.
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Callable, Dict, Optional

import numpy as np
import torch

from ..rollout import make_schedule, rollout
from ..utils.nca_init import trainer_seed
from .pool import SamplePool


@dataclass
class ClassicTrainConfig:
    batch_size: int = 16
    pool_size: int = 1024
    nca_steps_min: int = 48                 # config["training"]["nca_steps_min"/"nca_steps_max"]
    nca_steps_max: int = 80
    long_prob: float = 0.25                 # train_intermediate_loss.py:170-171
    long_min: int = 200
    long_max: int = 400
    fire_rate_min: float = 0.5              # :245 uniform_(0.5, 1.0)
    fire_rate_max: float = 1.0
    learning_rate: float = 2e-4
    weight_decay: float = 1e-5
    loss_alpha_thr: float = 0.2             # masked_loss(alpha_thr=0.2, lam_area=5e-5)  :253
    loss_lam_area: float = 5e-5
    stability_threshold: float = 0.01       # :258
    stability_steps: int = 24               # :262
    stability_weight: float = 0.5           # :267
    clip_grad_norm: float = 0.5             # :283
    reset_worst_prob: float = 0.10          # :177
    random_reseed_prob: float = 0.05        # :178
    rollout_impl: str = "auto"


def masked_loss(pred: torch.Tensor, target: torch.Tensor, alpha_thr: float = 0.2, lam_area: float = 5e-5):
    """train_intermediate_loss.py:37-51: MSE inside the TARGET's alpha support + a tiny area penalty; per sample."""
    target_mask = (target[:, 3:4] > alpha_thr).float()
    mse = ((pred - target) ** 2) * target_mask
    denom = target_mask.sum(dim=(1, 2, 3)) + 1e-8
    per_sample = mse.sum(dim=(1, 2, 3)) / denom
    return per_sample + lam_area * pred[:, 3:4].mean(dim=(1, 2, 3))


class ClassicNCATrainer:
    def __init__(self, model, target: torch.Tensor, cfg: ClassicTrainConfig, *, seed_fn: Optional[Callable] = None):
        self.model, self.cfg = model, cfg
        self.device = next(model.parameters()).device
        self.target = target.to(self.device).float().contiguous()           # [4,H,W]
        self.n_ch, self.img = model.n_channels, model.img_size
        self.seed_fn = seed_fn or (lambda batch_size=1: trainer_seed(self.n_ch, self.img, batch_size, self.device))
        self.pool = SamplePool(cfg.pool_size, self.seed_fn, device=self.device)
        self.opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=cfg.learning_rate,
                                    weight_decay=cfg.weight_decay)
        self.last: Dict = {}

    def _draw_rollout(self, n_rows: int, steps_host: np.ndarray, T: int):
        """Per step: fire rate (one device uniform) then the fire uniforms of the samples still running, in index order
        (nca.py:81-83 draws torch.rand(n_active,1,H,W) inside model(state[mask]))."""
        cfg, dev, H, W = self.cfg, self.device, self.img, self.img
        fr_dev = torch.empty(T, dtype=torch.float32, device=dev)
        fire_u = torch.empty(T, n_rows, 1, H, W, dtype=torch.float32, device=dev)
        for t in range(T):
            act = np.nonzero(steps_host > t)[0]
            if len(act) == 0:                               # `if mask.any()`: nothing is drawn for this step
                fr_dev[t] = 1.0
                continue
            fr_dev[t:t + 1].uniform_(cfg.fire_rate_min, cfg.fire_rate_max)
            if len(act) == n_rows:
                torch.rand(n_rows, 1, H, W, out=fire_u[t])
            else:
                fire_u[t, torch.as_tensor(act, device=dev)] = torch.rand(len(act), 1, H, W, device=dev)
        sched = make_schedule(self.model, n_rows, H, W, T, fire_rate=[0.0] * T, steps=steps_host.tolist(), seed=0, device=dev)
        sched.fire_rate = fr_dev
        sched.fire_u = fire_u.view(T, n_rows, H, W)
        return sched

    def train_step(self) -> Dict:
        cfg, dev = self.cfg, self.device
        B = cfg.batch_size
        idx, batch = self.pool.sample(B)                                          # :232
        if random.random() < cfg.long_prob:                                       # :234
            lo, hi = cfg.long_min, cfg.long_max
        else:
            lo, hi = cfg.nca_steps_min, cfg.nca_steps_max
        nca_steps = torch.randint(lo, hi + 1, (B,), device=dev)                   # :239
        steps_host = nca_steps.cpu().numpy()
        sched = self._draw_rollout(B, steps_host, int(steps_host.max()))
        state = rollout(self.model, batch.contiguous(), sched, impl=cfg.rollout_impl)          # :243-247
        target_b = self.target.unsqueeze(0).expand(B, -1, -1, -1)
        per_sample = masked_loss(state[:, :4], target_b, cfg.loss_alpha_thr, cfg.loss_lam_area)   # :253
        loss = per_sample.mean()
        with torch.no_grad():
            close = per_sample < cfg.stability_threshold                           # :258
        stab = None
        n_close = int(close.sum())
        if n_close > 0:                                                            # :260-267
            K = cfg.stability_steps
            sched2 = self._draw_rollout(n_close, np.full(n_close, K), K)
            stab_state = rollout(self.model, state[close].contiguous(), sched2, impl=cfg.rollout_impl)
            stab = torch.nn.functional.mse_loss(stab_state[:, :4], target_b[close])
            loss = loss + cfg.stability_weight * stab
        n_reset = int(cfg.reset_worst_prob * B)                                    # :269-272
        worst = torch.topk(per_sample, n_reset).indices if n_reset > 0 else None
        do_reseed = random.random() < cfg.random_reseed_prob                       # :274-277
        rand_idx = int(torch.randint(0, B, (1,), device=dev).item()) if do_reseed else None
        self.opt.zero_grad(set_to_none=True)                                       # :280-284
        loss.backward()
        params = [p for p in self.model.parameters() if p.requires_grad]
        gnorm = torch.nn.utils.clip_grad_norm_(params, cfg.clip_grad_norm)
        grads = {n: p.grad.detach().clone() for n, p in self.model.named_parameters() if p.grad is not None}
        self.opt.step()
        new_states = state.detach()
        if worst is not None or do_reseed:                                         # :287-296
            new_states = new_states.clone()
            if worst is not None:
                new_states[worst] = self.seed_fn(len(worst)).detach()
            if do_reseed:
                new_states[rand_idx:rand_idx + 1] = self.seed_fn(1).detach()
        self.pool.replace(idx, new_states)
        self.last = {"per_sample": per_sample.detach(), "loss": loss.detach(), "stab": None if stab is None else stab.detach(),
                     "close": close, "steps": steps_host, "worst": worst, "grad_norm": gnorm, "idx": idx, "grads": grads}
        return self.last
