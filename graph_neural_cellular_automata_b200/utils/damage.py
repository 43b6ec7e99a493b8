"""Damage operators (drop-in for the reference's utils/damage.py).

Every default damage kind is a multiplicative {0,1} (or soft) mask, so each operator here builds the mask
`D` on the device and applies `state *= D` with one fused kernel (gnca_apply_mask); the rollout can also take
`D` and apply it in-kernel at an arbitrary step (Schedule.damage / damage_step).  RNG consumption follows the
reference call for call (same `torch.randint` / `torch.rand_like` / `random` draws in the same order) but without
the per-sample `int(...)` host synchronisations: the drawn positions stay on the device.
"""
from __future__ import annotations

import ctypes as C
import random

import torch

from .. import _lib
from .. import functional as GF


def _apply(state: torch.Tensor, D: torch.Tensor) -> None:
    if not state.is_cuda:
        raise RuntimeError("damage operators run on CUDA tensors only (no CPU fallback)")
    if not state.is_contiguous() or state.dtype != torch.float32:
        raise RuntimeError("state must be a contiguous fp32 tensor")
    D = D.expand_as(state).contiguous()
    _lib.check(_lib.load().gnca_apply_mask(state.numel(), GF._ptr(state), GF._ptr(D), GF._stream()), "gnca_apply_mask")


def dense_mask(D: torch.Tensor, state: torch.Tensor) -> torch.Tensor:
    """The [B,C,H,W] multiplicative mask a builder's (broadcastable) result stands for."""
    return D.expand_as(state).contiguous()


def _grid(state):
    B, Cc, H, W = state.shape
    yy = torch.arange(H, device=state.device).view(1, H, 1)
    xx = torch.arange(W, device=state.device).view(1, 1, W)
    return B, Cc, H, W, yy, xx


def _draw_pairs(B, lo_y, hi_y, lo_x, hi_x, device):
    """B x (y, x) draws in the reference's order: randint(y) then randint(x) per sample (damage.py:21-22)."""
    vals = []
    for _ in range(B):
        vals.append(torch.randint(lo_y, hi_y, (1,), device=device))
        vals.append(torch.randint(lo_x, hi_x, (1,), device=device))
    v = torch.cat(vals).view(B, 2)
    return v[:, 0].view(B, 1, 1), v[:, 1].view(B, 1, 1)


@torch.no_grad()
def square_mask(state, size):
    B, Cc, H, W, yy, xx = _grid(state)
    y, x = _draw_pairs(B, 0, max(1, H - size + 1), 0, max(1, W - size + 1), state.device)
    hit = (yy >= y) & (yy < y + size) & (xx >= x) & (xx < x + size)
    return (~hit).float().unsqueeze(1)


@torch.no_grad()
def circle_mask(state, radius):
    B, Cc, H, W, yy, xx = _grid(state)
    cy, cx = _draw_pairs(B, radius, max(radius + 1, H - radius), radius, max(radius + 1, W - radius), state.device)
    hit = ((yy - cy) ** 2 + (xx - cx) ** 2) <= radius ** 2
    return (~hit).float().unsqueeze(1)


@torch.no_grad()
def stripe_mask(state, width, orientation="auto"):
    B, Cc, H, W, yy, xx = _grid(state)
    if orientation == "auto":
        orientation = "h" if random.random() < 0.5 else "v"
    if orientation == "h":
        s0 = torch.randint(0, max(1, H - width + 1), (1,), device=state.device)
        hit = ((yy >= s0) & (yy < s0 + width)).expand(1, H, W)
    else:
        s0 = torch.randint(0, max(1, W - width + 1), (1,), device=state.device)
        hit = ((xx >= s0) & (xx < s0 + width)).expand(1, H, W)
    return (~hit).float().unsqueeze(1).expand(B, 1, H, W)


@torch.no_grad()
def alpha_dropout_mask(state, p, alpha_thr=0.1, hard=True):
    alpha = state[:, 3:4]
    drop = (torch.rand_like(alpha) < p).float() * (alpha > alpha_thr).float()
    if hard:
        return 1.0 - drop
    D = torch.ones_like(state)
    D[:, 3:4] = 1.0 - drop
    return D


@torch.no_grad()
def salt_pepper_mask(state, p):
    D = torch.ones_like(state)
    D[:, 3:4] = 1.0 - (torch.rand_like(state[:, 3:4]) < p).float()
    return D


@torch.no_grad()
def gaussian_mask(state, radius, softness=0.35):
    B, Cc, H, W, yy, xx = _grid(state)
    cy, cx = _draw_pairs(B, radius, max(radius + 1, H - radius), radius, max(radius + 1, W - radius), state.device)
    r2 = ((yy - cy) ** 2 + (xx - cx) ** 2).float()
    m = torch.exp(-(r2 / (2.0 * (radius * max(1e-6, softness)) ** 2)))
    return (1.0 - m).clamp(0.0, 1.0).unsqueeze(1)


# ---- in-place operators with the reference's names / signatures (damage.py:16-98) -------------------------
@torch.no_grad()
def cutout_square_(state, size):
    if size > 0:
        _apply(state, square_mask(state, size))


@torch.no_grad()
def cutout_circle_(state, radius):
    if radius > 0:
        _apply(state, circle_mask(state, radius))


@torch.no_grad()
def stripe_wipe_(state, width, orientation="auto"):
    if width > 0:
        _apply(state, stripe_mask(state, width, orientation))


@torch.no_grad()
def alpha_dropout_(state, p, alpha_thr=0.1, hard=True):
    if p > 0:
        _apply(state, alpha_dropout_mask(state, p, alpha_thr, hard))


@torch.no_grad()
def salt_pepper_alpha_(state, p):
    if p > 0:
        _apply(state, salt_pepper_mask(state, p))


@torch.no_grad()
def gaussian_hole_(state, radius, softness=0.35):
    if radius > 0:
        _apply(state, gaussian_mask(state, radius, softness))


@torch.no_grad()
def hidden_scramble_(state, sigma=0.2):
    """Additive noise on hidden channels (weight 0 / absent in the default policy; kept for API parity)."""
    B, Cc, H, W = state.shape
    if Cc <= 4 or sigma <= 0:
        return
    noise = torch.randn(B, Cc - 4, H, W, device=state.device) * sigma
    state[:, 4:] = (state[:, 4:] + noise).clamp_(0.0, 1.0)


def _draw_policy(state, dmg_cfg, epoch):
    """The batch-level draws of apply_damage_policy_ (damage.py:101-121), in the reference's order: torch.rand(1) gate,
    random.choices(kind), random.randint(size).  Returns (kind, size) or None (no damage this batch)."""
    start_ep = int(dmg_cfg.get("start_epoch", dmg_cfg.get("damage_start_epoch", 100)))
    prob = float(dmg_cfg.get("prob", dmg_cfg.get("damage_prob", 0.0)))
    if epoch < start_ep or prob <= 0:
        return None
    if torch.rand(1, device=state.device).item() > prob:
        return None
    kinds = dmg_cfg.get("kinds", {"square": 1.0})
    names, weights = zip(*kinds.items())
    kind = random.choices(names, weights=weights, k=1)[0]
    size_min = int(dmg_cfg.get("size_min", dmg_cfg.get("damage_patch_size", 8)))
    size_max = int(dmg_cfg.get("size_max", max(size_min, 14)))
    return kind, int(random.randint(size_min, size_max))


def _policy_mask(state, dmg_cfg, kind, size):
    """Mask of the drawn (kind, size) with the kind's own geometry draws (damage.py:122-138); None = nothing to apply."""
    if kind == "circle":
        r = size // 2 if size > 1 else 1
        return circle_mask(state, r) if r > 0 else None
    if kind == "stripes":
        w = int(dmg_cfg.get("stripe_width", size))
        return stripe_mask(state, w, "auto") if w > 0 else None
    if kind == "alpha_drop":
        p = float(dmg_cfg.get("alpha_dropout_p", 0.1))
        return alpha_dropout_mask(state, p, float(dmg_cfg.get("alpha_thr", 0.1)), True) if p > 0 else None
    if kind == "saltpepper":
        p = float(dmg_cfg.get("salt_pepper_p", 0.02))
        return salt_pepper_mask(state, p) if p > 0 else None
    if kind == "gaussian":
        return gaussian_mask(state, max(1, size // 2), float(dmg_cfg.get("gaussian_softness", 0.35)))
    return square_mask(state, size) if size > 0 else None      # "square" and the reference's fallback


@torch.no_grad()
def sample_damage_mask(state, dmg_cfg, epoch):
    """The policy of apply_damage_policy_ (damage.py:101-138) returning the multiplicative mask instead of applying it
    (None = no damage this batch).  PURE: `state` is only read (shape, device, alpha for alpha_drop).  The additive
    `hidden_noise` kind (weight 0 / absent in the default policy) has no mask: it yields None here and is applied by
    apply_damage_policy_."""
    drawn = _draw_policy(state, dmg_cfg, epoch)
    if drawn is None or drawn[0] == "hidden_noise":
        return None
    return _policy_mask(state, dmg_cfg, *drawn)


@torch.no_grad()
def apply_damage_policy_(state, dmg_cfg, epoch):
    drawn = _draw_policy(state, dmg_cfg, epoch)
    if drawn is None:
        return
    if drawn[0] == "hidden_noise":
        hidden_scramble_(state, float(dmg_cfg.get("hidden_noise_sigma", 0.0)))
        return
    D = _policy_mask(state, dmg_cfg, *drawn)
    if D is not None:
        _apply(state, D)
