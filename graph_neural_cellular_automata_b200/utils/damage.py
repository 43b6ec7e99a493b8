"""Damage operators (drop-in for the reference's utils/damage.py).

Every default damage kind is a multiplicative {0,1} (or soft) per-CELL mask, so a damage event is a small DESCRIPTOR --
kind, size, per-sample positions (the `torch.randint` draws, kept on the device), or the `torch.rand_like` uniforms for
the two stochastic kinds -- that one CUDA kernel (`gnca_damage_plane`, include/gnca.h) evaluates in closed form into a
plane [B,H,W].  The in-place operators multiply the state by that plane (`gnca_apply_plane`); a rollout takes the
`Damage` object and applies it in-kernel at an arbitrary step (Schedule.damage / damage_step, regeneration protocol of
test_graph_augmented_regeneration.py:185-189).  No [B,C,H,W] mask is ever built (1 GiB at 256x256x32, B=128) and no
eager-torch mask arithmetic runs.

RNG consumption follows the reference call for call (same `torch.randint` / `torch.rand_like` / `random` draws in the
same order) but without the per-sample `int(...)` host synchronisations: the drawn positions stay on the device.
`fast=True` (the trainer's production mode) draws the geometry with ONE host call instead of 2B device calls -- same
distribution, different stream, nothing waits for the device.
"""
from __future__ import annotations

import ctypes as C
import random
from dataclasses import dataclass
from typing import Optional

import torch

from .. import _lib
from .. import functional as GF
from .._lib import GncaDamage

DK_SQUARE, DK_CIRCLE, DK_STRIPE_H, DK_STRIPE_V, DK_GAUSSIAN, DK_ALPHA_DROP, DK_SALTPEPPER = 1, 2, 3, 4, 5, 6, 7
LAYOUT_DENSE, LAYOUT_PLANE, LAYOUT_PLANE_ALPHA = 0, 1, 2


@dataclass
class Damage:
    """An evaluated damage event: per-cell plane [B,H,W] + how it applies (all channels / alpha only)."""
    plane: torch.Tensor
    layout: int = LAYOUT_PLANE

    def dense(self, state: torch.Tensor) -> torch.Tensor:
        """The [B,C,H,W] multiplicative mask this event stands for (tests / inspection only)."""
        B, Cc, H, W = state.shape
        if self.layout == LAYOUT_PLANE_ALPHA:
            D = torch.ones_like(state)
            D[:, 3] = self.plane
            return D
        return self.plane.unsqueeze(1).expand(B, Cc, H, W).contiguous()

    def expand_as(self, state: torch.Tensor) -> torch.Tensor:       # call shape of the round-1 mask tensors
        return self.dense(state)

    def take(self, lo: int, hi: int) -> "Damage":
        return Damage(self.plane[lo:hi].contiguous(), self.layout)


def dense_mask(D, state: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] mask of a `Damage` (or of a broadcastable mask tensor)."""
    return D.dense(state) if isinstance(D, Damage) else D.expand_as(state).contiguous()


def _check(state: torch.Tensor) -> None:
    if not state.is_cuda:
        raise RuntimeError("damage operators run on CUDA tensors only (no CPU fallback)")
    if not state.is_contiguous() or state.dtype != torch.float32:
        raise RuntimeError("state must be a contiguous fp32 tensor")


def _plane(state: torch.Tensor, kind: int, size: int = 0, *, pos: Optional[torch.Tensor] = None,
           rand: Optional[torch.Tensor] = None, p: float = 0.0, alpha_thr: float = 0.0, softness: float = 0.35) -> Damage:
    _check(state)
    B, Cc, H, W = state.shape
    plane = torch.empty(B, H, W, dtype=torch.float32, device=state.device)
    if pos is not None:
        pos = pos.to(device=state.device, dtype=torch.int64).contiguous()
    if rand is not None:
        rand = GF._require_cuda_f32(rand, "rand")
    d = GncaDamage(int(kind), int(size), float(softness), float(p), float(alpha_thr), 0,
                   C.c_void_p(0 if pos is None else pos.data_ptr()), C.c_void_p(0 if rand is None else rand.data_ptr()))
    layout = C.c_int32(0)
    _lib.check(_lib.load().gnca_damage_plane(C.byref(d), B, Cc, H, W, GF._ptr(state), GF._ptr(plane), C.byref(layout),
                                             GF._stream()), "gnca_damage_plane")
    return Damage(plane, int(layout.value))


def apply_damage_(state: torch.Tensor, D) -> None:
    """state *= D in place (D: `Damage`, or a mask tensor broadcastable to the state)."""
    _check(state)
    lib = _lib.load()
    if isinstance(D, Damage):
        B, Cc, H, W = state.shape
        _lib.check(lib.gnca_apply_plane(B, Cc, H, W, GF._ptr(state), GF._ptr(D.plane), int(D.layout), GF._stream()),
                   "gnca_apply_plane")
    else:
        M = D.expand_as(state).contiguous()
        _lib.check(lib.gnca_apply_mask(state.numel(), GF._ptr(state), GF._ptr(M), GF._stream()), "gnca_apply_mask")


def _draw_pairs(B, lo_y, hi_y, lo_x, hi_x, device, fast=False):
    """[B,2] (y, x) draws.  Reference order: randint(y) then randint(x) per sample (damage.py:21-22), one device call
    each, results never read back; fast: two host calls for the whole batch."""
    if fast:
        return torch.stack([torch.randint(lo_y, hi_y, (B,)), torch.randint(lo_x, hi_x, (B,))], 1).pin_memory().to(device, non_blocking=True)
    vals = []
    for _ in range(B):
        vals.append(torch.randint(lo_y, hi_y, (1,), device=device))
        vals.append(torch.randint(lo_x, hi_x, (1,), device=device))
    return torch.cat(vals).view(B, 2)


# ---- mask builders: descriptor -> Damage --------------------------------------------------------------------
@torch.no_grad()
def square_mask(state, size, fast=False) -> Damage:
    B, Cc, H, W = state.shape
    pos = _draw_pairs(B, 0, max(1, H - size + 1), 0, max(1, W - size + 1), state.device, fast)
    return _plane(state, DK_SQUARE, size, pos=pos)


@torch.no_grad()
def circle_mask(state, radius, fast=False) -> Damage:
    B, Cc, H, W = state.shape
    pos = _draw_pairs(B, radius, max(radius + 1, H - radius), radius, max(radius + 1, W - radius), state.device, fast)
    return _plane(state, DK_CIRCLE, radius, pos=pos)


@torch.no_grad()
def stripe_mask(state, width, orientation="auto", fast=False) -> Damage:
    B, Cc, H, W = state.shape
    if orientation == "auto":
        orientation = "h" if random.random() < 0.5 else "v"
    n = H if orientation == "h" else W
    s0 = torch.randint(0, max(1, n - width + 1), (1,)) if fast else torch.randint(0, max(1, n - width + 1), (1,), device=state.device)
    pos = torch.cat([s0.to(state.device), torch.zeros(1, dtype=torch.int64, device=state.device)]).view(1, 2)
    return _plane(state, DK_STRIPE_H if orientation == "h" else DK_STRIPE_V, width, pos=pos)


@torch.no_grad()
def alpha_dropout_mask(state, p, alpha_thr=0.1, hard=True):
    _check(state)
    rand = torch.rand_like(state[:, 3:4])                                   # damage.py:60
    if hard:
        return _plane(state, DK_ALPHA_DROP, rand=rand, p=p, alpha_thr=alpha_thr)
    # soft variant (not in the default policy): alpha channel only, still gated by the alive test
    D = _plane(state, DK_ALPHA_DROP, rand=rand, p=p, alpha_thr=alpha_thr)
    return Damage(D.plane, LAYOUT_PLANE_ALPHA)


@torch.no_grad()
def salt_pepper_mask(state, p) -> Damage:
    _check(state)
    return _plane(state, DK_SALTPEPPER, rand=torch.rand_like(state[:, 3:4]), p=p)       # damage.py:71


@torch.no_grad()
def gaussian_mask(state, radius, softness=0.35, fast=False) -> Damage:
    B, Cc, H, W = state.shape
    pos = _draw_pairs(B, radius, max(radius + 1, H - radius), radius, max(radius + 1, W - radius), state.device, fast)
    return _plane(state, DK_GAUSSIAN, radius, pos=pos, softness=softness)


# ---- in-place operators with the reference's names / signatures (damage.py:16-98) -------------------------
@torch.no_grad()
def cutout_square_(state, size):
    if size > 0:
        apply_damage_(state, square_mask(state, size))


@torch.no_grad()
def cutout_circle_(state, radius):
    if radius > 0:
        apply_damage_(state, circle_mask(state, radius))


@torch.no_grad()
def stripe_wipe_(state, width, orientation="auto"):
    if width > 0:
        apply_damage_(state, stripe_mask(state, width, orientation))


@torch.no_grad()
def alpha_dropout_(state, p, alpha_thr=0.1, hard=True):
    if p > 0:
        apply_damage_(state, alpha_dropout_mask(state, p, alpha_thr, hard))


@torch.no_grad()
def salt_pepper_alpha_(state, p):
    if p > 0:
        apply_damage_(state, salt_pepper_mask(state, p))


@torch.no_grad()
def gaussian_hole_(state, radius, softness=0.35):
    if radius > 0:
        apply_damage_(state, gaussian_mask(state, radius, softness))


@torch.no_grad()
def hidden_scramble_(state, sigma=0.2):
    """Additive noise on hidden channels (weight 0 / absent in the default policy; kept for API parity)."""
    B, Cc, H, W = state.shape
    if Cc <= 4 or sigma <= 0:
        return
    noise = torch.randn(B, Cc - 4, H, W, device=state.device) * sigma
    state[:, 4:] = (state[:, 4:] + noise).clamp_(0.0, 1.0)


def _draw_policy(state, dmg_cfg, epoch, fast=False):
    """The batch-level draws of apply_damage_policy_ (damage.py:101-121), in the reference's order: torch.rand(1) gate,
    random.choices(kind), random.randint(size).  Returns (kind, size) or None (no damage this batch)."""
    start_ep = int(dmg_cfg.get("start_epoch", dmg_cfg.get("damage_start_epoch", 100)))
    prob = float(dmg_cfg.get("prob", dmg_cfg.get("damage_prob", 0.0)))
    if epoch < start_ep or prob <= 0:
        return None
    gate = torch.rand(1).item() if fast else torch.rand(1, device=state.device).item()
    if gate > prob:
        return None
    kinds = dmg_cfg.get("kinds", {"square": 1.0})
    names, weights = zip(*kinds.items())
    kind = random.choices(names, weights=weights, k=1)[0]
    size_min = int(dmg_cfg.get("size_min", dmg_cfg.get("damage_patch_size", 8)))
    size_max = int(dmg_cfg.get("size_max", max(size_min, 14)))
    return kind, int(random.randint(size_min, size_max))


def _policy_mask(state, dmg_cfg, kind, size, fast=False):
    """Damage of the drawn (kind, size) with the kind's own geometry draws (damage.py:122-138); None = nothing to apply."""
    if kind == "circle":
        r = size // 2 if size > 1 else 1
        return circle_mask(state, r, fast) if r > 0 else None
    if kind == "stripes":
        w = int(dmg_cfg.get("stripe_width", size))
        return stripe_mask(state, w, "auto", fast) if w > 0 else None
    if kind == "alpha_drop":
        p = float(dmg_cfg.get("alpha_dropout_p", 0.1))
        return alpha_dropout_mask(state, p, float(dmg_cfg.get("alpha_thr", 0.1)), True) if p > 0 else None
    if kind == "saltpepper":
        p = float(dmg_cfg.get("salt_pepper_p", 0.02))
        return salt_pepper_mask(state, p) if p > 0 else None
    if kind == "gaussian":
        return gaussian_mask(state, max(1, size // 2), float(dmg_cfg.get("gaussian_softness", 0.35)), fast)
    return square_mask(state, size, fast) if size > 0 else None      # "square" and the reference's fallback


@torch.no_grad()
def sample_damage_mask(state, dmg_cfg, epoch, fast=False):
    """The policy of apply_damage_policy_ (damage.py:101-138) returning the `Damage` instead of applying it (None = no
    damage this batch).  PURE: `state` is only read (shape, device, alpha for alpha_drop).  The additive `hidden_noise`
    kind (weight 0 / absent in the default policy) has no mask: it yields None here and is applied by
    apply_damage_policy_."""
    drawn = _draw_policy(state, dmg_cfg, epoch, fast)
    if drawn is None or drawn[0] == "hidden_noise":
        return None
    return _policy_mask(state, dmg_cfg, *drawn, fast=fast)


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
        apply_damage_(state, D)
