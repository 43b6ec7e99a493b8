"""CPU oracle for the graph-NCA step / rollout.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain PyTorch on the CPU, the algorithm of the reference's hot path
(`src/modules/{perception,nca,ncagraph,graph_augmentation}.py`, the rollout / loss / optimiser
lines of `src/training/train_graph_augmented_nca.py`, `src/utils/damage.py`).  It exists so that
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline leg can check / time the CUDA
product against something that travels to the GPU box (the reference checkout does not).
Nothing under `graph_neural_cellular_automata_b200/` imports it.

Parity pin: the reference ships no tests or golden vectors, so the oracle is pinned against the
reference's OWN modules imported in the build container (`tests/golden/make_golden.py` writes the
fixtures, `tests/test_oracle_golden.py` checks the oracle against them).

All randomness is explicit input (fire uniforms, chosen offsets, damage geometry) so the same draw
can be fed to the oracle, the reference and the CUDA path.  All functions work in the dtype of
`x` (fp32 for parity, fp64 as the accuracy yard-stick) and are differentiable through autograd.

Parameter dictionaries use the reference's state-dict keys, e.g.
`update_net.0.weight`, `norm.weight`, `graph.msg_proj.weight`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]
Offset = Tuple[int, int]


@dataclass
class StepConfig:
    """Knobs of one CA step (ctor arguments of `NeuralCA` nca.py:18-27 / `NeuralCAGraph` ncagraph.py:29-47)."""

    update_gain: float = 0.1
    alpha_thr: float = 0.1
    use_groupnorm: bool = True
    graph: bool = True                 # False -> classic NeuralCA (nca.py:64-105)
    message_gain: float = 0.5
    hidden_only: bool = True
    alive_to_alive: bool = True
    zero_padded_shift: bool = True     # module default; trainer forces False (train...:132)
    graph_alpha_thr: Optional[float] = None   # GraphAugmentation.alpha_thr (defaults to alpha_thr)
    gn_eps: float = 1e-3               # nca.py:51 / ncagraph.py:68


# ---------------------------------------------------------------------------------------------
# a1  perception  (perception.py:9-26)
# ---------------------------------------------------------------------------------------------
def perception(x: torch.Tensor) -> torch.Tensor:
    """identity ‖ sobel_x ‖ sobel_y of every channel, zero halo, cross-correlation.

    perception.py:9-10 kernels, :16 zero padding, :25 channel regroup to [id(C), sx(C), sy(C)].
    Written with explicit shifted slices (not conv2d) so it is an independent restatement.
    """
    B, C, H, W = x.shape
    xp = F.pad(x, (1, 1, 1, 1))

    def at(i: int, j: int) -> torch.Tensor:      # in[y+i, x+j]
        return xp[:, :, 1 + i:1 + i + H, 1 + j:1 + j + W]

    sx = (at(-1, -1) - at(-1, 1)) + 2.0 * (at(0, -1) - at(0, 1)) + (at(1, -1) - at(1, 1))
    sy = (at(-1, -1) + 2.0 * at(-1, 0) + at(-1, 1)) - (at(1, -1) + 2.0 * at(1, 0) + at(1, 1))
    return torch.cat([x, sx, sy], dim=1)


# ---------------------------------------------------------------------------------------------
# a7  alive mask  (nca.py:55-62, ncagraph.py:85-92, graph_augmentation.py:116-117)
# ---------------------------------------------------------------------------------------------
def alive_mask(x: torch.Tensor, alpha_thr: float) -> torch.Tensor:
    """`maxpool3x3(alpha) > thr` as float [B,1,H,W]; halo is -inf like F.max_pool2d. Non-differentiable."""
    with torch.no_grad():
        a = x[:, 3:4]
        B, _, H, W = a.shape
        ap = F.pad(a, (1, 1, 1, 1), value=float("-inf"))
        m = ap[:, :, 0:H, 0:W]
        for i in range(3):
            for j in range(3):
                m = torch.maximum(m, ap[:, :, i:i + H, j:j + W])
        return (m > alpha_thr).to(x.dtype)


# ---------------------------------------------------------------------------------------------
# a4  offsets and shifts  (graph_augmentation.py:73-102)
# ---------------------------------------------------------------------------------------------
def build_offsets(radius: int) -> List[Offset]:
    """All (dy,dx) with max(|dy|,|dx|)<=r minus the 3x3 block, dy-major (graph_augmentation.py:73-83)."""
    return [(dy, dx)
            for dy in range(-radius, radius + 1)
            for dx in range(-radius, radius + 1)
            if max(abs(dy), abs(dx)) > 1]


def shift_torus(t: torch.Tensor, dy: int, dx: int) -> torch.Tensor:
    """out[y,x] = in[(y-dy)%H, (x-dx)%W]  (graph_augmentation.py:94-97, torch.roll semantics)."""
    H, W = t.shape[-2:]
    ys = (torch.arange(H) - dy) % H
    xs = (torch.arange(W) - dx) % W
    return t[..., ys, :][..., :, xs]


def shift_zero_pad(t: torch.Tensor, dy: int, dx: int) -> torch.Tensor:
    """The reference's `_shift2d_pad` (graph_augmentation.py:85-92) INCLUDING its x-axis no-op:
    it pads `left` columns on the left and then slices from `left`, so dx never moves anything;
    only the dy shift (zero filled) happens:  out[y,x] = in[y-dy, x] if 0<=y-dy<H else 0."""
    H = t.shape[-2]
    out = torch.zeros_like(t)
    if dy >= 0:
        if dy < H:
            out[..., dy:, :] = t[..., :H - dy, :]
    else:
        if -dy < H:
            out[..., :H + dy, :] = t[..., -dy:, :]
    return out


def _conv1x1(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    y = torch.einsum("oc,bchw->bohw", w.reshape(w.shape[0], w.shape[1]), x)
    if b is not None:
        y = y + b.view(1, -1, 1, 1)
    return y


# ---------------------------------------------------------------------------------------------
# a3  graph augmentation  (graph_augmentation.py:104-169)
# ---------------------------------------------------------------------------------------------
def graph_message(x: torch.Tensor, p: Params, chosen: Sequence[Offset], cfg: StepConfig,
                  prefix: str = "graph.", return_attention_map: bool = False):
    """Aggregated mid-range message [B,C,H,W] (+ attention map [B,H,W]).

    Follows the reference op for op: three 1x1 projections (:109-111), pooled query (:114),
    sender alive mask (:116-117), per-offset shifted K / M / mask (:126-133), pooled logits
    (:136-138), max-subtracted softmax over offsets with |scaling|+1e-6 temperature (:150-154),
    weighted sum (:156-158), optional min-max normalised attention map (:160-167).
    """
    B, C, H, W = x.shape
    Q = _conv1x1(x, p[prefix + "query_proj.weight"], p[prefix + "query_proj.bias"])
    K = _conv1x1(x, p[prefix + "key_proj.weight"], p[prefix + "key_proj.bias"])
    M = _conv1x1(x, p[prefix + "msg_proj.weight"], p[prefix + "msg_proj.bias"])
    q_pooled = Q.mean(dim=(2, 3))
    thr = cfg.alpha_thr if cfg.graph_alpha_thr is None else cfg.graph_alpha_thr
    a_send = alive_mask(x, thr) if cfg.alive_to_alive else None
    shift = shift_zero_pad if cfg.zero_padded_shift else shift_torus

    if len(chosen) == 0:
        agg = torch.zeros_like(M)
        if return_attention_map:
            return agg, torch.zeros(B, H, W, dtype=x.dtype)
        return agg

    msgs, logits = [], []
    for (dy, dx) in chosen:
        k_s = shift(K, dy, dx)
        m_s = shift(M, dy, dx)
        if a_send is not None:
            m_s = m_s * shift(a_send, dy, dx)
        logits.append((q_pooled * k_s.mean(dim=(2, 3))).sum(dim=1))
        msgs.append(m_s)
    L = torch.stack(logits, dim=0)                      # [N,B]
    L = L - L.max(dim=0, keepdim=True).values
    denom = p[prefix + "scaling"].abs() + 1e-6
    wt = torch.softmax(L / denom, dim=0).view(len(chosen), B, 1, 1, 1)
    weighted = torch.stack(msgs, dim=0) * wt
    agg = weighted.sum(dim=0)
    if return_attention_map:
        attn = weighted.abs().mean(dim=2).sum(dim=0)
        lo = attn.amin(dim=(1, 2), keepdim=True)
        hi = attn.amax(dim=(1, 2), keepdim=True)
        return agg, (attn - lo) / (hi - lo + 1e-8)
    return agg


def attention_weights(x: torch.Tensor, p: Params, chosen: Sequence[Offset], cfg: StepConfig,
                      prefix: str = "graph.") -> torch.Tensor:
    """Per-sample softmax weights [N,B] only (same arithmetic as graph_message; diagnostics)."""
    Q = _conv1x1(x, p[prefix + "query_proj.weight"], p[prefix + "query_proj.bias"])
    K = _conv1x1(x, p[prefix + "key_proj.weight"], p[prefix + "key_proj.bias"])
    shift = shift_zero_pad if cfg.zero_padded_shift else shift_torus
    qp = Q.mean(dim=(2, 3))
    L = torch.stack([(qp * shift(K, dy, dx).mean(dim=(2, 3))).sum(dim=1) for dy, dx in chosen], 0)
    L = L - L.max(dim=0, keepdim=True).values
    return torch.softmax(L / (p[prefix + "scaling"].abs() + 1e-6), dim=0)


# ---------------------------------------------------------------------------------------------
# a2, a5, a6, a8, a9, a10  the step  (ncagraph.py:106-168, nca.py:64-105)
# ---------------------------------------------------------------------------------------------
def nca_step(x: torch.Tensor, p: Params, cfg: StepConfig, fire_rate: float = 1.0,
             fire_u: Optional[torch.Tensor] = None, chosen: Sequence[Offset] = (),
             return_attention: bool = False, return_aux: bool = False):
    """One CA step.  `fire_u` are the uniforms the reference would have drawn with
    `torch.rand(B,1,H,W)` (needed iff fire_rate < 1.0); `chosen` is the `random.sample` result."""
    B, C, H, W = x.shape
    y = perception(x)                                                        # ncagraph.py:128
    h = torch.relu(_conv1x1(y, p["update_net.0.weight"], p["update_net.0.bias"]))
    dx = _conv1x1(h, p["update_net.2.weight"], None)                         # :131
    attn = None
    if cfg.graph:
        if return_attention:
            m, attn = graph_message(x, p, chosen, cfg, return_attention_map=True)   # :134-138
        else:
            m = graph_message(x, p, chosen, cfg)
        if cfg.hidden_only and C >= 4:                                       # :94-104
            m = torch.cat([torch.zeros_like(m[:, :4]), m[:, 4:]], dim=1)
        dx = dx + torch.tanh(m) * cfg.message_gain                           # :141
    fire = None
    if fire_rate < 1.0:                                                      # :144-146 (`<=`)
        assert fire_u is not None, "fire_rate < 1 needs the recorded uniforms"
        fire = (fire_u <= fire_rate).to(x.dtype)
        dx = dx * fire
    pre = alive_mask(x, cfg.alpha_thr)                                       # :149-150
    u = dx * pre
    if cfg.use_groupnorm:                                                    # :153 GroupNorm(1,C)
        mu = u.mean(dim=(1, 2, 3), keepdim=True)
        var = u.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
        z = (u - mu) / torch.sqrt(var + cfg.gn_eps)
        z = z * p["norm.weight"].view(1, -1, 1, 1) + p["norm.bias"].view(1, -1, 1, 1)
    else:
        z = u
    xt = x + torch.tanh(z) * cfg.update_gain                                 # :154-155
    post = alive_mask(xt, cfg.alpha_thr)                                     # :158
    gate = torch.ones_like(xt)
    gate[:, 3:4] = post                                                      # :159-166 alpha only
    out = xt * gate
    if return_aux:
        return out, {"u": u, "pre": pre, "post": post, "fire": fire, "attn": attn}
    return (out, attn) if return_attention else out


def rollout(x0: torch.Tensor, p: Params, cfg: StepConfig, fire_rates: Sequence[float],
            fire_us: Optional[Sequence[Optional[torch.Tensor]]] = None,
            chosens: Optional[Sequence[Sequence[Offset]]] = None,
            message_gains: Optional[Sequence[float]] = None,
            steps: Optional[torch.Tensor] = None) -> torch.Tensor:
    """T sequential steps with the trainer's per-sample step counts (train...:302-324):
    at step t only samples with steps[b] > t advance; `fire_us[t]` has one row per ACTIVE sample
    in index order (that is what `torch.rand(n_active,1,H,W)` yields in the reference)."""
    T = len(fire_rates)
    x = x0
    for t in range(T):
        c = StepConfig(**{**cfg.__dict__})
        if message_gains is not None:
            c.message_gain = float(message_gains[t])
        ch = chosens[t] if chosens is not None else ()
        fu = fire_us[t] if fire_us is not None else None
        if steps is None:
            x = nca_step(x, p, c, fire_rates[t], fu, ch)
        else:
            mask = steps > t
            if not bool(mask.any()):
                continue
            new = nca_step(x[mask], p, c, fire_rates[t], fu, ch)
            x = x.clone()
            x[mask] = new
    return x


# ---------------------------------------------------------------------------------------------
# a12  loss  (train_graph_augmented_nca.py:52-61)   /  classic masked loss (train_intermediate_loss.py:37-51)
# ---------------------------------------------------------------------------------------------
def loss_premult_rgba(pred4: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """Per-sample MSE of [rgb*a, a] vs a premultiplied target [B,4,H,W] -> [B]."""
    rgba = torch.cat([pred4[:, :3] * pred4[:, 3:4], pred4[:, 3:4]], dim=1)
    return ((rgba - target) ** 2).mean(dim=(1, 2, 3))


def masked_loss(pred4: torch.Tensor, target: torch.Tensor, alpha_thr: float = 0.2,
                lam_area: float = 5e-5) -> torch.Tensor:
    tm = (target[:, 3:4] > alpha_thr).to(pred4.dtype)
    per = (((pred4 - target) ** 2) * tm).sum(dim=(1, 2, 3)) / (tm.sum(dim=(1, 2, 3)) + 1e-8)
    return per + lam_area * pred4[:, 3:4].mean(dim=(1, 2, 3))


# ---------------------------------------------------------------------------------------------
# a13  gradient post-processing + Adam  (train...:370-375, torch.optim.Adam defaults, L2 wd)
# ---------------------------------------------------------------------------------------------
def normalise_grads_(grads: Dict[str, Optional[torch.Tensor]]) -> None:
    """Per-parameter-tensor g /= (||g||_2 + 1e-8), skipping tensors without a gradient."""
    for g in grads.values():
        if g is not None:
            g.div_(g.norm() + 1e-8)


def adam_step_(param: torch.Tensor, grad: torch.Tensor, m: torch.Tensor, v: torch.Tensor, step: int,
               lr: float, weight_decay: float = 0.0, beta1: float = 0.9, beta2: float = 0.999,
               eps: float = 1e-8) -> None:
    """torch.optim.Adam single-tensor update (coupled L2 weight decay, no amsgrad); `step` is 1-based."""
    g = grad + weight_decay * param if weight_decay != 0.0 else grad
    m.mul_(beta1).add_(g, alpha=1.0 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    param.addcdiv_(m, denom, value=-lr / bc1)


# ---------------------------------------------------------------------------------------------
# a15  damage as multiplicative masks  (damage.py:16-98)
# ---------------------------------------------------------------------------------------------
def damage_mask(kind: str, B: int, C: int, H: int, W: int, *, size: int = 0,
                pos: Optional[Sequence[Tuple[int, int]]] = None, orientation: str = "h",
                rand: Optional[torch.Tensor] = None, p: float = 0.0, alpha: Optional[torch.Tensor] = None,
                alpha_thr: float = 0.1, softness: float = 0.35, dtype=torch.float32) -> torch.Tensor:
    """Return D [B,C,H,W] such that the reference's in-place damage equals `state * D`.

    kind / geometry follow apply_damage_policy_ (damage.py:101-138):
      square  (:16-24)  pos[b]=(y,x) top-left, side `size`
      circle  (:27-37)  pos[b]=(cy,cx), radius `size`  (caller passes size//2)
      stripes (:40-51)  one band for the whole batch: pos[0]=(y0 or x0, _), width `size`
      alpha_drop hard (:54-66)  rand[B,1,H,W] < p on alive alpha -> zero all channels
      saltpepper (:69-73)  rand < p -> zero alpha only
      gaussian (:87-98)  pos[b]=(cy,cx), radius `size`: all channels * clamp(1-exp(-r2/(2(R*soft)^2)),0,1)
    """
    D = torch.ones(B, C, H, W, dtype=dtype)
    yy = torch.arange(H).view(H, 1).to(dtype)
    xx = torch.arange(W).view(1, W).to(dtype)
    if kind == "square":
        for b in range(B):
            y, x = pos[b]
            D[b, :, y:y + size, x:x + size] = 0.0
    elif kind == "circle":
        for b in range(B):
            cy, cx = pos[b]
            D[b, :, ((yy - cy) ** 2 + (xx - cx) ** 2) <= size ** 2] = 0.0
    elif kind == "stripes":
        s0 = pos[0][0]
        if orientation == "h":
            D[:, :, s0:s0 + size, :] = 0.0
        else:
            D[:, :, :, s0:s0 + size] = 0.0
    elif kind == "alpha_drop":
        drop = (rand < p).to(dtype) * (alpha > alpha_thr).to(dtype)
        D = D * (1.0 - drop)
    elif kind == "saltpepper":
        D[:, 3:4] = 1.0 - (rand < p).to(dtype)
    elif kind == "gaussian":
        for b in range(B):
            cy, cx = pos[b]
            r2 = (yy - cy) ** 2 + (xx - cx) ** 2
            m = torch.exp(-(r2 / (2.0 * (size * max(1e-6, softness)) ** 2)))
            D[b] = D[b] * (1.0 - m).clamp(0.0, 1.0)
    else:
        raise ValueError(kind)
    return D


# ---------------------------------------------------------------------------------------------
# seeds (nca_init.py:4-6, train_graph_augmented_nca.py:108-114)
# ---------------------------------------------------------------------------------------------
def make_seed(n_channels: int, img_size: int, batch_size: int = 1, dtype=torch.float32) -> torch.Tensor:
    g = torch.zeros(batch_size, n_channels, img_size, img_size, dtype=dtype)
    g[:, 3:, img_size // 2, img_size // 2] = 1.0
    return g


def trainer_seed(n_channels: int, img_size: int, hidden_noise: torch.Tensor) -> torch.Tensor:
    """alpha=1, hidden = 0.01*N(0,1) at the centre cell; `hidden_noise` is the [B,C-4] randn draw."""
    B = hidden_noise.shape[0]
    g = torch.zeros(B, n_channels, img_size, img_size, dtype=hidden_noise.dtype)
    g[:, 3, img_size // 2, img_size // 2] = 1.0
    if n_channels > 4:
        g[:, 4:, img_size // 2, img_size // 2] = 0.01 * hidden_noise
    return g
