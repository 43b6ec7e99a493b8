"""Rollout = T CA steps in ONE library call (gnca_rollout_fwd / gnca_rollout_bwd).

The reference has no rollout function: every script runs a Python `for` over `model(state[mask], fire_rate=fr)`
(train_graph_augmented_nca.py:305-321, test_graph_augmented_regeneration.py:183-194, ...).  The contract of this
extension is "equals T sequential `forward` calls given the same random draws".  A `Schedule` holds those draws
on the device: per-step fire rate, message gain and offsets, per-sample step counts, and the fire uniforms
(either a recorded/pre-drawn tensor or an in-kernel Philox stream).
"""
from __future__ import annotations

import ctypes as C
import random
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib
from . import functional as GF
from ._lib import GncaSchedule

Offset = Tuple[int, int]


@dataclass
class Schedule:
    """Device-resident per-step schedule of a rollout (mirror of `gnca_schedule` in include/gnca.h)."""
    T: int
    k: int
    fire_rate: torch.Tensor                 # [T] f32
    message_gain: torch.Tensor              # [T] f32
    offsets: Optional[torch.Tensor]         # [T,k,2] int8
    steps: Optional[torch.Tensor] = None    # [B] int32
    fire_u: Optional[torch.Tensor] = None   # [T,B,H,W] f32 (row b = sample b)
    philox_seed: int = 0
    philox_offset: int = 0
    damage: Optional[torch.Tensor] = None   # multiplicative mask: [B,C,H,W] (layout 0) or a per-cell plane [B,H,W]
    damage_step: int = 0
    damage_layout: int = 0                  # 0 dense, 1 plane on every channel, 2 plane on alpha only (include/gnca.h)
    total_updates: Optional[int] = None     # sum_b steps_b (host int, for throughput accounting)
    max_offset: int = 0                     # max(|dy|,|dx|) over the offsets (halo depth of the resident kernel)

    def c_struct(self) -> GncaSchedule:
        p = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
        return GncaSchedule(self.T, self.k, p(self.fire_rate), p(self.message_gain), p(self.offsets), p(self.steps),
                            p(self.fire_u), C.c_uint64(self.philox_seed), C.c_uint64(self.philox_offset),
                            p(self.damage), self.damage_step, self.max_offset, int(self.damage_layout), 0)


class _PinnedRing:
    """A few reusable pinned staging buffers (pinning memory per call costs ~100 us); a buffer is reused only after
    the event recorded behind its last H2D copy has completed."""

    def __init__(self, slots: int = 8):
        self.slots, self.buf, self.ev, self.i = slots, [None] * slots, [None] * slots, 0

    def get(self, nbytes: int):
        i = self.i
        self.i = (i + 1) % self.slots
        if self.ev[i] is not None:
            self.ev[i].synchronize()
        if self.buf[i] is None or self.buf[i].numel() < nbytes:
            self.buf[i] = torch.empty(max(nbytes, 1 << 14), dtype=torch.uint8).pin_memory()
        return i, self.buf[i]

    def mark(self, i: int):
        ev = torch.cuda.Event()
        ev.record()
        self.ev[i] = ev


_RING = _PinnedRing()
_TORCH_DTYPE = {np.dtype(np.float32): torch.float32, np.dtype(np.int8): torch.int8, np.dtype(np.int32): torch.int32,
                np.dtype(np.uint8): torch.uint8, np.dtype(np.int64): torch.int64}


def _upload(arrs: Sequence[np.ndarray], device) -> List[torch.Tensor]:
    """One pinned staging buffer, one async H2D copy, typed device views."""
    sizes = [a.nbytes for a in arrs]
    offs, o = [], 0
    for s in sizes:
        offs.append(o)
        o += (s + 15) // 16 * 16
    total = max(o, 16)
    on_cuda = torch.device(device).type == "cuda"
    if on_cuda:
        slot, host = _RING.get(total)
    else:
        host = torch.empty(total, dtype=torch.uint8)
    hv = host.numpy()
    for a, off in zip(arrs, offs):
        hv[off:off + a.nbytes].view(a.dtype)[:] = a.reshape(-1)
    if on_cuda:
        dev = torch.empty(total, dtype=torch.uint8, device=device)
        dev.copy_(host[:total], non_blocking=True)
        _RING.mark(slot)
    else:
        dev = host[:total].clone()
    out = []
    for a, off in zip(arrs, offs):
        t = dev[off:off + a.nbytes].view(_TORCH_DTYPE[a.dtype]).view(a.shape)
        out.append(t)
    return out


def make_schedule(model, B: int, H: int, W: int, T: int, *, fire_rate: Union[float, Sequence[float]] = 1.0,
                  message_gains: Optional[Sequence[float]] = None, message_every: int = 1,
                  offsets: Optional[Sequence[Sequence[Offset]]] = None, steps: Optional[Sequence[int]] = None,
                  fire: str = "philox", fire_u: Optional[torch.Tensor] = None, seed: Optional[int] = None,
                  damage: Optional[torch.Tensor] = None, damage_step: int = 0, device=None) -> Schedule:
    """Build the schedule the way T sequential `forward` calls would consume randomness.

    * offsets: drawn with `random.sample(model.graph.offsets, k)` once per step (graph models), exactly like
      graph_augmentation.py:120-121, unless given.
    * fire="torch": one `torch.rand(B,1,H,W)` per step on the model device (bit-identical stream to T forward
      calls with a fixed fire_rate < 1); fire="philox": in-kernel Philox4x32-10 keyed by `seed` (no HBM traffic,
      statistically equivalent, different numbers); or pass recorded `fire_u` [T,B,H,W].
    * damage: a `utils.damage.Damage` (descriptor evaluated to a per-cell plane, applied in-kernel before step
      `damage_step`) or a dense [B,C,H,W] mask tensor.
    * message_gains: per-step `model.message_gain` (the trainer sets gain or 0 per step, train...:312-319);
      default = model.message_gain on steps with t % message_every == 0, else 0.
    """
    device = device or next(model.parameters()).device
    is_graph = bool(getattr(model, "_is_graph", False))
    fr = np.full(T, float(fire_rate), np.float32) if np.isscalar(fire_rate) else np.asarray(fire_rate, np.float32)
    assert fr.shape == (T,)
    if is_graph:
        if offsets is None:
            off = model.graph.draw_offsets_array(T)          # same python-RNG stream as T forward calls
        elif isinstance(offsets, np.ndarray):
            off = offsets.astype(np.int8, copy=False).reshape(T, -1, 2)
        else:
            k_ = len(offsets[0]) if T > 0 else 0
            flat = [v for st in offsets for o in st for v in o]
            off = np.asarray(flat, dtype=np.int8).reshape(T, k_, 2) if k_ > 0 else np.zeros((T, 0, 2), np.int8)
        k = off.shape[1]
        if message_gains is None:
            g = float(model.message_gain)
            message_gains = [g if (message_every <= 1 or t % message_every == 0) else 0.0 for t in range(T)]
    else:
        k, off = 0, np.zeros((T, 0, 2), np.int8)
        message_gains = [0.0] * T
    gains = np.asarray(message_gains, np.float32)
    arrs = [fr, gains, off if off.size else np.zeros(2, np.int8)]
    if steps is not None:
        st = np.asarray(steps, np.int32)
        assert st.shape == (B,)
        arrs.append(st)
        total = int(np.minimum(st, T).clip(min=0).sum())
    else:
        total = B * T
    dev = _upload(arrs, device)
    damage_layout = 0
    if damage is not None and hasattr(damage, "plane"):              # utils.damage.Damage
        damage_layout, damage = int(damage.layout), GF._require_cuda_f32(damage.plane, "damage plane")
        if tuple(damage.shape) != (B, H, W):
            raise ValueError(f"damage plane has shape {tuple(damage.shape)}, expected {(B, H, W)}")
    elif damage is not None:
        damage = GF._require_cuda_f32(damage, "damage")
    sched = Schedule(T=T, k=k, fire_rate=dev[0], message_gain=dev[1], offsets=dev[2] if off.size else None,
                     steps=dev[3] if steps is not None else None, damage=damage, damage_step=int(damage_step),
                     damage_layout=damage_layout,
                     total_updates=total, max_offset=int(np.abs(off).max()) if off.size else 0)
    needs_fire = bool((fr < 1.0).any())
    if fire_u is not None:
        sched.fire_u = GF._require_cuda_f32(fire_u, "fire_u").view(T, B, H, W)
    elif needs_fire and fire == "torch":
        buf = torch.empty(T, B, 1, H, W, dtype=torch.float32, device=device)
        for t in range(T):
            if fr[t] < 1.0:
                torch.rand(B, 1, H, W, out=buf[t])
        sched.fire_u = buf.view(T, B, H, W)
    elif needs_fire:
        if fire != "philox":
            raise ValueError("fire must be 'torch' or 'philox'")
        sched.philox_seed = int(seed if seed is not None else random.getrandbits(63))
    return sched


class History:
    """What the forward keeps for BPTT.  Two formats:
    * records (`bptt`): the replicated-state kernel's per-active-cell records + bitmaps + statistics -- consumed by the
      cluster-resident backward (gnca_rollout_bwd_bptt); `x` is only filled when the caller asked for the history;
    * dense: x_t for every step and (optionally) the masked pre-norm update of the active cells + the GroupNorm
      statistics, consumed by the streaming backward (gnca_rollout_bwd)."""

    def __init__(self, x, u=None, stats=None, bptt=None, shape=None):
        self.x, self.u, self.stats, self.bptt = x, u, stats, bptt
        self._shape = tuple(shape) if shape is not None else tuple(x.shape)

    @property
    def shape(self):
        return self._shape

    def __getitem__(self, i):
        return self.x[i]


_QUANT_SLACK_BYTES = 8 << 30        # never over-allocate more than this for the sake of a stable block size


def _round_up_steps(T: int, q: int = 64) -> int:
    return T if T <= 16 else -(-T // q) * q


BPTT_MAX_BYTES = 48 << 30      # above this the record buffer is not worth it: dense history + streaming backward


def rollout_fwd_raw(desc, packed, x0: torch.Tensor, sched: Schedule, *, history: bool, impl: int = 0,
                    keep_u: bool = True, keep_x: bool = True):
    """gnca_rollout_fwd without autograd: returns (x_T, History or None).  `keep_x=False`: the caller only needs the
    history for the backward (the resident BPTT path then skips storing x_t)."""
    x0 = GF._require_cuda_f32(x0, "x0")
    B, Cc, H, W = x0.shape
    if Cc != desc.C:
        raise ValueError(f"x0 has {Cc} channels, model has {desc.C}")
    lib = _lib.load()
    T = sched.T
    xT = torch.empty_like(x0)
    nbytes = lib.gnca_rollout_workspace_bytes(C.byref(desc), B, H, W, T)
    ws = GF._WS.get(x0.device, nbytes)
    cs = sched.c_struct()
    if history and impl in (0, 2) and T > 0:
        nb = lib.gnca_bptt_bytes(C.byref(desc), B, H, W, T)
        if 0 < nb <= BPTT_MAX_BYTES:
            # the record buffer is GBs (T*B*H*W*1.4 KB worst case): sized for T rounded up to 64 steps, so that rollouts of
            # slightly different lengths (per-iteration step counts of the trainer) ask the caching allocator for the SAME
            # block instead of a fresh cudaMalloc whenever a new maximum shows up (70 ms for 29 GB in the long regime)
            nb_q = lib.gnca_bptt_bytes(C.byref(desc), B, H, W, _round_up_steps(T))
            if nb <= nb_q <= BPTT_MAX_BYTES and nb_q - nb <= _QUANT_SLACK_BYTES:
                nb = nb_q
            bptt = torch.empty(nb, dtype=torch.uint8, device=x0.device)
            hist = torch.empty(T + 1, B, Cc, H, W, dtype=torch.float32, device=x0.device) if keep_x else None
            rc = lib.gnca_rollout_fwd_bptt(C.byref(desc), GF._ptr(packed), B, H, W, C.byref(cs), GF._ptr(x0), GF._ptr(xT),
                                           GF._ptr(hist), GF._ptr(bptt), nb, GF._ptr(ws), ws.numel(), GF._stream())
            if rc != _lib.GNCA_ERR_UNSUPPORTED:
                _lib.check(rc, "gnca_rollout_fwd_bptt")
                return xT, History(hist, bptt=bptt, shape=(T + 1, B, Cc, H, W))
    hist = uh = sh = None
    if history:
        Tq = _round_up_steps(T)              # same quantised allocation for the dense history (prefix views)
        if (Tq - T) * B * Cc * H * W * 4 > _QUANT_SLACK_BYTES:
            Tq = T
        hist = torch.empty(Tq + 1, B, Cc, H, W, dtype=torch.float32, device=x0.device)[:T + 1]
        if keep_u and T > 0:
            uh = torch.empty(Tq, B, Cc, H, W, dtype=torch.float32, device=x0.device)[:T]
            sh = torch.empty(Tq, B, 2, dtype=torch.float32, device=x0.device)[:T]
    _lib.check(lib.gnca_rollout_fwd(C.byref(desc), GF._ptr(packed), B, H, W, C.byref(cs), GF._ptr(x0), GF._ptr(xT),
                                    GF._ptr(hist), GF._ptr(sh), GF._ptr(uh), GF._ptr(ws), ws.numel(), int(impl),
                                    GF._stream()), "gnca_rollout_fwd")
    return xT, (History(hist, uh, sh) if history else None)


def rollout_bwd_raw(desc, packed, hist, sched: Schedule, gT: torch.Tensor, *, gflat=None, impl: int = 0):
    """gnca_rollout_bwd without autograd: returns (dL/dx_0, flat parameter gradient in canonical layout).
    `gflat` (optional) is accumulated into."""
    if not isinstance(hist, History):
        hist = History(hist)
    _, B, Cc, H, W = hist.shape
    gT = GF._require_cuda_f32(gT, "grad_output")
    lib = _lib.load()
    g0 = torch.empty(B, Cc, H, W, dtype=torch.float32, device=gT.device)
    if gflat is None:
        lay = GF.param_layout(desc)
        gflat = torch.zeros(lay.total, dtype=torch.float32, device=gT.device)
    nbytes = lib.gnca_rollout_workspace_bytes(C.byref(desc), B, H, W, sched.T)
    ws = GF._WS.get(gT.device, nbytes)
    cs = sched.c_struct()
    if hist.bptt is not None:
        _lib.check(lib.gnca_rollout_bwd_bptt(C.byref(desc), GF._ptr(packed), B, H, W, C.byref(cs), GF._ptr(hist.bptt),
                                             hist.bptt.numel(), GF._ptr(gT), GF._ptr(g0), GF._ptr(gflat), GF._ptr(ws),
                                             ws.numel(), GF._stream()), "gnca_rollout_bwd_bptt")
        return g0, gflat
    _lib.check(lib.gnca_rollout_bwd(C.byref(desc), GF._ptr(packed), B, H, W, C.byref(cs), GF._ptr(hist.x),
                                    GF._ptr(hist.stats), GF._ptr(hist.u), GF._ptr(gT), GF._ptr(g0), GF._ptr(gflat),
                                    GF._ptr(ws), ws.numel(), int(impl), GF._stream()), "gnca_rollout_bwd")
    return g0, gflat


class _RolloutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, cfg, *params):
        xT, hist = rollout_fwd_raw(cfg["desc"], cfg["packed"], x0, cfg["schedule"],
                                   history=bool(cfg["need_grad"] or cfg["history"]), impl=cfg["impl"],
                                   keep_x=bool(cfg["history"]))
        ctx.cfg = cfg
        ctx.param_shapes = [p.shape for p in params]
        ctx.hist = hist
        if cfg["history"]:
            ctx.mark_non_differentiable(hist.x)
            return xT, hist.x
        return xT

    @staticmethod
    def backward(ctx, gT, *unused):
        cfg = ctx.cfg
        if ctx.hist is None:
            raise RuntimeError("rollout was run without gradient history")
        g0, gflat = rollout_bwd_raw(cfg["desc"], cfg["packed"], ctx.hist, cfg["schedule"], gT, impl=cfg["impl"])
        offs = GF.segment_offsets(cfg["desc"])
        grads = [gflat[offs[i]:offs[i + 1]].view(ctx.param_shapes[i]) for i in range(len(ctx.param_shapes))]
        ctx.hist = None
        return (g0, None, *grads)


IMPL = {"auto": 0, "streaming": 1, "resident": 2, "banded": 3}


def rollout(model, x0: torch.Tensor, schedule: Schedule, *, return_history: bool = False, impl: str = "auto"):
    """x_T (and optionally the [T+1,B,C,H,W] history) of T steps from x0 under `schedule`.  Differentiable
    w.r.t. x0 and the model parameters (BPTT stores x_t only and recomputes the rest)."""
    if not x0.is_cuda:
        raise RuntimeError(f"rollout: x0 is on {x0.device}; CUDA (sm_100a) only, no CPU fallback")
    ps = model.canonical_params()
    need_grad = torch.is_grad_enabled() and (x0.requires_grad or any(p.requires_grad for p in ps))
    if not need_grad and not return_history:        # inference: no autograd node, no history (same C call underneath)
        return rollout_fwd_raw(model.model_desc(), model.packed_weights(), x0, schedule, history=False, impl=IMPL[impl])[0]
    cfg = {"desc": model.model_desc(), "packed": model.packed_weights(), "schedule": schedule, "need_grad": need_grad,
           "history": bool(return_history), "impl": IMPL[impl]}
    return _RolloutFn.apply(x0, cfg, *ps)
