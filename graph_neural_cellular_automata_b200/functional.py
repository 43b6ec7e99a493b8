"""Host side of the graph-NCA operators: thin autograd.Function wrappers over the C ABI of libgnca.so.

Every operator here launches hand-written sm_100a kernels on `torch.cuda.current_stream()`; PyTorch only
provides device memory, streams and autograd plumbing.  Inputs must be CUDA fp32 tensors -- there is no CPU
or eager-PyTorch fallback (north star: "no CPU fallback"); anything else raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import GncaModel, GncaLayout

Offset = Tuple[int, int]

# canonical order of the flat parameter / gradient buffer (gnca_param_layout)
CLASSIC_KEYS = ("update_net.0.weight", "update_net.0.bias", "update_net.2.weight", "norm.weight", "norm.bias")
GRAPH_KEYS = CLASSIC_KEYS + ("graph.msg_proj.weight", "graph.msg_proj.bias", "graph.query_proj.weight",
                             "graph.query_proj.bias", "graph.key_proj.weight", "graph.key_proj.bias", "graph.scaling")
_LAYOUT_FIELDS = ("w1", "b1", "w2", "gamma", "beta", "wm", "bm", "wq", "bq", "wk", "bk", "scaling")


def _require_cuda_f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} is on {t.device}: the graph-NCA operators are CUDA (sm_100a) only and "
                           "deliberately have no CPU fallback")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} has dtype {t.dtype}; the kernels compute in fp32 only")
    return t.contiguous()


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def make_model_desc(C_: int, hidden: int, d_model: int, *, graph: bool, torus: bool, hidden_only: bool,
                    alive_to_alive: bool, groupnorm: bool, update_gain: float, alpha_thr: float,
                    graph_alpha_thr: float, gn_eps: float = 1e-3) -> GncaModel:
    flags = 0
    if graph:
        flags |= _lib.GNCA_F_GRAPH
        if torus:
            flags |= _lib.GNCA_F_TORUS
        if hidden_only:
            flags |= _lib.GNCA_F_HIDDEN_ONLY
        if alive_to_alive:
            flags |= _lib.GNCA_F_ALIVE_TO_ALIVE
    if groupnorm:
        flags |= _lib.GNCA_F_GROUPNORM
    return GncaModel(C_, hidden, d_model if graph else 0, flags, float(update_gain), float(alpha_thr),
                     float(graph_alpha_thr), float(gn_eps))


def param_layout(desc: GncaModel) -> GncaLayout:
    lay = GncaLayout()
    _lib.check(_lib.load().gnca_param_layout(C.byref(desc), C.byref(lay)), "gnca_param_layout")
    return lay


def segment_offsets(desc: GncaModel) -> Sequence[int]:
    """Start offsets of each canonical tensor in the flat buffer + the total (len = n_tensors + 1)."""
    lay = param_layout(desc)
    offs = [getattr(lay, f) for f in _LAYOUT_FIELDS if getattr(lay, f) >= 0]
    return offs + [lay.total]


class _Workspace:
    """Grow-only per-device scratch buffer handed to the C ABI."""

    def __init__(self):
        self.buf: Dict[torch.device, torch.Tensor] = {}

    def get(self, device: torch.device, nbytes: int) -> torch.Tensor:
        b = self.buf.get(device)
        if b is None or b.numel() < nbytes:
            b = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=device)
            self.buf[device] = b
        return b


_WS = _Workspace()


def pack_weights(desc: GncaModel, flat_params: torch.Tensor) -> torch.Tensor:
    lay = param_layout(desc)
    if flat_params.numel() != lay.total:
        raise RuntimeError(f"flat parameter buffer has {flat_params.numel()} floats, layout needs {lay.total}")
    packed = torch.empty(lay.packed_total, dtype=torch.float32, device=flat_params.device)
    _lib.check(_lib.load().gnca_pack_weights(C.byref(desc), _ptr(flat_params), _ptr(packed), _stream()),
               "gnca_pack_weights")
    return packed


def _offsets_array(chosen: Sequence[Offset]):
    k = len(chosen)
    if k > _lib.GNCA_MAX_K:
        raise RuntimeError(f"{k} offsets per step exceed GNCA_MAX_K={_lib.GNCA_MAX_K}")
    arr = (C.c_int32 * max(1, 2 * k))()
    for i, (dy, dx) in enumerate(chosen):
        arr[2 * i], arr[2 * i + 1] = int(dy), int(dx)
    return arr, k


# ------------------------------------------------------------------------------------------------------
# perception.py:21-26
# ------------------------------------------------------------------------------------------------------
class _PerceptionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _require_cuda_f32(x, "x")
        B, Cc, H, W = x.shape
        y = torch.empty(B, 3 * Cc, H, W, dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().gnca_perception_fwd(B, Cc, H, W, _ptr(x), _ptr(y), _stream()), "gnca_perception_fwd")
        ctx.shape = (B, Cc, H, W)
        return y

    @staticmethod
    def backward(ctx, gy):
        B, Cc, H, W = ctx.shape
        gy = _require_cuda_f32(gy, "grad")
        gx = torch.empty(B, Cc, H, W, dtype=torch.float32, device=gy.device)
        _lib.check(_lib.load().gnca_perception_bwd(B, Cc, H, W, _ptr(gy), _ptr(gx), _stream()), "gnca_perception_bwd")
        return gx


def perception(x: torch.Tensor) -> torch.Tensor:
    """[B,C,H,W] -> [B,3C,H,W] identity|sobel_x|sobel_y (zero halo); differentiable w.r.t. x."""
    return _PerceptionFn.apply(x)


def alive_mask(x: torch.Tensor, alpha_thr: float) -> torch.Tensor:
    """nca.py:55-62 -- (maxpool3x3(alpha) > thr) as float [B,1,H,W]; never differentiable."""
    x = _require_cuda_f32(x.detach(), "x")
    B, Cc, H, W = x.shape
    if Cc < 4:
        raise RuntimeError("alive mask needs an alpha channel (C >= 4)")
    m = torch.empty(B, 1, H, W, dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().gnca_alive_mask(B, Cc, H, W, _ptr(x), float(alpha_thr), _ptr(m), _stream()),
               "gnca_alive_mask")
    return m


# ------------------------------------------------------------------------------------------------------
# one CA step (nca.py:64-105 / ncagraph.py:106-168)
# ------------------------------------------------------------------------------------------------------
class _StepFn(torch.autograd.Function):
    """x' = step(x).  Non-tensor context travels in `cfg`; parameters arrive as separate tensors in canonical
    order so autograd routes the flat gradient buffer back to each nn.Parameter as a view."""

    @staticmethod
    def forward(ctx, x, fire_u, cfg, *params):
        desc: GncaModel = cfg["desc"]
        x = _require_cuda_f32(x, "x")
        B, Cc, H, W = x.shape
        if Cc != desc.C:
            raise RuntimeError(f"x has {Cc} channels, model has {desc.C}")
        fire_rate = float(cfg["fire_rate"])
        if fire_rate < 1.0:
            fire_u = _require_cuda_f32(fire_u, "fire_u")
            if fire_u.numel() != B * H * W:
                raise RuntimeError("fire_u must hold B*H*W uniforms")
        else:
            fire_u = None
        packed = cfg["packed"]
        chosen = cfg["chosen"]
        arr, k = _offsets_array(chosen)
        lib = _lib.load()
        out = torch.empty_like(x)
        u = torch.empty_like(x)
        stats = torch.empty(B, 2, dtype=torch.float32, device=x.device)
        attn = torch.empty(B, H, W, dtype=torch.float32, device=x.device) if cfg.get("want_attn") else None
        nbytes = lib.gnca_step_workspace_bytes(C.byref(desc), B, H, W)
        ws = _WS.get(x.device, nbytes)
        _lib.check(lib.gnca_step_fwd(C.byref(desc), _ptr(packed), B, H, W, _ptr(x), _ptr(out), _ptr(fire_u),
                                     fire_rate, arr, k, float(cfg["message_gain"]), _ptr(u), _ptr(stats), _ptr(attn),
                                     _ptr(ws), ws.numel(), _stream()), "gnca_step_fwd")
        ctx.cfg = cfg
        ctx.chosen_arr = (arr, k)
        ctx.n_params = len(params)
        ctx.param_shapes = [p.shape for p in params]
        ctx.save_for_backward(x, fire_u if fire_u is not None else x.new_empty(0), u, stats)
        if attn is not None:
            ctx.mark_non_differentiable(attn)
            return out, attn
        return out

    @staticmethod
    def backward(ctx, gout, *unused):
        cfg = ctx.cfg
        desc: GncaModel = cfg["desc"]
        x, fire_u, u, stats = ctx.saved_tensors
        B, Cc, H, W = x.shape
        gout = _require_cuda_f32(gout, "grad_output")
        lib = _lib.load()
        lay = param_layout(desc)
        gx = torch.empty_like(x)
        gflat = torch.zeros(lay.total, dtype=torch.float32, device=x.device)
        arr, k = ctx.chosen_arr
        nbytes = lib.gnca_step_workspace_bytes(C.byref(desc), B, H, W)
        ws = _WS.get(x.device, nbytes)
        fu = fire_u if fire_u.numel() else None
        _lib.check(lib.gnca_step_bwd(C.byref(desc), _ptr(cfg["packed"]), B, H, W, _ptr(x), _ptr(fu),
                                     float(cfg["fire_rate"]), arr, k, float(cfg["message_gain"]), _ptr(u),
                                     _ptr(stats), _ptr(gout), _ptr(gx), _ptr(gflat), _ptr(ws), ws.numel(),
                                     _stream()), "gnca_step_bwd")
        offs = segment_offsets(desc)
        grads = [gflat[offs[i]:offs[i + 1]].view(ctx.param_shapes[i]) for i in range(ctx.n_params)]
        return (gx, None, None, *grads)


def nca_step(x: torch.Tensor, params: Sequence[torch.Tensor], desc: GncaModel, packed: torch.Tensor, *,
             fire_rate: float = 1.0, fire_u: Optional[torch.Tensor] = None, chosen: Sequence[Offset] = (),
             message_gain: float = 0.0, want_attn: bool = False):
    """Functional form of NeuralCA.forward / NeuralCAGraph.forward.  `params` in canonical order
    (CLASSIC_KEYS / GRAPH_KEYS); `packed` = pack_weights(desc, flat(params))."""
    cfg = {"desc": desc, "packed": packed, "fire_rate": fire_rate, "chosen": tuple(chosen),
           "message_gain": message_gain, "want_attn": want_attn}
    return _StepFn.apply(x, fire_u, cfg, *params)
