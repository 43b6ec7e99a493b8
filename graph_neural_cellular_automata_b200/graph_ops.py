"""Stand-alone GraphAugmentation.forward on the GPU (gnca_graph_fwd/_bwd)."""
from __future__ import annotations

import ctypes as C
from typing import Sequence, Tuple

import torch

from . import _lib
from . import functional as GF


def _graph_desc(g):
    return GF.make_model_desc(g.n_channels, 4, g.d_model, graph=True, torus=not g.zero_padded_shift, hidden_only=False,
                              alive_to_alive=g.alive_to_alive, groupnorm=False, update_gain=0.0,
                              alpha_thr=float(g.alpha_thr), graph_alpha_thr=float(g.alpha_thr))


def _graph_flat(g, desc):
    """Canonical flat buffer with only the graph slots filled (the MLP slots are unused by gnca_graph_*)."""
    lay = GF.param_layout(desc)
    dev = g.msg_proj.weight.device
    flat = torch.zeros(lay.total, dtype=torch.float32, device=dev)
    for off, p in ((lay.wm, g.msg_proj.weight), (lay.bm, g.msg_proj.bias), (lay.wq, g.query_proj.weight),
                   (lay.bq, g.query_proj.bias), (lay.wk, g.key_proj.weight), (lay.bk, g.key_proj.bias),
                   (lay.scaling, g.scaling)):
        flat[off:off + p.numel()] = p.detach().reshape(-1)
    return flat, lay


class _GraphFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, cfg, wm, bm, wq, bq, wk, bk, scaling):
        g = cfg["module"]
        x = GF._require_cuda_f32(x, "x")
        B, Cc, H, W = x.shape
        desc = _graph_desc(g)
        flat, lay = _graph_flat(g, desc)
        packed = GF.pack_weights(desc, flat)
        arr, k = GF._offsets_array(cfg["chosen"])
        lib = _lib.load()
        msg = torch.empty_like(x)
        attn = torch.empty(B, H, W, dtype=torch.float32, device=x.device) if cfg["want_attn"] else None
        nbytes = lib.gnca_step_workspace_bytes(C.byref(desc), B, H, W)
        ws = GF._WS.get(x.device, nbytes)
        _lib.check(lib.gnca_graph_fwd(C.byref(desc), GF._ptr(packed), B, H, W, GF._ptr(x), arr, k, GF._ptr(msg),
                                      GF._ptr(attn), GF._ptr(ws), ws.numel(), GF._stream()), "gnca_graph_fwd")
        ctx.cfg, ctx.desc, ctx.packed, ctx.arr, ctx.lay = cfg, desc, packed, (arr, k), lay
        ctx.save_for_backward(x)
        if attn is not None:
            ctx.mark_non_differentiable(attn)
            return msg, attn
        return msg

    @staticmethod
    def backward(ctx, gmsg, *unused):
        (x,) = ctx.saved_tensors
        B, Cc, H, W = x.shape
        desc, lay = ctx.desc, ctx.lay
        gmsg = GF._require_cuda_f32(gmsg, "grad")
        lib = _lib.load()
        gx = torch.empty_like(x)
        gflat = torch.zeros(lay.total, dtype=torch.float32, device=x.device)
        arr, k = ctx.arr
        nbytes = lib.gnca_step_workspace_bytes(C.byref(desc), B, H, W)
        ws = GF._WS.get(x.device, nbytes)
        _lib.check(lib.gnca_graph_bwd(C.byref(desc), GF._ptr(ctx.packed), B, H, W, GF._ptr(x), arr, k, GF._ptr(gmsg),
                                      GF._ptr(gx), GF._ptr(gflat), GF._ptr(ws), ws.numel(), GF._stream()),
                   "gnca_graph_bwd")
        g = ctx.cfg["module"]
        d, Cn = g.d_model, g.n_channels
        seg = lambda off, n, shape: gflat[off:off + n].view(shape)
        return (gx, None, seg(lay.wm, Cn * Cn, (Cn, Cn, 1, 1)), seg(lay.bm, Cn, (Cn,)),
                seg(lay.wq, d * Cn, (d, Cn, 1, 1)), seg(lay.bq, d, (d,)), seg(lay.wk, d * Cn, (d, Cn, 1, 1)),
                seg(lay.bk, d, (d,)), gflat[lay.scaling].view(()))


def graph_message(g, x: torch.Tensor, chosen: Sequence[Tuple[int, int]], return_attention_map: bool = False):
    """GraphAugmentation.forward (graph_augmentation.py:104-169) for already drawn offsets."""
    if not x.is_cuda:
        raise RuntimeError(f"GraphAugmentation: input on {x.device}; CUDA (sm_100a) only, no CPU fallback")
    cfg = {"module": g, "chosen": tuple(chosen), "want_attn": bool(return_attention_map)}
    return _GraphFn.apply(x, cfg, g.msg_proj.weight, g.msg_proj.bias, g.query_proj.weight, g.query_proj.bias,
                          g.key_proj.weight, g.key_proj.bias, g.scaling)
