"""Checkpoint payloads in the reference's format (train_graph_augmented_nca.py:196-240, 405-416):
{"epoch", "model_state", "optimizer_state" (torch.optim.Adam format), "scheduler_state" (StepLR format), "config",
"param_count", "global_step"} -- a run can resume from a checkpoint the reference wrote and the reference can resume
from one written here."""
from __future__ import annotations

import glob
import os
import re
from typing import Optional, Tuple

import torch


def _epoch_num(path: str) -> int:
    m = re.search(r"nca_epoch(\d+)", os.path.basename(path))
    return int(m.group(1)) if m else -1


def steplr_state(base_lr: float, step_size: int, gamma: float, last_epoch: int) -> dict:
    """torch.optim.lr_scheduler.StepLR.state_dict() after `last_epoch` scheduler steps"""
    lr = base_lr * gamma ** (max(last_epoch, 0) // step_size)
    return {"step_size": step_size, "gamma": gamma, "base_lrs": [base_lr], "last_epoch": last_epoch, "verbose": False,
            "_step_count": last_epoch + 1, "_get_lr_called_within_step": False, "_last_lr": [lr]}


def save_checkpoint(path: str, model, optimizer, epoch: int, global_step: Optional[int] = None, config=None,
                    scheduler_state: Optional[dict] = None) -> dict:
    if hasattr(optimizer, "torch_state_dict"):
        # group["lr"] = what the scheduler holds at this epoch (decayed), group["initial_lr"] = the base rate
        cur = scheduler_state["_last_lr"][0] if scheduler_state and scheduler_state.get("_last_lr") else None
        opt_state = optimizer.torch_state_dict(current_lr=cur)
    else:
        opt_state = optimizer.state_dict()
    payload = {"epoch": int(epoch), "model_state": {k: v.detach().cpu() for k, v in model.state_dict().items()},
               "optimizer_state": opt_state, "scheduler_state": scheduler_state, "config": config,
               "param_count": sum(p.numel() for p in model.parameters() if p.requires_grad),
               "global_step": int(epoch if global_step is None else global_step)}
    torch.save(payload, path)
    return payload


def load_checkpoint(path_or_payload, model, optimizer=None) -> Tuple[dict, list, list]:
    """Returns (payload, missing_keys, unexpected_keys); mirrors the reference's tolerant resume (strict=False, an
    incompatible optimizer state is reported, not fatal)."""
    payload = path_or_payload if isinstance(path_or_payload, dict) else torch.load(path_or_payload, map_location="cpu",
                                                                                   weights_only=False)
    state = payload["model_state"] if "model_state" in payload else payload          # bare state dicts load too
    missing, unexpected = model.load_state_dict(state, strict=False)
    if hasattr(model, "invalidate_packed"):
        model.invalidate_packed()
    if optimizer is not None and isinstance(payload, dict) and payload.get("optimizer_state") is not None:
        if hasattr(optimizer, "load_torch_state_dict"):
            if hasattr(optimizer, "flat"):          # flat parameter copy follows the freshly loaded weights
                optimizer.flat.copy_(torch.cat([p.detach().reshape(-1) for p in model.canonical_params()]))
            optimizer.load_torch_state_dict(payload["optimizer_state"])
        else:
            optimizer.load_state_dict(payload["optimizer_state"])
    return payload, list(missing), list(unexpected)


def pick_resume(ckpt_dir: str):
    """The reference's resume precedence (train...:196-219): the candidate with the largest (epoch, global_step)."""
    cand = []
    latest = os.path.join(ckpt_dir, "nca_latest.pt")
    if os.path.exists(latest):
        cand.append(latest)
    cand += sorted(glob.glob(os.path.join(ckpt_dir, "nca_epoch*_final.pt")))
    cand += sorted(glob.glob(os.path.join(ckpt_dir, "nca_*_last.pt")))
    cand += sorted(glob.glob(os.path.join(ckpt_dir, "nca_crash_ep*.pt")))
    cand += sorted(glob.glob(os.path.join(ckpt_dir, "nca_epoch*.pt")), key=_epoch_num)
    best = (None, None, -1, -1)
    for p in cand:
        try:
            payload = torch.load(p, map_location="cpu", weights_only=False)
            ep = int(payload.get("epoch", -1))
            gs = int(payload.get("global_step", ep))
        except Exception:
            continue
        if ep > best[2] or (ep == best[2] and gs > best[3]):
            best = (p, payload, ep, gs)
    return best[0], best[1]
