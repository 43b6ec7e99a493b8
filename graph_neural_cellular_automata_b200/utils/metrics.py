"""The trainer's per-step quality metrics on the device (train_graph_augmented_nca.py:405-422 computes them with
numpy/skimage on sample 0 after a host copy): pixel-perfection and PSNR on premultiplied RGBA, for the whole batch,
without leaving the GPU.  (SSIM stays host-side, off the hot loop, as in the reference.)"""
from __future__ import annotations

import torch


def premultiply(state: torch.Tensor) -> torch.Tensor:
    pred = state[:, :4]
    return torch.cat([pred[:, :3] * pred[:, 3:4], pred[:, 3:4]], dim=1)


def pixel_perfect(state: torch.Tensor, target: torch.Tensor, eps: float = 0.05) -> torch.Tensor:
    """fraction of pixels whose 4 premultiplied RGBA channels are all within eps of the target; per sample [B]"""
    diff = (premultiply(state) - target.unsqueeze(0)).abs()
    return (diff < eps).all(dim=1).float().mean(dim=(1, 2))


def psnr_rgb(state: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """skimage.metrics.peak_signal_noise_ratio(data_range=1) on the clipped premultiplied RGB; per sample [B]"""
    a = premultiply(state)[:, :3].clamp(0, 1)
    b = target[:3].unsqueeze(0).clamp(0, 1)
    mse = ((a - b) ** 2).mean(dim=(1, 2, 3))
    return 10.0 * torch.log10(1.0 / mse)
