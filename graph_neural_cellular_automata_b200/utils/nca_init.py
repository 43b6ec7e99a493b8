"""Seeds (drop-in for the reference's utils/nca_init.py:4-6 and the trainer's seed_fn, train...:108-114)."""
import torch


def make_seed(n_channels, img_size, batch_size=1, device="cpu"):
    """Single live cell at the centre: alpha and every hidden channel 1.0, RGB 0."""
    c = img_size // 2
    grid = torch.zeros(batch_size, n_channels, img_size, img_size, device=device)
    grid[:, 3:, c, c] = 1.0
    return grid


def trainer_seed(n_channels, img_size, batch_size=1, device="cpu"):
    """The graph trainer's seed_fn: alpha 1, hidden 0.01*N(0,1) at the centre cell (one randn_like draw)."""
    c = img_size // 2
    g = torch.zeros(batch_size, n_channels, img_size, img_size, device=device)
    g[:, 3:4, c, c] = 1.0
    if n_channels > 4:
        g[:, 4:, c, c] = 0.01 * torch.randn_like(g[:, 4:, c, c])
    return g
