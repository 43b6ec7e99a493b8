"""Baseline legs of bench.py: the REFERENCE's own PyTorch path, timed on the box's host cores (cpu_baseline, the
`--impl reference` arm) and in PyTorch eager on the B200 (gpu_eager: the like-for-like GPU comparator BASELINE.md
section 3 / SURVEY 8d ask for).

Which code runs: the unmodified reference modules staged in `oracle/_ref/` by `oracle/build_ref.py` (kind "reference");
if that directory was never staged, the restatement in `oracle/nca_oracle.py` behind the same call shape (kind "port").
The training step is the loop body of /root/reference/src/training/train_graph_augmented_nca.py:289-391 restated here
around those modules (the trainer script itself is a monolithic `main()` that needs matplotlib / tensorboard / a PNG and
cannot be imported): pool sample, per-sample step counts, per-step fire rate and message gating,
`state[mask] = model(state[mask], fire_rate=fr)`, premultiplied loss, backward, per-tensor gradient normalisation, Adam,
worst-k reseed, pool replace.  This file is measurement infrastructure: the product never imports it.
"""
from __future__ import annotations

import os
import random
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def _weights(name):
    return {k: torch.from_numpy(v) for k, v in np.load(os.path.join(GOLDEN, name)).items()}


class _PortGraph(torch.nn.Module):
    """oracle/nca_oracle.py behind the reference's module call shape (fallback when oracle/_ref is absent)."""

    def __init__(self, params):
        super().__init__()
        from oracle import nca_oracle as O
        self.O = O
        self.p = torch.nn.ParameterDict({k.replace(".", "__"): torch.nn.Parameter(v.clone(), requires_grad="perception" not in k)
                                         for k, v in params.items() if v.is_floating_point()})
        self.message_gain = 0.25
        self.offsets = O.build_offsets(4)

    def forward(self, x, fire_rate=1.0):
        O = self.O
        p = {k.replace("__", "."): v for k, v in self.p.items()}
        cfg = O.StepConfig(update_gain=0.05, alpha_thr=0.12, graph=True, message_gain=float(self.message_gain),
                           hidden_only=True, zero_padded_shift=False)
        chosen = random.sample(self.offsets, 8)
        fu = torch.rand(x.shape[0], 1, x.shape[2], x.shape[3], device=x.device) if fire_rate < 1.0 else None
        return O.nca_step(x, p, cfg, fire_rate, fu, chosen)


def make_reference_graph(device="cpu"):
    """(model, kind): the reference NeuralCAGraph with the trained gecko weights (torus shift, as the trainer forces)."""
    sys.path.insert(0, ROOT)
    from oracle.build_ref import import_reference
    ref = import_reference()
    w = _weights("weights_graph_ep960.npz")
    if ref is not None:
        _, NeuralCAGraph, _ = ref
        m = NeuralCAGraph(16, update_hidden=128, img_size=40, update_gain=0.05, alpha_thr=0.12, use_groupnorm=True,
                          message_gain=0.25, hidden_only=True, graph_d_model=16, graph_attention_radius=4,
                          graph_num_neighbors=8, graph_gating_hidden=32, graph_zero_padded_shift=False)
        missing, unexpected = m.load_state_dict(w, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        return m.to(device), "reference"
    return _PortGraph(w).to(device), "port"


def make_seed(B, device):
    g = torch.zeros(B, 16, 40, 40, device=device)
    g[:, 3:, 20, 20] = 1.0                                   # utils/nca_init.py:4-6
    return g


def _sync(device):
    if torch.device(device).type == "cuda":
        torch.cuda.synchronize()


def forward_rollout(model, B, T, fire_rate, device, message_every=1):
    """BASELINE configs[1]: T forward calls from the seed, no grad (the loop of the reference's test scripts)."""
    x = make_seed(B, device)
    base = 0.25
    with torch.no_grad():
        for t in range(T):
            model.message_gain = base if (message_every <= 1 or t % message_every == 0) else 0.0
            x = model(x, fire_rate=fire_rate)
    model.message_gain = base
    return x


def time_forward(B, T, fire_rate, device, reps=2, threads=None):
    """cell-updates/s of the reference forward rollout on `device` (best of `reps` after a short warm-up)."""
    if torch.device(device).type == "cpu":
        threads = threads or os.cpu_count() or 1
        torch.set_num_threads(threads)
    model, kind = make_reference_graph(device)
    random.seed(42); torch.manual_seed(42)
    forward_rollout(model, B, min(T, 8), fire_rate, device)
    _sync(device)
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        forward_rollout(model, B, T, fire_rate, device)
        _sync(device)
        best = min(best, time.perf_counter() - t0)
    return {"value": B * T * 1600 / best, "unit": "cell-updates/s", "kind": kind,
            "cores": threads if torch.device(device).type == "cpu" else None, "seconds": best}


class ReferenceTrainer:
    """train_graph_augmented_nca.py:289-391 around the reference modules (short regime by default)."""

    def __init__(self, device, batch=32, pool_size=1024, steps_range=(48, 80), seed=1234):
        random.seed(seed); torch.manual_seed(seed)
        self.device, self.B = device, batch
        self.model, self.kind = make_reference_graph(device)
        self.target = torch.from_numpy(np.load(os.path.join(GOLDEN, "target_gecko_surrogate.npy"))).to(device)
        self.params = [p for p in self.model.parameters() if p.requires_grad]
        self.opt = torch.optim.Adam(self.params, lr=2e-4, weight_decay=1e-5)          # train...:143-147
        self.lo, self.hi = steps_range
        self.pool = [self.seed_fn(1)[0] for _ in range(pool_size)]                     # pool.py:12-18
        # mixed ages: an eighth of the pool pre-rolled 16..48 steps, no grad
        with torch.no_grad():
            for i0 in range(0, pool_size // 8, 64):
                x = torch.stack(self.pool[i0:i0 + 64])
                for _ in range(random.randint(16, 48)):
                    x = self.model(x, fire_rate=0.6)
                for j in range(x.shape[0]):
                    self.pool[i0 + j] = x[j]

    def seed_fn(self, n):                                                             # train...:108-114
        g = torch.zeros(n, 16, 40, 40, device=self.device)
        g[:, 3:4, 20, 20] = 1.0
        g[:, 4:, 20, 20] = 0.01 * torch.randn_like(g[:, 4:, 20, 20])
        return g

    def step(self):
        B, dev, model = self.B, self.device, self.model
        idx = random.sample(range(len(self.pool)), B)                                 # pool.py:21-29
        state = torch.stack([self.pool[i].clone() for i in idx])
        nca_steps = torch.randint(self.lo, self.hi + 1, (B,), device=dev)             # train...:297-301
        max_steps = int(nca_steps.max().item())
        for t in range(max_steps):                                                    # train...:305-324
            mask = nca_steps > t
            if not mask.any():
                break
            fr = float(torch.empty(1, device=dev).uniform_(0.5, 0.9).item())
            model.message_gain = 0.25 if t % 3 == 0 else 0.0
            new = model(state[mask], fire_rate=fr)
            state = state.clone()
            state[mask] = new
        model.message_gain = 0.25
        pred = state[:, :4]
        rgba = torch.cat([pred[:, :3] * pred[:, 3:4], pred[:, 3:4]], dim=1)           # train...:52-61
        per = torch.nn.functional.mse_loss(rgba, self.target.unsqueeze(0).expand(B, -1, -1, -1), reduction="none").mean(dim=(1, 2, 3))
        loss = per.mean()
        self.opt.zero_grad(set_to_none=True)
        loss.backward()                                                               # train...:367-368
        for p in self.params:                                                         # train...:370-373
            if p.grad is not None:
                p.grad /= (p.grad.norm() + 1e-8)
        self.opt.step()
        with torch.no_grad():                                                         # train...:378-391
            worst = torch.topk(per, int(0.10 * B)).indices
            state = state.detach()
            state[worst] = self.seed_fn(len(worst))
            for j, i in enumerate(idx):
                self.pool[i] = state[j]
        return float(loss.item()), int(nca_steps.sum().item()) * 1600


def time_train_step(device, batch=32, steps=1, warmup=0, threads=None):
    """cell-updates/s (fwd+bwd) of the reference training iteration, short regime (steps ~ randint(48,80))."""
    if torch.device(device).type == "cpu":
        threads = threads or os.cpu_count() or 1
        torch.set_num_threads(threads)
    tr = ReferenceTrainer(device, batch=batch)
    for _ in range(warmup):
        tr.step()
    _sync(device)
    t0 = time.perf_counter()
    updates = 0
    for _ in range(steps):
        updates += tr.step()[1]
    _sync(device)
    dt = time.perf_counter() - t0
    return {"value": updates / dt, "unit": "cell-updates/s", "kind": tr.kind,
            "cores": threads if torch.device(device).type == "cpu" else None, "seconds_per_step": dt / steps,
            "sample": f"{steps} training iteration(s), B={batch}, steps~randint(48,80), fr~U(0.5,0.9), msg_every=3, "
                      f"loss + backward + grad-normalise + Adam + worst-k reseed"}
