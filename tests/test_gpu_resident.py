"""GPU tests of the cluster-resident kernels (replicated-state forward gnca_rep.cu, resident BPTT gnca_rep_bwd.cu)
against the streaming per-step kernels, which the other test files pin to the oracle / golden fixtures.
Covers what the golden fixtures do not: damage at a step, ragged step counts, every cluster size (incl. the
overflow paths of the in-smem buffers), other grid shapes, different pre-alive / sender-alive thresholds (the
general mask path), per-step gains / fire rates, GroupNorm off."""
import os
import random

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import nca_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import graph_neural_cellular_automata_b200 as G
    from graph_neural_cellular_automata_b200 import functional as GF
    from graph_neural_cellular_automata_b200.rollout import make_schedule, rollout
    from test_gpu_step import graph_model, classic_model, T32, DEV


def _grown_state(m, B, H, W, steps=24, seed=0):
    """seed states grown for a few steps with the streaming kernels (alive blobs of different ages)"""
    torch.manual_seed(seed); random.seed(seed)
    from graph_neural_cellular_automata_b200.utils.nca_init import make_seed
    x = make_seed(16, H, B, device=DEV) if H == W else None
    if x is None:
        x = torch.zeros(B, 16, H, W, device=DEV)
        x[:, 3:, H // 2, W // 2] = 1.0
    with torch.no_grad():
        s = make_schedule(m, B, H, W, steps, fire_rate=0.6, seed=seed + 1)
        x = rollout(m, x, s, impl="streaming")
    return x


def _sched(m, B, H, W, T, seed, **kw):
    random.seed(seed)
    fu = torch.rand(T, B, H, W, device=DEV, generator=torch.Generator(device=DEV).manual_seed(seed))
    frs = [0.5 + 0.4 * ((7 * t) % 10) / 10 for t in range(T)]
    gains = None
    if getattr(m, "_is_graph", False):
        gains = [0.25 if t % 3 != 1 else 0.0 for t in range(T)]
    return make_schedule(m, B, H, W, T, fire_rate=frs, fire_u=fu, message_gains=gains, **kw)


def _loss(xT):
    return (xT[:, :4] ** 2).mean() + 0.1 * xT[:, 4:].mean()


def _run(m, x0, sched, impl):
    for p in m.parameters():
        p.grad = None
    x = x0.clone().requires_grad_(True)
    xT, hist = rollout(m, x, sched, return_history=True, impl=impl)
    _loss(xT).backward()
    grads = {n: (p.grad.clone() if p.grad is not None else None) for n, p in m.named_parameters()}
    return xT.detach(), hist.detach(), x.grad.clone(), grads


def _compare(m, x0, sched, tol_state=2e-6, tol_grad=2e-5):
    a = _run(m, x0, sched, "streaming")
    b = _run(m, x0, sched, "resident")
    assert rel_err(b[0].cpu(), a[0].cpu()) < tol_state
    assert rel_err(b[1].cpu(), a[1].cpu()) < tol_state
    assert torch.equal(GF.alive_mask(b[0], float(m.alpha_thr)), GF.alive_mask(a[0], float(m.alpha_thr)))
    assert rel_err(b[2].cpu(), a[2].cpu()) < tol_grad, rel_err(b[2].cpu(), a[2].cpu())
    for n, ga in a[3].items():
        gb = b[3][n]
        if ga is None or float(ga.abs().max()) == 0.0:
            assert gb is None or float(gb.abs().max()) <= 1e-8, n
            continue
        if any(s in n for s in ("query_proj", "key_proj", "scaling")):
            assert float(gb.abs().max()) <= 1e-8, n
            continue
        assert rel_err(gb.cpu(), ga.cpu()) < tol_grad, (n, rel_err(gb.cpu(), ga.cpu()))


@pytest.fixture(autouse=True)
def _clean_env():
    yield
    os.environ.pop("GNCA_RESIDENT_NC", None)
    os.environ.pop("GNCA_REP_SYNC", None)


@pytest.mark.parametrize("nc", ["8", "4", "2", "1"])
def test_every_cluster_size_fwd_bwd(nc):
    """NC = 2 and 1 exercise the global overflow of the in-smem u buffer; all sizes the balanced split."""
    os.environ["GNCA_RESIDENT_NC"] = nc
    m = graph_model(True)
    x0 = _grown_state(m, 3, 40, 40, steps=30)
    _compare(m, x0, _sched(m, 3, 40, 40, 6, seed=5))


def test_barrier_sync_mode_matches(monkeypatch):
    monkeypatch.setenv("GNCA_REP_SYNC", "barrier")       # this test only: every other test runs the default st.async exchange
    m = graph_model(True)
    x0 = _grown_state(m, 2, 40, 40, steps=20)
    _compare(m, x0, _sched(m, 2, 40, 40, 5, seed=6))


def test_damage_and_ragged_steps():
    m = graph_model(True)
    B, T = 4, 9
    x0 = _grown_state(m, B, 40, 40, steps=28)
    D = O.damage_mask("circle", B, 16, 40, 40, size=10, pos=[(20, 20), (18, 22), (25, 15), (12, 30)]).to(DEV)
    _compare(m, x0, _sched(m, B, 40, 40, T, seed=7, damage=D, damage_step=4, steps=[9, 3, 6, 0]))
    _compare(m, x0, _sched(m, B, 40, 40, T, seed=8, damage=D, damage_step=0, steps=[2, 9, 9, 5]))


def test_classic_model():
    m = classic_model()
    x0 = _grown_state(m, 5, 40, 40, steps=32)
    _compare(m, x0, _sched(m, 5, 40, 40, 7, seed=9))


@pytest.mark.parametrize("H,W", [(32, 32), (24, 40), (40, 24), (16, 64)])
def test_other_grids(H, W):
    m = graph_model(True)
    x0 = _grown_state(m, 2, H, W, steps=16)
    _compare(m, x0, _sched(m, 2, H, W, 5, seed=10))


def test_different_sender_threshold_general_mask_path():
    """model.alpha_thr != graph.alpha_thr: pre_alive(t+1) is no longer post_alive(t) of the same bitmap"""
    m = graph_model(True)
    m.graph.alpha_thr = 0.3
    x0 = _grown_state(m, 2, 40, 40, steps=26)
    _compare(m, x0, _sched(m, 2, 40, 40, 6, seed=11))


def test_large_batch_waves():
    """more samples than co-resident clusters of the preferred size"""
    m = graph_model(True)
    x0 = _grown_state(m, 40, 40, 40, steps=12)
    s = _sched(m, 40, 40, 40, 4, seed=12)
    with torch.no_grad():
        a = rollout(m, x0, s, impl="streaming")
        b = rollout(m, x0, s, impl="resident")
    assert rel_err(b.cpu(), a.cpu()) < 2e-6


def test_inference_does_not_allocate_history():
    m = graph_model(True)
    x0 = _grown_state(m, 2, 40, 40, steps=10)
    s = _sched(m, 2, 40, 40, 4, seed=13)
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    with torch.no_grad():
        rollout(m, x0, s, impl="resident")
    torch.cuda.synchronize()
    assert torch.cuda.max_memory_allocated() - base < 64 << 20


def test_resident_fwd_bwd_is_bitwise_deterministic():
    """No atomics, fixed reduction orders, fixed slot order: repeated runs must agree bit for bit (a race between the
    CTAs of a cluster or between warps would show up here as run-to-run noise)."""
    m = graph_model(True)
    B, T = 8, 14
    x0 = _grown_state(m, B, 40, 40, steps=30, seed=3)
    s = _sched(m, B, 40, 40, T, seed=21)
    ref = None
    for rep in range(4):
        out = _run(m, x0, s, "resident")
        flat = [out[0], out[1], out[2]] + [g for g in out[3].values() if g is not None]
        if ref is None:
            ref = flat
        else:
            for a, b in zip(ref, flat):
                assert torch.equal(a, b), rep


def test_batch_larger_than_any_cluster_split():
    """B = 80 > 74 co-resident 2-CTA clusters: one CTA per sample (NC = 1) in the forward and the backward"""
    m = graph_model(True)
    x0 = _grown_state(m, 80, 40, 40, steps=14, seed=4)
    _compare(m, x0, _sched(m, 80, 40, 40, 3, seed=22))


def test_banded_kernel_32_channels_vs_oracle():
    """40x40x32: the replicated-state kernel takes C = 16 only (a 40x40x32 replica is 205 KB), so the
    `resident` implementation runs the banded cluster kernel (gnca_resident.cu, k_resident_fwd<32>).  10 steps from
    ragged blobs against the fp64 oracle fed the same uniforms and offsets, and against the streaming kernels."""
    torch.manual_seed(11); random.seed(11)
    C, H, W, B, T = 32, 40, 40, 3, 10
    m = G.NeuralCAGraph(C, update_hidden=128, img_size=H, update_gain=0.1, alpha_thr=0.1, message_gain=0.3,
                        hidden_only=True, graph_zero_padded_shift=False)
    with torch.no_grad():
        m.update_net[2].weight.normal_(0, 0.05)
        m.norm.weight.uniform_(0.5, 1.5); m.norm.bias.normal_(0, 0.1)
    p = {k: v.detach().double() for k, v in m.state_dict().items()}
    m = m.to(DEV)
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    disk = (((yy - 19) ** 2 + (xx - 21) ** 2) < 11 ** 2).float()
    x0 = torch.rand(B, C, H, W) * disk
    x0[1, 3] *= (torch.rand(H, W) > 0.4).float()
    fu = torch.rand(T, B, H, W)
    chosen = [random.sample(m.graph.offsets, 8) for _ in range(T)]
    frs = [0.5 + 0.04 * t for t in range(T)]
    gains = [0.3 if t % 3 != 1 else 0.0 for t in range(T)]
    sched = make_schedule(m, B, H, W, T, fire_rate=frs, fire_u=fu.to(DEV), offsets=chosen, message_gains=gains)
    with torch.no_grad():
        os.environ["GNCA_DEBUG"] = "1"
        try:
            res = rollout(m, x0.to(DEV), sched, impl="resident")
        finally:
            os.environ.pop("GNCA_DEBUG", None)
        stream = rollout(m, x0.to(DEV), sched, impl="streaming")
    ref = x0.double()
    for t in range(T):
        cfg = O.StepConfig(update_gain=0.1, alpha_thr=0.1, graph=True, message_gain=gains[t], hidden_only=True,
                           zero_padded_shift=False)
        ref = O.nca_step(ref, p, cfg, frs[t], fu[t].unsqueeze(1).double(), chosen[t])
    assert rel_err(res.cpu(), ref.float()) < 1e-5, rel_err(res.cpu(), ref.float())
    assert rel_err(stream.cpu(), ref.float()) < 1e-5
    assert torch.equal(GF.alive_mask(res, 0.1).cpu(), O.alive_mask(ref.float(), 0.1))
    assert rel_err(res.cpu(), stream.cpu()) < 2e-6
