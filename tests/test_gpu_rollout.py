"""GPU parity of the rollout entry points (gnca_rollout_fwd/_bwd through rollout.py) against golden fixtures."""
import os
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden, load_params, max_rel, rel_err
from oracle import nca_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import graph_neural_cellular_automata_b200 as G
    from graph_neural_cellular_automata_b200 import functional as GF
    from graph_neural_cellular_automata_b200.rollout import make_schedule, rollout
    from test_gpu_step import graph_model, classic_model, ocfg, T32, tup, DEV

IMPLS = ["streaming", "resident", "banded"]


def _supported(impl, fn):
    from graph_neural_cellular_automata_b200._lib import GncaError
    try:
        return fn()
    except GncaError as e:
        if impl in ("resident", "banded") and "unsupported" in str(e):
            pytest.skip("resident kernel not available for this configuration")
        raise


@pytest.mark.parametrize("impl", IMPLS)
def test_rollout_matches_golden_torus(impl):
    g = load_golden("graph_torus_rollout.npz")
    m = graph_model(True)
    x0 = T32(g["x_0"]).to(DEV)
    T = 48
    sched = make_schedule(m, 2, 40, 40, T, fire_rate=float(g["fire_rate"]), offsets=[tup(c) for c in g["chosen"]],
                          fire_u=T32(g["fire_u"]).to(DEV))
    with torch.no_grad():
        xT, hist = _supported(impl, lambda: rollout(m, x0, sched, return_history=True, impl=impl))
    assert torch.equal(hist[T], xT)
    for t in (1, 8, 16, 32, 48):
        ref = T32(g[f"x_{t}"])
        assert rel_err(hist[t].cpu(), ref) < 1e-5, (t, rel_err(hist[t].cpu(), ref))
        assert torch.equal(GF.alive_mask(hist[t], 0.12).cpu(), O.alive_mask(ref, 0.12)), t


@pytest.mark.parametrize("impl", IMPLS)
def test_rollout_equals_sequential_steps(impl):
    """The contract of the extension: bit-identical to T module.step() calls with the same draws."""
    m = graph_model(True)
    g = load_golden("graph_torus_step.npz")
    x0 = T32(g["x_in"]).to(DEV)
    T = 7
    random.seed(3)
    offs = [m.graph.draw_offsets() for _ in range(T)]
    fu = torch.rand(T, 2, 40, 40, device=DEV)
    gains = [0.25 if t % 3 == 0 else 0.0 for t in range(T)]
    frs = [0.5 + 0.05 * t for t in range(T)]
    sched = make_schedule(m, 2, 40, 40, T, fire_rate=frs, offsets=offs, fire_u=fu, message_gains=gains)
    with torch.no_grad():
        xT = _supported(impl, lambda: rollout(m, x0, sched, impl=impl))
        x = x0
        for t in range(T):
            x = m.step(x, frs[t], fire_u=fu[t].unsqueeze(1), chosen=offs[t], message_gain=gains[t])
    if impl == "streaming":
        assert torch.equal(xT, x)
    else:
        assert rel_err(xT.cpu(), x.cpu()) < 1e-6


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("fname,kind", [("graph_torus_grads.npz", "graph"), ("graph_torus_grads_ragged.npz", "graph"),
                                        ("classic_grads.npz", "classic"), ("graph_zeropad_grads.npz", "zeropad")])
def test_rollout_grads_golden(impl, fname, kind):
    g = load_golden(fname)
    m = graph_model(True) if kind == "graph" else (graph_model(False) if kind == "zeropad" else classic_model())
    T = len(g["gains"])
    B = g["x0"].shape[0]
    x0 = T32(g["x0"]).to(DEV).requires_grad_(True)
    steps = g["steps"].tolist() if "steps" in g else None
    sched = make_schedule(m, B, 40, 40, T, fire_rate=g["fire_rates"].tolist(),
                          offsets=[tup(c) for c in g["chosen"]] if kind != "classic" else None,
                          fire_u=T32(g["fire_u"]).to(DEV), message_gains=g["gains"].tolist(), steps=steps)
    target = T32(np.load(os.path.join(GOLDEN, "target_gecko_surrogate.npy"))).to(DEV)
    xT = _supported(impl, lambda: rollout(m, x0, sched, impl=impl))
    pred = xT[:, :4]
    rgba = torch.cat([pred[:, :3] * pred[:, 3:4], pred[:, 3:4]], 1)
    per = ((rgba - target.unsqueeze(0)) ** 2).mean(dim=(1, 2, 3))
    per.mean().backward()
    assert rel_err(xT.detach().cpu(), g["x_T"]) < 1e-5
    assert rel_err(per.detach().cpu(), g["per_sample"]) < 1e-5       # 1e-5 rel loss
    assert rel_err(x0.grad.cpu(), g["grad_x0"]) < 1e-4
    named = dict(m.named_parameters())
    for k, v in g.items():
        if not k.startswith("grad:") or v.size == 0:
            continue
        name = k[5:]
        ours = named[name].grad if named[name].grad is not None else torch.zeros_like(named[name])
        if kind == "graph" and any(s in name for s in ("query_proj", "key_proj", "scaling")):
            assert float(ours.abs().max()) <= 1e-8, name
        elif kind == "zeropad" and any(s in name for s in ("query_proj", "key_proj", "scaling")):
            assert rel_err(ours.cpu(), v) < 1e-3, (name, rel_err(ours.cpu(), v))   # tiny (~1e-3) second-order path
        else:
            assert rel_err(ours.cpu(), v) < 1e-4, (name, rel_err(ours.cpu(), v))


@pytest.mark.parametrize("impl", IMPLS)
def test_philox_fire_stream(impl):
    """In-kernel Philox fire masks are THE stream of tests/philox_replica.py, bit for bit: a rollout that draws its
    fire uniforms in the kernel (seed, offset) equals the same rollout fed the replica's uniforms as recorded `fire_u`
    (same kernels, only the source of the uniforms differs -> torch.equal), for ragged step counts, per-step fire rates
    and a non-zero stream offset.  This is the path bench.py times (fire="philox")."""
    from philox_replica import fire_uniforms
    m = graph_model(True)
    x0 = T32(load_golden("graph_torus_step.npz")["x_in"]).to(DEV)
    x0 = torch.cat([x0, x0.flip(0), x0], 0).contiguous()                 # B = 6
    B, T = 6, 9
    frs = [0.5, 0.9, 0.3, 0.7, 0.55, 0.5, 0.99, 0.1, 0.5]
    steps = [9, 4, 0, 9, 7, 1]
    for seed, offset in ((11, 0), (2 ** 40 + 7, 12345)):
        random.seed(1)
        s1 = make_schedule(m, B, 40, 40, T, fire_rate=frs, seed=seed, steps=steps)
        s1.philox_offset = offset
        offs = s1.offsets.cpu().numpy()
        u = torch.from_numpy(fire_uniforms(seed, offset, T, B, 40, 40)).to(DEV)
        s2 = make_schedule(m, B, 40, 40, T, fire_rate=frs, offsets=offs, fire_u=u, steps=steps)
        with torch.no_grad():
            a = _supported(impl, lambda: rollout(m, x0, s1, impl=impl))
            b = rollout(m, x0, s2, impl=impl)
        assert torch.equal(a, b), (impl, seed)
    random.seed(1)
    s3 = make_schedule(m, B, 40, 40, T, fire_rate=frs, seed=12, steps=steps)
    with torch.no_grad():
        c = rollout(m, x0, s3, impl=impl)
    assert not torch.equal(a, c) and torch.isfinite(c).all()


@pytest.mark.parametrize("impl", ["resident", "streaming"])
def test_philox_rollout_bench_shape_vs_oracle(impl):
    """BASELINE configs[1] exactly as bench.py runs it (B=8, T=96, 40x40x16, fire 0.5 from the in-kernel Philox stream,
    seed growth, trained weights, torus, 8 of 72 offsets per step) against the ORACLE fed the replica's uniforms:
    final state <= 1e-5 rel-Frobenius, alive mask bit-exact."""
    from philox_replica import fire_uniforms
    from oracle.nca_oracle import make_seed
    m = graph_model(True)
    B, T, seed = 8, 96, 4242
    x0 = make_seed(16, 40, B)
    random.seed(42)
    sched = make_schedule(m, B, 40, 40, T, fire_rate=0.5, seed=seed)
    offs = sched.offsets.cpu().numpy()
    with torch.no_grad():
        xT = _supported(impl, lambda: rollout(m, x0.to(DEV), sched, impl=impl))
    u = torch.from_numpy(fire_uniforms(seed, 0, T, B, 40, 40))
    p = load_params("weights_graph_ep960.npz")
    with torch.no_grad():
        ref = O.rollout(x0, p, ocfg(True), [0.5] * T, [u[t].unsqueeze(1) for t in range(T)],
                        [tup(offs[t]) for t in range(T)])
    assert float((ref[:, 3] > 0.12).float().mean()) > 0.05            # the pattern grew: a live test
    assert rel_err(xT.cpu(), ref) < 1e-5, rel_err(xT.cpu(), ref)
    assert torch.equal(GF.alive_mask(xT, 0.12).cpu(), O.alive_mask(ref, 0.12))


def test_damage_in_rollout():
    """A multiplicative damage mask at step t (regeneration test, test_graph_augmented_regeneration.py:185-189)."""
    m = graph_model(True)
    g = load_golden("graph_torus_step.npz")
    x0 = T32(g["x_in"]).to(DEV)
    T, td = 6, 3
    D = O.damage_mask("circle", 2, 16, 40, 40, size=6, pos=[(20, 20), (18, 22)]).to(DEV)
    random.seed(9)
    offs = [m.graph.draw_offsets() for _ in range(T)]
    fu = torch.rand(T, 2, 40, 40, device=DEV)
    sched = make_schedule(m, 2, 40, 40, T, fire_rate=0.5, offsets=offs, fire_u=fu, damage=D, damage_step=td)
    with torch.no_grad():
        xT = rollout(m, x0, sched, impl="streaming")
        x = x0
        for t in range(T):
            if t == td:
                x = x * D
            x = m.step(x, 0.5, fire_u=fu[t].unsqueeze(1), chosen=offs[t])
    assert torch.equal(xT, x)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("kind", ["square", "circle", "stripes", "gaussian", "alpha_drop", "saltpepper"])
def test_damage_descriptor_in_rollout(impl, kind):
    """A damage DESCRIPTOR (utils/damage.py `Damage`: per-cell plane, all channels or alpha only) applied in-kernel at
    step td == rollout to td, multiply by the dense mask, continue; and forward + all gradients are bit-identical to the
    same rollout fed the equivalent dense [B,C,H,W] mask (checks the plane indexing of every kernel, backward included)."""
    from graph_neural_cellular_automata_b200.utils import damage as DM
    m = graph_model(True)
    g = load_golden("graph_torus_step.npz")
    x0 = T32(g["x_in"]).to(DEV)
    T, td = 6, 3
    torch.manual_seed(4); random.seed(4)
    Dm = {"square": lambda: DM.square_mask(x0, 9), "circle": lambda: DM.circle_mask(x0, 6),
          "stripes": lambda: DM.stripe_mask(x0, 5, "auto"), "gaussian": lambda: DM.gaussian_mask(x0, 6, 0.35),
          "alpha_drop": lambda: DM.alpha_dropout_mask(x0, 0.3, 0.12, True), "saltpepper": lambda: DM.salt_pepper_mask(x0, 0.2)}[kind]()
    dense = Dm.dense(x0)
    assert Dm.layout == (2 if kind == "saltpepper" else 1) and float((dense != 1).sum()) > 0
    if kind == "saltpepper":
        assert bool((dense[:, :3] == 1).all()) and bool((dense[:, 4:] == 1).all())
    offs = [m.graph.draw_offsets() for _ in range(T)]
    fu = torch.rand(T, 2, 40, 40, device=DEV)

    def run(dmg):
        for p_ in m.parameters():
            p_.grad = None
        sched = make_schedule(m, 2, 40, 40, T, fire_rate=0.5, offsets=offs, fire_u=fu, damage=dmg, damage_step=td)
        xg = x0.clone().requires_grad_(True)
        xT = _supported(impl, lambda: rollout(m, xg, sched, impl=impl))
        xT[:, :4].square().mean().backward()
        return xT.detach(), xg.grad.clone(), {n: p_.grad.clone() for n, p_ in m.named_parameters() if p_.grad is not None}

    a, ga, pa = run(Dm)
    b, gb, pb = run(dense)
    assert torch.equal(a, b) and torch.equal(ga, gb)
    assert all(torch.equal(pa[n], pb[n]) for n in pa)
    with torch.no_grad():
        s1 = make_schedule(m, 2, 40, 40, td, fire_rate=0.5, offsets=offs[:td], fire_u=fu[:td].contiguous())
        s2 = make_schedule(m, 2, 40, 40, T - td, fire_rate=0.5, offsets=offs[td:], fire_u=fu[td:].contiguous())
        mid = rollout(m, x0, s1, impl="streaming")
        two = rollout(m, (mid * dense).contiguous(), s2, impl="streaming")
    assert rel_err(a.cpu(), two.cpu()) < 1e-6


@pytest.mark.parametrize("impl", ["auto", "resident", "streaming"])
def test_zeropad_rollout_matches_reference(impl):
    """The MODULE DEFAULT graph shift (zero-padded: dy-only senders + per-sample softmax weights over the offsets,
    graph_augmentation.py:85-92,136-154) through the rollout entry point: 12 reference steps from an aged state
    (`graph_zeropad_rollout.npz`).  `auto` / `resident` run the ZP instantiation of the replicated cluster kernel, which
    recomputes the attention weights from the resident state every step."""
    g = load_golden("graph_zeropad_rollout.npz")
    m = graph_model(False)
    x0 = T32(g["x_0"]).to(DEV)
    T = 12
    sched = make_schedule(m, 2, 40, 40, T, fire_rate=float(g["fire_rate"]), offsets=[tup(c) for c in g["chosen"]],
                          fire_u=T32(g["fire_u"]).to(DEV))
    with torch.no_grad():
        xT = _supported(impl, lambda: rollout(m, x0, sched, impl=impl))
    ref = T32(g["x_12"])
    assert rel_err(xT.cpu(), ref) < 1e-5, rel_err(xT.cpu(), ref)
    assert torch.equal(GF.alive_mask(xT, 0.12).cpu(), O.alive_mask(ref, 0.12))
