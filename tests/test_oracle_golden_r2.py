"""Pins oracle/nca_oracle.py against the round-2 fixtures recorded from the REFERENCE modules
(tests/golden/make_golden_r2.py): the constructor-flag branches (hidden_only=False, graph_alive_to_alive=False,
use_groupnorm=False, C=4, zero-pad with both), the 64-step gradient case and BASELINE configs[1] at the bench shape.
Fire uniforms are regenerated from the Philox replica (seed stored in the fixture); initial states from round-1 fixtures.
CPU only; runs in the `-m "not gpu"` suite."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_golden, load_params, rel_err
from oracle import nca_oracle as O
from philox_replica import fire_uniforms

T32 = lambda a: torch.from_numpy(np.asarray(a)).float()
tup = lambda ch: [tuple(int(v) for v in o) for o in ch]

# name -> (oracle StepConfig kwargs, weights, C)
FLAG_CASES = {
    "hidden_all": (dict(hidden_only=False), "graph", 16),
    "no_a2a": (dict(alive_to_alive=False), "graph", 16),
    "no_gn": (dict(use_groupnorm=False), "graph", 16),
    "zeropad_hidden_all_no_a2a": (dict(hidden_only=False, alive_to_alive=False, zero_padded_shift=True), "graph", 16),
    "c4_hidden_all": (dict(hidden_only=False), "own", 4),
    "classic_no_gn": (dict(use_groupnorm=False, graph=False, update_gain=0.1, alpha_thr=0.1), "classic", 16),
}


def flag_case_inputs(name):
    """(fixture, oracle config, parameter dict, x0) of a flags_<name>.npz case."""
    kw, wsrc, C = FLAG_CASES[name]
    g = load_golden(f"flags_{name}.npz")
    base = dict(update_gain=0.05, alpha_thr=0.12, graph=True, message_gain=0.25, hidden_only=True, zero_padded_shift=False)
    base.update(kw)
    cfg = O.StepConfig(**base)
    if wsrc == "own":
        p = {k[2:]: T32(v) for k, v in g.items() if k.startswith("w:")}
    else:
        p = load_params("weights_graph_ep960.npz" if wsrc == "graph" else "weights_classic_ep990.npz")
    x0 = T32(load_golden("graph_torus_rollout.npz")["x_48"])[:, :C].contiguous()
    return g, cfg, p, x0


def oracle_grad_case(g, cfg, p, x0, target):
    """Rollout + premultiplied loss + backward through the oracle, draws from the fixture / the Philox replica."""
    B, C, H, W = x0.shape
    T = len(g["fire_rates"])
    u = torch.from_numpy(fire_uniforms(int(g["philox_seed"]), 0, T, B, H, W))
    steps = torch.from_numpy(g["steps"]).long() if "steps" in g else None
    po = {k: v.clone().requires_grad_(v.is_floating_point() and "perception" not in k) for k, v in p.items()}
    xo = x0.clone().requires_grad_(True)
    fus = [u[t].unsqueeze(1) if steps is None else u[t][steps > t].unsqueeze(1) for t in range(T)]
    chosens = [tup(c) for c in g["chosen"]] if "chosen" in g else None
    if steps is not None and chosens is not None:      # the reference draws offsets only on steps with an active sample
        it = iter(chosens)
        chosens = [next(it) if bool((steps > t).any()) else () for t in range(T)]
    xT = O.rollout(xo, po, cfg, g["fire_rates"].tolist(), fus, chosens,
                   g["gains"].tolist() if cfg.graph else None, steps)
    per = O.loss_premult_rgba(xT[:, :4], target[:min(4, C)].unsqueeze(0).expand(B, -1, -1, -1))
    per.mean().backward()
    return xT.detach(), per.detach(), xo.grad, po


def check_param_grads(g, get_grad, tol=1e-4, skip_tiny=("query_proj", "key_proj", "scaling"), torus=True):
    for k, v in g.items():
        if not k.startswith("grad:") or v.size == 0:
            continue
        name = k[5:]
        ours = get_grad(name)
        if any(s in name for s in skip_tiny):
            if torus:
                assert ours is None or float(ours.abs().max()) <= 1e-8, name      # true gradient is 0
            else:
                assert rel_err(ours, v) < 1e-2, (name, rel_err(ours, v))
            continue
        assert ours is not None, name
        assert rel_err(ours, v) < tol, (name, rel_err(ours, v))


@pytest.mark.parametrize("name", sorted(FLAG_CASES))
def test_oracle_flag_branches(name):
    g, cfg, p, x0 = flag_case_inputs(name)
    target = T32(np.load(os.path.join(GOLDEN, "target_gecko_surrogate.npy")))
    B, C, H, W = x0.shape
    u = torch.from_numpy(fire_uniforms(int(g["philox_seed"]), 0, 1, B, H, W))
    c1 = O.StepConfig(**{**cfg.__dict__, "message_gain": float(g["gains"][0])})
    with torch.no_grad():
        x1 = O.nca_step(x0, p, c1, float(g["fire_rates"][0]), u[0].unsqueeze(1), tup(g["chosen"][0]) if cfg.graph else ())
    assert rel_err(x1, g["x_1"]) < 1e-6, rel_err(x1, g["x_1"])
    xT, per, gx, po = oracle_grad_case(g, cfg, p, x0, target)
    assert rel_err(xT, g["x_T"]) < 1e-6
    assert rel_err(per, g["per_sample"]) < 1e-5
    assert rel_err(gx, g["grad_x0"]) < 1e-4
    check_param_grads(g, lambda n: po[n].grad if n in po else None, torus=not cfg.zero_padded_shift)


def test_oracle_64_step_gradients():
    g = load_golden("grads64_b8.npz")
    x48 = T32(load_golden("graph_torus_rollout.npz")["x_48"])
    x0 = torch.cat([O.make_seed(16, 40, 4), x48, x48.flip(0)], 0)
    cfg = O.StepConfig(update_gain=0.05, alpha_thr=0.12, graph=True, message_gain=0.25, hidden_only=True,
                       zero_padded_shift=False)
    target = T32(np.load(os.path.join(GOLDEN, "target_gecko_surrogate.npy")))
    xT, per, gx, po = oracle_grad_case(g, cfg, load_params("weights_graph_ep960.npz"), x0, target)
    assert rel_err(xT, g["x_T"]) < 1e-5
    assert rel_err(per, g["per_sample"]) < 1e-5
    assert rel_err(gx, g["grad_x0"]) < 1e-4
    check_param_grads(g, lambda n: po[n].grad if n in po else None)


def test_oracle_bench_shape_forward():
    g = load_golden("c2_bench_shape.npz")
    B, T = 8, 96
    u = torch.from_numpy(fire_uniforms(int(g["philox_seed"]), 0, T, B, 40, 40))
    cfg = O.StepConfig(update_gain=0.05, alpha_thr=0.12, graph=True, message_gain=0.25, hidden_only=True,
                       zero_padded_shift=False)
    with torch.no_grad():
        x = O.rollout(O.make_seed(16, 40, B), load_params("weights_graph_ep960.npz"), cfg, [0.5] * T,
                      [u[t].unsqueeze(1) for t in range(T)], [tup(c) for c in g["chosen"]])
    assert rel_err(x, g["x_96"]) < 1e-5
    assert torch.equal(O.alive_mask(x, 0.12), O.alive_mask(T32(g["x_96"]), 0.12))
