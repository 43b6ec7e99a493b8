import sys, os, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, numpy as np
from conftest import load_golden
from test_gpu_step import graph_model, T32, DEV
from graph_neural_cellular_automata_b200.rollout import make_schedule, rollout
from graph_neural_cellular_automata_b200.utils import damage as DM
from graph_neural_cellular_automata_b200 import functional as GF
m = graph_model(True)
x0 = T32(load_golden("graph_torus_step.npz")["x_in"]).to(DEV)
T, td = 6, 3
random.seed(4); torch.manual_seed(4)
offs = [m.graph.draw_offsets() for _ in range(T)]
fu = torch.rand(T, 2, 40, 40, device=DEV)
for orient in ("h", "v"):
    for s0 in (3, 17, 30):
        pos = torch.tensor([[s0, 0]], dtype=torch.int64, device=DEV)
        Dm = DM._plane(x0, DM.DK_STRIPE_H if orient == "h" else DM.DK_STRIPE_V, 5, pos=pos)
        out = {}
        for impl in ("streaming", "resident"):
            sched = make_schedule(m, 2, 40, 40, T, fire_rate=0.5, offsets=offs, fire_u=fu, damage=Dm, damage_step=td)
            with torch.no_grad():
                xT, hist = rollout(m, x0, sched, return_history=True, impl=impl)
            out[impl] = hist
        d = [(float((out["resident"][t] - out["streaming"][t]).abs().max())) for t in range(T + 1)]
        am = [int((GF.alive_mask(out["resident"][t], 0.12) != GF.alive_mask(out["streaming"][t], 0.12)).sum()) for t in range(T + 1)]
        print(orient, s0, "max|d| per step:", " ".join(f"{v:.1e}" for v in d), " alive-mask mismatches:", am)
        t = td + 1
        dd = (out["resident"][t] - out["streaming"][t]).abs()
        if float(dd.max()) > 1e-5:
            bad = (dd.amax(1) > 1e-5).nonzero()
            print("    step", t, "cells differing > 1e-5:", len(bad), "first:", bad[:8].tolist(), " per-sample max:", dd.amax(dim=(1, 2, 3)).tolist())
print("---- gradient mode (records) ----")
torch.manual_seed(4); random.seed(4)
Dm = DM.stripe_mask(x0, 5, "auto")
dense = Dm.dense(x0)
rows = (dense[0, 0] == 0).all(1).nonzero().flatten().tolist(); cols = (dense[0, 0] == 0).all(0).nonzero().flatten().tolist()
print("stripe rows", rows, "cols", cols)
offs = [m.graph.draw_offsets() for _ in range(T)]
fu = torch.rand(T, 2, 40, 40, device=DEV)
res = {}
for impl in ("streaming", "resident"):
    for grad in (False, True):
        sched = make_schedule(m, 2, 40, 40, T, fire_rate=0.5, offsets=offs, fire_u=fu, damage=Dm, damage_step=td)
        xg = x0.clone().requires_grad_(grad)
        with torch.set_grad_enabled(grad):
            for p_ in m.parameters(): p_.requires_grad_(grad)
            res[(impl, grad)] = rollout(m, xg, sched, impl=impl).detach()
ref = res[("streaming", False)]
for k, v in res.items():
    d = (v - ref).abs()
    print(k, f"max|d| {float(d.max()):.2e}", "per-sample", d.amax(dim=(1, 2, 3)).tolist(), " cells>1e-5:", int((d.amax(1) > 1e-5).sum()))
print("---- sweep of stripe positions with the failing draws ----")
with torch.no_grad():
    for p_ in m.parameters(): p_.requires_grad_(False)
    for orient in ("h", "v"):
        line = []
        for s0 in range(0, 36):
            pos = torch.tensor([[s0, 0]], dtype=torch.int64, device=DEV)
            Dq = DM._plane(x0, DM.DK_STRIPE_H if orient == "h" else DM.DK_STRIPE_V, 5, pos=pos)
            r = {}
            for impl in ("streaming", "resident"):
                sched = make_schedule(m, 2, 40, 40, T, fire_rate=0.5, offsets=offs, fire_u=fu, damage=Dq, damage_step=td)
                r[impl] = rollout(m, x0, sched, impl=impl)
            d = (r["resident"] - r["streaming"]).abs().amax(dim=(1, 2, 3))
            line.append("X" if float(d.max()) > 1e-5 else ".")
        print(orient, "".join(line))
    # the failing one: which step first?
    pos = torch.tensor([[29, 0]], dtype=torch.int64, device=DEV)
    Dq = DM._plane(x0, DM.DK_STRIPE_H, 5, pos=pos)
    for tdd in range(0, T):
        r = {}
        for impl in ("streaming", "resident"):
            sched = make_schedule(m, 2, 40, 40, T, fire_rate=0.5, offsets=offs, fire_u=fu, damage=Dq, damage_step=tdd)
            r[impl] = rollout(m, x0, sched, impl=impl)
        print("damage_step", tdd, "max|d| per sample", (r["resident"] - r["streaming"]).abs().amax(dim=(1, 2, 3)).tolist())
    for Tq in range(td + 1, T + 1):
        r = {}
        for impl in ("streaming", "resident"):
            sched = make_schedule(m, 2, 40, 40, Tq, fire_rate=0.5, offsets=offs[:Tq], fire_u=fu[:Tq].contiguous(), damage=Dq, damage_step=td)
            r[impl] = rollout(m, x0, sched, impl=impl)
        dd = (r["resident"] - r["streaming"]).abs()
        print("T", Tq, "max|d| per sample", dd.amax(dim=(1, 2, 3)).tolist(), "cells>1e-5 in sample 0:", int((dd[0].amax(0) > 1e-5).sum()),
              "alive", int((r["streaming"][0, 3] > 0.12).sum()))
