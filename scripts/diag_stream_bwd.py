"""Where does the streaming BPTT deviate from the resident one on the 64-step case?  (sample 7 of grads64_b8)"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import load_golden, rel_err, GOLDEN
from oracle import nca_oracle as O
import test_gpu_r2 as T
from philox_replica import fire_uniforms
from graph_neural_cellular_automata_b200.rollout import make_schedule, rollout
g = load_golden("grads64_b8.npz")
x48 = T.T32(load_golden("graph_torus_rollout.npz")["x_48"])
x0 = torch.cat([O.make_seed(16, 40, 4), x48, x48.flip(0)], 0)
target = T.T32(np.load(os.path.join(GOLDEN, "target_gecko_surrogate.npy"))).cuda()
chosen = [T.tup(c) for c in g["chosen"]]
u = torch.from_numpy(fire_uniforms(int(g["philox_seed"]), 0, 64, 8, 40, 40)).cuda()
for Tq in (4, 8, 16, 24, 32, 48, 64):
    res = {}
    for impl in ("resident", "streaming"):
        m = T._graph_model()
        sched = make_schedule(m, 8, 40, 40, Tq, fire_rate=g["fire_rates"][:Tq].tolist(), offsets=chosen[:Tq],
                              message_gains=g["gains"][:Tq].tolist(), fire_u=u[:Tq].contiguous())
        xg = x0.cuda().requires_grad_(True)
        xT = rollout(m, xg, sched, impl=impl)
        T._loss(xT, target).mean().backward()
        res[impl] = (xg.grad.clone(), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None})
    a, b = res["resident"][0], res["streaming"][0]
    per = [rel_err(b[i].cpu(), a[i].cpu()) for i in range(8)]
    dd = (b[7] - a[7]).abs()
    idx = int(dd.flatten().argmax()); c, r = divmod(idx, 1600); y, x = divmod(r, 40)
    print(f"T={Tq:2d} streaming vs resident dL/dx0 per sample:", " ".join(f"{v:.1e}" for v in per), f"| sample 7 max at ch {c} ({y},{x}) ncells>1e-3*max: {int((dd.amax(0) > 1e-3 * a[7].abs().max()).sum())}")
