"""Per-sample / per-tensor error of the 64-step gradient case for every rollout implementation (GPU diagnostic)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import load_golden, rel_err
from oracle import nca_oracle as O
import test_gpu_r2 as T

x48 = T.T32(load_golden("graph_torus_rollout.npz")["x_48"])
x0 = torch.cat([O.make_seed(16, 40, 4), x48, x48.flip(0)], 0)
for impl in ("streaming", "resident", "banded"):
    for fire in ("philox", "recorded"):
        try:
            g, m, xT, per, gx = T._grad_case("grads64_b8.npz", x0, impl, fire)
        except Exception as e:
            print(impl, fire, "ERR", str(e)[:80]); continue
        print(f"{impl:10s} {fire:9s} x_T {rel_err(xT, g['x_T']):.2e} per {rel_err(per, g['per_sample']):.2e} grad_x0 {rel_err(gx, g['grad_x0']):.2e}",
              " per-sample:", " ".join(f"{rel_err(gx[b], g['grad_x0'][b]):.1e}" for b in range(8)))
        named = dict(m.named_parameters())
        print("     ", " ".join(f"{k[5:].replace('update_net','un').replace('graph.','g.')}={rel_err(named[k[5:]].grad.cpu(), v):.1e}"
                                for k, v in g.items() if k.startswith('grad:') and v.size and named[k[5:]].grad is not None))
