"""Single-step backward (streaming kernels via module.step autograd) vs oracle autograd on the states of the 64-step case."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import load_golden, load_params, rel_err
from oracle import nca_oracle as O
import test_gpu_r2 as T
from philox_replica import fire_uniforms
from graph_neural_cellular_automata_b200.rollout import make_schedule, rollout
g = load_golden("grads64_b8.npz")
x48 = T.T32(load_golden("graph_torus_rollout.npz")["x_48"])
x0 = torch.cat([O.make_seed(16, 40, 4), x48, x48.flip(0)], 0)
chosen = [T.tup(c) for c in g["chosen"]]
u = torch.from_numpy(fire_uniforms(int(g["philox_seed"]), 0, 64, 8, 40, 40))
m = T._graph_model()
sched = make_schedule(m, 8, 40, 40, 64, fire_rate=g["fire_rates"].tolist(), offsets=chosen, message_gains=g["gains"].tolist(), fire_u=u.cuda())
with torch.no_grad():
    xT, hist = rollout(m, x0.cuda(), sched, return_history=True, impl="streaming")
p = load_params("weights_graph_ep960.npz")
torch.manual_seed(0)
gup = torch.randn(8, 16, 40, 40)
for t in range(12, 30):
    xt = hist[t].detach().clone()
    fr, gain = float(g["fire_rates"][t]), float(g["gains"][t])
    xg = xt.clone().requires_grad_(True)
    out = m.step(xg, fr, fire_u=u[t].unsqueeze(1).cuda(), chosen=chosen[t], message_gain=gain)
    (out * gup.cuda()).sum().backward()
    xo = xt.cpu().double().requires_grad_(True)
    cfg = O.StepConfig(update_gain=0.05, alpha_thr=0.12, graph=True, message_gain=gain, hidden_only=True, zero_padded_shift=False)
    ref, aux = O.nca_step(xo, {k: v.double() for k, v in p.items()}, cfg, fr, u[t].unsqueeze(1).double(), chosen[t], return_aux=True)
    (ref * gup.double()).sum().backward()
    per = [rel_err(xg.grad[i].cpu(), xo.grad[i]) for i in range(8)]
    nact = (aux["pre"] * aux["fire"]).sum(dim=(1, 2, 3)).int().tolist()
    # active cells per 256-cell chunk of sample 7
    act7 = (aux["pre"][7, 0] * aux["fire"][7, 0]).flatten()
    per_chunk = [int(act7[i:i + 256].sum()) for i in range(0, 1600, 256)]
    flag = "  <<<" if max(per) > 1e-4 else ""
    print(f"t={t:2d} gain={gain:.2f} fwd {rel_err(out.detach().cpu(), ref.detach()):.1e} dL/dx per sample:", " ".join(f"{v:.0e}" for v in per), "| active/chunk s7:", per_chunk, flag)
    for pp in m.parameters(): pp.grad = None
print("---- t = 18, sample 7 ----")
t = 18
xt = hist[t].detach().clone()
fr, gain = float(g["fire_rates"][t]), float(g["gains"][t])
xg = xt.clone().requires_grad_(True)
out = m.step(xg, fr, fire_u=u[t].unsqueeze(1).cuda(), chosen=chosen[t], message_gain=gain)
(out * gup.cuda()).sum().backward()
xo = xt.cpu().double().requires_grad_(True)
cfg = O.StepConfig(update_gain=0.05, alpha_thr=0.12, graph=True, message_gain=gain, hidden_only=True, zero_padded_shift=False)
ref, aux = O.nca_step(xo, {k: v.double() for k, v in p.items()}, cfg, fr, u[t].unsqueeze(1).double(), chosen[t], return_aux=True)
(ref * gup.double()).sum().backward()
d = (xg.grad[7].cpu().double() - xo.grad[7]).abs()
print("grad norm", float(xo.grad[7].norm()), "err norm", float((xg.grad[7].cpu().double() - xo.grad[7]).norm()))
bad = (d.amax(0) > 1e-4 * float(xo.grad[7].abs().max())).nonzero()
print("cells with error:", bad.tolist()[:30], "n", len(bad))
print("per-channel max err:", [f"{float(v):.1e}" for v in d.amax(dim=(1, 2))])
print("chosen offsets:", chosen[t])
alpha = xt[7, 3].cpu()
for (yy, xx) in bad.tolist()[:6]:
    print(f"  cell ({yy},{xx}) alpha {float(alpha[yy, xx]):.6f} pre {float(aux['pre'][7,0,yy,xx])} fire {float(aux['fire'][7,0,yy,xx])} post {float(aux['post'][7,0,yy,xx])} "
          f"maxpool_alpha {float(torch.nn.functional.max_pool2d(alpha[None,None], 3, 1, 1)[0,0,yy,xx]):.6f}")
# sender alive values near threshold anywhere?
mp = torch.nn.functional.max_pool2d(alpha[None, None], 3, 1, 1)[0, 0]
near = ((mp - 0.12).abs() < 1e-5).nonzero()
print("cells with maxpool(alpha) within 1e-5 of the threshold:", near.tolist(), [float(mp[a, b]) for a, b in near.tolist()])
yv = O.perception(xt.cpu().double())[7, :, 35, 34]
W1 = p["update_net.0.weight"].double().view(128, 48); b1 = p["update_net.0.bias"].double()
pre = W1 @ yv + b1
j = int(pre.abs().argmin())
pre32 = (p["update_net.0.weight"].view(128, 48) @ O.perception(xt.cpu())[7, :, 35, 34] + p["update_net.0.bias"])
print(f"cell (35,34): hidden unit {j} pre-activation fp64 {float(pre[j]):.3e} (fp32 matmul: {float(pre32[j]):.3e}); next smallest |pre| {float(pre.abs().sort().values[1]):.3e}")
