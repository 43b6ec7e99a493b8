import sys, os, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import graph_neural_cellular_automata_b200 as G
from graph_neural_cellular_automata_b200 import functional as GF
from oracle import nca_oracle as O
from conftest import max_rel, rel_err
torch.manual_seed(1); random.seed(1)
C, Hh, Ww, B, hid = 16, 128, 128, 20, 128
m = G.NeuralCAGraph(C, update_hidden=hid, img_size=Hh, update_gain=0.1, alpha_thr=0.1, message_gain=0.3, hidden_only=True, graph_zero_padded_shift=False)
with torch.no_grad():
    m.update_net[2].weight.normal_(0, 0.05)
    m.norm.weight.uniform_(0.5, 1.5); m.norm.bias.normal_(0, 0.1)
p = {k: v.detach().clone() for k, v in m.state_dict().items()}
m = m.to("cuda")
yy, xx = torch.meshgrid(torch.arange(Hh), torch.arange(Ww), indexing="ij")
disk = (((yy - Hh / 2) ** 2 + (xx - Ww / 2) ** 2) < (0.3 * Hh) ** 2).float()
x = torch.rand(B, C, Hh, Ww) * disk
x[B // 2:, 3] *= (torch.rand(B - B // 2, Hh, Ww) > 0.5).float()
fus = [torch.rand(B, 1, Hh, Ww) for _ in range(2)]
chosen = [random.sample(m.graph.offsets, 8) for _ in range(2)]
cfg = O.StepConfig(update_gain=0.1, alpha_thr=0.1, graph=True, message_gain=0.3, hidden_only=True, zero_padded_shift=False)
ref = O.nca_step(x, p, cfg, 0.5, fus[0], chosen[0])
p64 = {k: v.double() for k, v in p.items()}
ref64, aux = O.nca_step(x.double(), p64, cfg, 0.5, fus[0].double(), chosen[0], return_aux=True)
with torch.no_grad():
    out = m.step(x.to("cuda"), 0.5, fire_u=fus[0].to("cuda"), chosen=chosen[0])
o = out.cpu()
print(f"threads {torch.get_num_threads()}  max_rel vs fp32 oracle {max_rel(o, ref):.3e}  vs fp64 oracle {max_rel(o, ref64):.3e}  fp32 oracle vs fp64 {max_rel(ref, ref64):.3e}")
d = (o.double() - ref64).abs()
flat = d.flatten().topk(3)
for v, i in zip(flat.values, flat.indices):
    b, r = divmod(int(i), C * Hh * Ww); c, r = divmod(r, Hh * Ww); y, xq = divmod(r, Ww)
    act = float(aux["pre"][b, 0, y, xq] * aux["fire"][b, 0, y, xq])
    print(f"   |d|={float(v):.2e} sample {b} ch {c} cell ({y},{xq}) act {act} ours {float(o[b,c,y,xq]):.6f} ref64 {float(ref64[b,c,y,xq]):.6f} u {float(aux['u'][b,c,y,xq]):.4f}")
mu = aux["u"].mean(dim=(1,2,3)); var = aux["u"].var(dim=(1,2,3), unbiased=False)
print("   rstd per sample:", " ".join(f"{float(1/(v+1e-3)**0.5):.1f}" for v in var))
# ---- gradient part of the test (2 steps), vs the fp64 oracle
xr = x.double().requires_grad_(True)
pr = {k: v.double().requires_grad_(v.dtype.is_floating_point) for k, v in p.items()}
s_ref = xr
for t in range(2):
    s_ref = O.nca_step(s_ref, pr, cfg, 0.5, fus[t].double(), chosen[t])
(s_ref[:, :4] ** 2).mean().backward()
xg = x.to("cuda").requires_grad_(True)
s_gpu = xg
for t in range(2):
    s_gpu = m.step(s_gpu, 0.5, fire_u=fus[t].to("cuda"), chosen=chosen[t])
(s_gpu[:, :4] ** 2).mean().backward()
print("NO_TC" if os.environ.get("GNCA_NO_TC") else "TC", f"state2 {rel_err(s_gpu.detach().cpu().double(), s_ref.detach()):.2e}  grad_x0 {rel_err(xg.grad.cpu().double(), xr.grad):.3e}",
      " per-sample:", " ".join(f"{rel_err(xg.grad[b].cpu().double(), xr.grad[b]):.1e}" for b in range(B)))
named = dict(m.named_parameters())
print("   ", " ".join(f"{n.split('.')[-2][-6:]}.{n.split('.')[-1][0]}={rel_err(named[n].grad.cpu().double(), pr[n].grad):.1e}" for n in
      ("update_net.0.weight", "update_net.0.bias", "update_net.2.weight", "norm.weight", "norm.bias", "graph.msg_proj.weight", "graph.msg_proj.bias")))
dg = (xg.grad[0].cpu().double() - xr.grad[0]).abs()
print("sample 0 grad: norm", float(xr.grad[0].norm()), "err norm", float((xg.grad[0].cpu().double() - xr.grad[0]).norm()))
top = dg.flatten().topk(8)
for v, i in zip(top.values, top.indices):
    c, r = divmod(int(i), Hh * Ww); y, xq = divmod(r, Ww)
    print(f"   |dg|={float(v):.2e} ch {c} cell ({y},{xq}) ours {float(xg.grad[0,c,y,xq]):.3e} ref {float(xr.grad[0,c,y,xq]):.3e}")
print("   cells with |dg| > 1e-9:", int((dg.amax(0) > 1e-9).sum()), "of", Hh * Ww)
