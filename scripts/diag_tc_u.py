import sys, os, random, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import graph_neural_cellular_automata_b200 as G
from graph_neural_cellular_automata_b200 import functional as GF, _lib
from oracle import nca_oracle as O
torch.manual_seed(1); random.seed(1)
Cc, Hh, Ww, B, hid = 16, 128, 128, 20, 128
m = G.NeuralCAGraph(Cc, update_hidden=hid, img_size=Hh, update_gain=0.1, alpha_thr=0.1, message_gain=0.3, hidden_only=True, graph_zero_padded_shift=False)
with torch.no_grad():
    m.update_net[2].weight.normal_(0, 0.05)
    m.norm.weight.uniform_(0.5, 1.5); m.norm.bias.normal_(0, 0.1)
p64 = {k: v.detach().clone().double() for k, v in m.state_dict().items()}
m = m.to("cuda")
yy, xx = torch.meshgrid(torch.arange(Hh), torch.arange(Ww), indexing="ij")
disk = (((yy - Hh / 2) ** 2 + (xx - Ww / 2) ** 2) < (0.3 * Hh) ** 2).float()
x = torch.rand(B, Cc, Hh, Ww) * disk
x[B // 2:, 3] *= (torch.rand(B - B // 2, Hh, Ww) > 0.5).float()
fus = [torch.rand(B, 1, Hh, Ww) for _ in range(2)]
chosen = [random.sample(m.graph.offsets, 8) for _ in range(2)]
cfg = O.StepConfig(update_gain=0.1, alpha_thr=0.1, graph=True, message_gain=0.3, hidden_only=True, zero_padded_shift=False)

def gpu_step(xin, t):
    desc, packed = m.model_desc(), m.packed_weights()
    lib = _lib.load()
    arr, k = GF._offsets_array(chosen[t])
    out = torch.empty_like(xin); u = torch.zeros_like(xin); stats = torch.empty(B, 2, device="cuda")
    ws = GF._WS.get(xin.device, lib.gnca_step_workspace_bytes(C.byref(desc), B, Hh, Ww))
    fu = fus[t].cuda()
    _lib.check(lib.gnca_step_fwd(C.byref(desc), GF._ptr(packed), B, Hh, Ww, GF._ptr(xin), GF._ptr(out), GF._ptr(fu), 0.5, arr, k, 0.3,
                                 GF._ptr(u), GF._ptr(stats), None, GF._ptr(ws), ws.numel(), GF._stream()), "step")
    return out, u, stats

xin = x.cuda()
for t in range(2):
    out, u, stats = gpu_step(xin, t)
    ref, aux = O.nca_step(xin.cpu().double(), p64, cfg, 0.5, fus[t].double(), chosen[t], return_aux=True)
    act = (aux["pre"] * aux["fire"]).bool().expand_as(u)
    du = ((u.cpu().double() - aux["u"]).abs() * act)
    mu = aux["u"].mean(dim=(1, 2, 3)); var = aux["u"].var(dim=(1, 2, 3), unbiased=False)
    print(f"step {t}: max|du| per sample:", " ".join(f"{float(v):.1e}" for v in du.amax(dim=(1, 2, 3))))
    print(f"         mean err per sample:", " ".join(f"{abs(float(stats[b,0]) - float(mu[b])):.1e}" for b in range(B)))
    print(f"         rstd relerr per sample:", " ".join(f"{abs(float(stats[b,1]) * float((var[b]+1e-3)**0.5) - 1):.1e}" for b in range(B)))
    print(f"         out max|d| per sample:", " ".join(f"{float(v):.1e}" for v in (out.cpu().double() - ref).abs().amax(dim=(1, 2, 3))))
    xin = out
