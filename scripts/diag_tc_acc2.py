import sys, os, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import graph_neural_cellular_automata_b200 as G
from oracle import nca_oracle as O
torch.manual_seed(1); random.seed(1)
C, Hh, Ww, B, hid = 16, 128, 128, 20, 128
m = G.NeuralCAGraph(C, update_hidden=hid, img_size=Hh, update_gain=0.1, alpha_thr=0.1, message_gain=0.3, hidden_only=True, graph_zero_padded_shift=False)
with torch.no_grad():
    m.update_net[2].weight.normal_(0, 0.05)
    m.norm.weight.uniform_(0.5, 1.5); m.norm.bias.normal_(0, 0.1)
p = {k: v.detach().clone() for k, v in m.state_dict().items()}
m = m.cuda()
yy, xx = torch.meshgrid(torch.arange(Hh), torch.arange(Ww), indexing="ij")
disk = (((yy - Hh / 2) ** 2 + (xx - Ww / 2) ** 2) < (0.3 * Hh) ** 2).float()
x = torch.rand(B, C, Hh, Ww) * disk
x[B // 2:, 3] *= (torch.rand(B - B // 2, Hh, Ww) > 0.5).float()
fus = [torch.rand(B, 1, Hh, Ww) for _ in range(2)]
chosen = [random.sample(m.graph.offsets, 8) for _ in range(2)]
cfg = O.StepConfig(update_gain=0.1, alpha_thr=0.1, graph=True, message_gain=0.3, hidden_only=True, zero_padded_shift=False)
ref, aux = O.nca_step(x, p, cfg, 0.5, fus[0], chosen[0], return_aux=True)
with torch.no_grad():
    out = m.step(x.cuda(), 0.5, fire_u=fus[0].cuda(), chosen=chosen[0])
d = (out.cpu() - ref).abs()
print("per-sample max|d|:", " ".join(f"{float(v):.1e}" for v in d.amax(dim=(1, 2, 3))))
b = int(d.amax(dim=(1, 2, 3)).argmax())
idx = d[b].flatten().argmax(); c, yx = divmod(int(idx), Hh * Ww); y, xq = divmod(yx, Ww)
print("worst: sample", b, "channel", c, "cell", (y, xq), "ours", float(out[b, c, y, xq]), "ref", float(ref[b, c, y, xq]),
      "pre", float(aux["pre"][b, 0, y, xq]), "fire", float(aux["fire"][b, 0, y, xq]), "u", float(aux["u"][b, c, y, xq]))
nbad = (d[b] > 1e-6).sum(dim=(1, 2))
print("cells with |d|>1e-6 per channel in that sample:", nbad.tolist())
bad = (d[b].amax(0) > 1e-6).nonzero()
print("bad cells (first 10):", bad[:10].tolist(), "of", len(bad))
u = aux["u"][b]; print("sample stats: u mean", float(u.mean()), "var", float(u.var(unbiased=False)), "active cells", int((aux["pre"][b] * aux["fire"][b]).sum()))
