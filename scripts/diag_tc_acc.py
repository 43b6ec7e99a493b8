"""One-step accuracy of the balanced k_update path (tensor-core or, with GNCA_NO_TC=1, FFMA) against the fp64 oracle."""
import sys, os, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import graph_neural_cellular_automata_b200 as G
from oracle import nca_oracle as O

def run(C, Hh, B, graph=True, gain=0.3):
    torch.manual_seed(1); random.seed(1)
    m = G.NeuralCAGraph(C, update_hidden=128, img_size=Hh, update_gain=0.1, alpha_thr=0.1, message_gain=gain,
                        hidden_only=True, graph_zero_padded_shift=False)
    with torch.no_grad():
        m.update_net[2].weight.normal_(0, 0.05)
        m.norm.weight.uniform_(0.5, 1.5); m.norm.bias.normal_(0, 0.1)
    p = {k: v.detach().clone().double() for k, v in m.state_dict().items()}
    m = m.cuda()
    yy, xx = torch.meshgrid(torch.arange(Hh), torch.arange(Hh), indexing="ij")
    disk = (((yy - Hh / 2) ** 2 + (xx - Hh / 2) ** 2) < (0.3 * Hh) ** 2).float()
    x = torch.rand(B, C, Hh, Hh) * disk
    fu = torch.rand(B, 1, Hh, Hh)
    chosen = random.sample(m.graph.offsets, 8)
    cfg = O.StepConfig(update_gain=0.1, alpha_thr=0.1, graph=True, message_gain=gain, hidden_only=True, zero_padded_shift=False)
    nb = min(B, 4)
    ref, aux = O.nca_step(x[:nb].double(), p, cfg, 0.5, fu[:nb].double(), chosen, return_aux=True)
    with torch.no_grad():
        out = m.step(x.cuda(), 0.5, fire_u=fu.cuda(), chosen=chosen)
    d = (out[:nb].cpu().double() - ref).abs()
    per_c = d.amax(dim=(0, 2, 3))
    u = aux["u"]
    print(f"C={C} {Hh}x{Hh} B={B} gain={gain}: max|d|={float(d.max()):.2e}  u std={float(u.std()):.3f} max|u|={float(u.abs().max()):.2f}  "
          f"per-channel max|d|: " + " ".join(f"{float(v):.0e}" for v in per_c[:8]))

print("GNCA_NO_TC =", os.environ.get("GNCA_NO_TC"))
run(16, 128, 20); run(16, 128, 20, gain=0.0); run(32, 128, 20); run(16, 256, 5); run(32, 256, 5); run(32, 64, 80)
