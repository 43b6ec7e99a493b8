import sys, os, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import graph_neural_cellular_automata_b200 as G
torch.manual_seed(1); random.seed(1)
for C, Hh, B in ((16, 128, 20), (32, 256, 5), (32, 64, 80)):
    Ww = Hh
    m = G.NeuralCAGraph(C, update_hidden=128, img_size=Hh, update_gain=0.1, alpha_thr=0.1, message_gain=0.3, hidden_only=True, graph_zero_padded_shift=False)
    with torch.no_grad():
        m.update_net[2].weight.normal_(0, 0.05)
        m.norm.weight.uniform_(0.5, 1.5); m.norm.bias.normal_(0, 0.1)
    m = m.cuda()
    yy, xx = torch.meshgrid(torch.arange(Hh), torch.arange(Ww), indexing="ij")
    disk = (((yy - Hh / 2) ** 2 + (xx - Ww / 2) ** 2) < (0.3 * Hh) ** 2).float()
    x = (torch.rand(B, C, Hh, Ww) * disk).cuda()
    fu = torch.rand(B, 1, Hh, Ww).cuda()
    chosen = random.sample(m.graph.offsets, 8)
    outs = []
    with torch.no_grad():
        for i in range(40):
            outs.append(m.step(x, 0.5, fire_u=fu, chosen=chosen).clone())
    torch.cuda.synchronize()
    ref = outs[-1]
    nd = [(int((o != ref).sum()), float((o - ref).abs().max())) for o in outs]
    print(f"C={C} {Hh}x{Hh} B={B}: runs differing from the last: {sum(1 for n, _ in nd if n)} of 40; first 6 (ncells, max|d|):", nd[:6])
    if nd[0][0]:
        d = (outs[0] - ref).abs()
        bad = (d.amax(1) > 0).nonzero()
        print("   differing cells in run 0 (b,y,x) first 12:", bad[:12].tolist(), " total", len(bad), " samples:", sorted(set(bad[:, 0].tolist())))
