"""40x40x32 forward rollout (B=8, T=96): banded cluster kernel vs the streaming kernels."""
import os, sys, random, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import graph_neural_cellular_automata_b200 as G
from graph_neural_cellular_automata_b200.rollout import make_schedule, rollout
torch.manual_seed(5); random.seed(5)
C, H, W, B, T = 32, 40, 40, 8, 96
m = G.NeuralCAGraph(C, update_hidden=128, img_size=H, update_gain=0.1, alpha_thr=0.1, message_gain=0.3, hidden_only=True, graph_zero_padded_shift=False)
with torch.no_grad():
    m.update_net[2].weight.normal_(0, 0.05)
m = m.cuda()
yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
x0 = (torch.rand(B, C, H, W) * (((yy - 19) ** 2 + (xx - 21) ** 2) < 81).float()).cuda()
for impl in ("resident", "streaming"):
    s = make_schedule(m, B, H, W, T, fire_rate=0.5, seed=1)
    with torch.no_grad():
        for _ in range(3): rollout(m, x0, s, impl=impl)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10): rollout(m, x0, s, impl=impl)
        torch.cuda.synchronize()
    print("%s: %.3f ms per rollout (B=%d T=%d %dx%dx%d)" % (impl, (time.perf_counter() - t0) / 10 * 1e3, B, T, H, W, C), flush=True)
