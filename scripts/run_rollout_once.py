"""Run the C2 rollout a few times (for ncu captures): python scripts/run_rollout_once.py [n] [impl]"""
import os, sys, random
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import graph_neural_cellular_automata_b200 as G
from graph_neural_cellular_automata_b200.rollout import make_schedule, rollout
from graph_neural_cellular_automata_b200.utils.nca_init import make_seed
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
impl = sys.argv[2] if len(sys.argv) > 2 else "auto"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
torch.manual_seed(0); random.seed(0)
m = G.NeuralCAGraph(16, update_hidden=128, img_size=40, update_gain=0.05, alpha_thr=0.12, message_gain=0.25,
                    graph_zero_padded_shift=False)
m.load_state_dict({k: torch.from_numpy(v) for k, v in np.load(os.path.join(ROOT, "tests/golden/weights_graph_ep960.npz")).items()}, strict=False)
m = m.cuda()
x0 = make_seed(16, 40, B, device="cuda")
with torch.no_grad():
    for i in range(n):
        s = make_schedule(m, B, 40, 40, 96, fire_rate=0.5, seed=100 + i)
        y = rollout(m, x0, s, impl=impl)
torch.cuda.synchronize()
print("alive fraction at T:", float((y[:, 3] > 0.12).float().mean()))
