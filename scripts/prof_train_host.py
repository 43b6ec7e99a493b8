"""Host-side profile of GraphNCATrainer.train_step (fast mode) on one GPU: cProfile over 40 steps + wall / device split."""
import cProfile, pstats, io, os, sys, random, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import graph_neural_cellular_automata_b200 as G
from graph_neural_cellular_automata_b200.training.trainer import GraphNCATrainer, TrainConfig
torch.manual_seed(1234); random.seed(1234)
m = G.NeuralCAGraph(16, update_hidden=128, img_size=40, update_gain=0.05, alpha_thr=0.12, message_gain=0.25, hidden_only=True, graph_zero_padded_shift=False)
m.load_state_dict({k: torch.from_numpy(v) for k, v in np.load(os.path.join(ROOT, "tests/golden/weights_graph_ep960.npz")).items()}, strict=False)
m = m.cuda()
target = torch.from_numpy(np.load(os.path.join(ROOT, "tests/golden/target_gecko_surrogate.npy"))).cuda()
tr = GraphNCATrainer(m, target, TrainConfig(batch_size=32, pool_size=1024, long_rollout_prob=0.0, fire="philox"))
for _ in range(5): tr.train_step(epoch=300)
torch.cuda.synchronize()
# host time only: how long until train_step RETURNS (device work still queued)
t0 = time.perf_counter(); ret = []
for _ in range(40):
    a = time.perf_counter(); tr.train_step(epoch=300); ret.append(time.perf_counter() - a)
torch.cuda.synchronize(); wall = time.perf_counter() - t0
print("wall per step %.3f ms; host return time per step: median %.3f ms, min %.3f, max %.3f" % (wall / 40 * 1e3, np.median(ret) * 1e3, min(ret) * 1e3, max(ret) * 1e3))
pr = cProfile.Profile(); pr.enable()
for _ in range(40): tr.train_step(epoch=300)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:5000])
