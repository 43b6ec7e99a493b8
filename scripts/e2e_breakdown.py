"""Host-side breakdown of the e2e path of bench.py (C2): where the time outside the kernel goes."""
import os, sys, time, random
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import graph_neural_cellular_automata_b200 as G
from graph_neural_cellular_automata_b200.rollout import make_schedule, rollout
from graph_neural_cellular_automata_b200.utils.nca_init import make_seed
m = G.NeuralCAGraph(16, update_hidden=128, img_size=40, update_gain=0.05, alpha_thr=0.12, message_gain=0.25,
                    graph_zero_padded_shift=False)
m.load_state_dict({k: torch.from_numpy(v) for k, v in np.load(os.path.join(ROOT, "tests/golden/weights_graph_ep960.npz")).items()}, strict=False)
m = m.cuda()
x0h = make_seed(16, 40, 8).pin_memory(); xTh = torch.empty_like(x0h).pin_memory()
def T(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
s = make_schedule(m, 8, 40, 40, 96, fire_rate=0.5, seed=1)
x0 = x0h.cuda()
print("make_schedule (incl. its H2D)   %.3f ms" % T(lambda: make_schedule(m, 8, 40, 40, 96, fire_rate=0.5, seed=1)))
print("draw_offsets_array only          %.3f ms" % T(lambda: m.graph.draw_offsets_array(96)))
print("h2d x0                           %.3f ms" % T(lambda: x0h.to("cuda", non_blocking=True)))
with torch.no_grad():
    print("rollout (device resident, sync)  %.3f ms" % T(lambda: rollout(m, x0, s)))
    def full():
        x = x0h.to("cuda", non_blocking=True); sc = make_schedule(m, 8, 40, 40, 96, fire_rate=0.5, seed=2)
        y = rollout(m, x, sc); xTh.copy_(y, non_blocking=True); torch.cuda.synchronize()
    print("full e2e                         %.3f ms" % T(full))
    def launch_only():
        t0 = time.perf_counter(); rollout(m, x0, s); return time.perf_counter() - t0
    torch.cuda.synchronize(); ls = [launch_only() for _ in range(20)]; torch.cuda.synchronize()
    print("rollout() host time to return    %.3f ms" % (np.median(ls) * 1e3))
    import cProfile, pstats
    for _ in range(20): full()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(200): full()
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(28)
