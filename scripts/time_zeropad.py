"""Zero-padded-shift (module default) forward rollout, C2 shape: resident (ZP instantiation of k_rep_fwd) vs streaming."""
import os, sys, random, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_step import graph_model
from oracle.nca_oracle import make_seed
from graph_neural_cellular_automata_b200.rollout import make_schedule, rollout
for torus in (False, True):
    m = graph_model(torus)
    x0 = make_seed(16, 40, 8).cuda()
    random.seed(1)
    sched = make_schedule(m, 8, 40, 40, 96, fire_rate=0.5, seed=5)
    for impl in ("resident", "streaming"):
        with torch.no_grad():
            for _ in range(3): rollout(m, x0, sched, impl=impl)
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): rollout(m, x0, sched, impl=impl)
            e1.record(); torch.cuda.synchronize()
        print(f"{'torus' if torus else 'zero-pad'} {impl:10s} {e0.elapsed_time(e1) / 10:.3f} ms per rollout (B=8, T=96, 40x40x16)")
