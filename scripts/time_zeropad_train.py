"""Training iteration of the MODULE DEFAULT graph shift (zero-padded) vs the trainer's torus: ms per iteration and the
library's per-kernel event timings (ids of include/gnca.h: 0 k_update, 1 k_apply, 2 resident fwd, 3 k_bwd_mlp / wgrad, 6 resident bwd)."""
import ctypes, os, sys, random, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import graph_neural_cellular_automata_b200 as G
from graph_neural_cellular_automata_b200 import _lib
from graph_neural_cellular_automata_b200.training.trainer import GraphNCATrainer, TrainConfig
lib = _lib.load()
target = torch.from_numpy(np.load(os.path.join(ROOT, "tests/golden/target_gecko_surrogate.npy"))).cuda()
QUICK = bool(os.environ.get("ZP_ONLY"))          # under ncu: the zero-pad model only, few iterations
for zp in ((os.environ["ZP_ONLY"] != "torus",) if QUICK else (False, True)):     # ZP_ONLY=torus: the torus model only
    torch.manual_seed(1234); random.seed(1234)
    m = G.NeuralCAGraph(16, update_hidden=128, img_size=40, update_gain=0.05, alpha_thr=0.12, message_gain=0.25, hidden_only=True,
                        graph_zero_padded_shift=zp)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in np.load(os.path.join(ROOT, "tests/golden/weights_graph_ep960.npz")).items()}, strict=False)
    m = m.cuda()
    tr = GraphNCATrainer(m, target, TrainConfig(batch_size=32, pool_size=1024, long_rollout_prob=0.0, fire="philox"))
    for _ in range(1 if QUICK else 4): tr.train_step(epoch=300)
    torch.cuda.synchronize()
    n = 1 if QUICK else 10
    t0 = time.perf_counter()
    for _ in range(n): out = tr.train_step(epoch=300)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / n * 1e3
    lib.gnca_profile_enable(1)
    for _ in range(3): tr.train_step(epoch=300)
    torch.cuda.synchronize(); lib.gnca_profile_enable(0)
    if QUICK: break
    parts = []
    for kid in range(8):
        a, c = ctypes.c_double(0), ctypes.c_ulonglong(0)
        lib.gnca_profile_read(kid, ctypes.byref(a), ctypes.byref(c))
        if c.value: parts.append("id%d %.2f ms/%d" % (kid, a.value / 3, c.value // 3))
    print("zero_padded_shift=%s: %.2f ms per iteration | per iteration: %s" % (zp, ms, ", ".join(parts)), flush=True)
