"""Trainer-level data-parallel equivalence on ONE GPU: the real `GraphNCATrainer.train_step` run by two processes on
cuda:0 (gloo backend, world 2, global batch 8) against one process (world 1), SURVEY section 4 "N-GPU equals 1-GPU":
the all-reduced flat gradient (<= 1e-6 rel), the per-sample losses, the worst-k reset indices (bit-exact), the updated
parameters and the whole replicated pool after two iterations (with the damage policy on).  fire="torch": every rank
draws the global batch's uniforms from the same seeded stream and takes its slice, so the two worlds see identical masks.
(The ranks never wait for each other on the device: the collectives are host-side, staged through gloo.)"""
import os
import random
import socket
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, load_params, rel_err

pytestmark = pytest.mark.gpu

DMG = {"start_epoch": 100, "prob": 1.0, "kinds": {"square": 0.4, "circle": 0.3, "gaussian": 0.3}, "size_min": 6, "size_max": 12,
       "gaussian_softness": 0.35}


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _run(rank, world, port, out, sharding="replicated"):
    if world > 1:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
    import graph_neural_cellular_automata_b200 as G
    from graph_neural_cellular_automata_b200.training.trainer import GraphNCATrainer, TrainConfig
    torch.cuda.set_device(0)
    torch.manual_seed(99); random.seed(99)
    m = G.NeuralCAGraph(16, update_hidden=128, img_size=40, update_gain=0.05, alpha_thr=0.12, message_gain=0.25,
                        hidden_only=True, graph_zero_padded_shift=False)
    m.load_state_dict(load_params("weights_graph_ep960.npz"), strict=False)
    m = m.cuda()
    target = torch.from_numpy(np.load(os.path.join(GOLDEN, "target_gecko_surrogate.npy"))).cuda()
    cfg = TrainConfig(batch_size=8, pool_size=32, nca_steps_min=10, nca_steps_max=16, long_rollout_prob=0.0, fire="torch",
                      reset_worst_prob=0.25, random_reseed_prob=1.0, damage=DMG, pool_sharding=sharding)
    tr = GraphNCATrainer(m, target, cfg)
    rec = []
    for it in range(2):
        o = tr.train_step(epoch=150)
        rec.append({"per": o["per_sample"].cpu(), "worst": o["worst"].cpu(), "gflat": o["gflat"].cpu(), "steps": o["steps"].copy()})
    torch.cuda.synchronize()
    # results go through files: a multiprocessing.Manager would FORK a server out of the (multi-threaded, CUDA-initialised)
    # pytest process, and that server aborted once inside a garbage collection
    torch.save({"rec": rec, "flat": tr.opt.flat.cpu(), "pool": tr.pool.pool.cpu()}, os.path.join(out, f"{world}_{rank}_{sharding}.pt"))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _load(out, world, rank, sharding):
    return torch.load(os.path.join(out, f"{world}_{rank}_{sharding}.pt"), weights_only=False)


def test_two_ranks_equal_one_rank():
    out = tempfile.mkdtemp(prefix="gnca_dp_")
    mp.spawn(_run, args=(1, 0, out), nprocs=1, join=True)
    mp.spawn(_run, args=(2, _free_port(), out), nprocs=2, join=True)
    one, r0, r1 = _load(out, 1, 0, "replicated"), _load(out, 2, 0, "replicated"), _load(out, 2, 1, "replicated")
    for it in range(2):
        a, b, c = one["rec"][it], r0["rec"][it], r1["rec"][it]
        assert np.array_equal(a["steps"], b["steps"]) and np.array_equal(a["steps"], c["steps"])     # same host RNG replay
        assert torch.equal(b["gflat"], c["gflat"])                                                   # all-reduce: identical on both ranks
        assert rel_err(b["gflat"], a["gflat"]) < 1e-6, (it, rel_err(b["gflat"], a["gflat"]))        # == the 1-GPU gradient
        assert torch.equal(b["per"], c["per"]) and rel_err(b["per"], a["per"]) < 1e-6
        assert torch.equal(a["worst"], b["worst"]) and torch.equal(a["worst"], c["worst"])           # bit-exact global worst-k
    assert torch.equal(r0["flat"], r1["flat"]) and rel_err(r0["flat"], one["flat"]) < 1e-6           # parameters after 2 Adam steps
    assert torch.equal(r0["pool"], r1["pool"])                                                       # replicated pool stays replicated
    assert rel_err(r0["pool"], one["pool"]) < 1e-5


def test_owner_sharded_pool_two_ranks():
    """pool_sharding="owner" (rank r owns pool_size / world slots, no state all-gather): both ranks agree on the
    all-reduced gradient, the global per-sample losses, the worst-k indices and the parameters; each rank's pool shard has
    pool_size / world slots, changed only in the slots it drew, with its members of the global worst-k set reseeded."""
    out = tempfile.mkdtemp(prefix="gnca_dp_")
    mp.spawn(_run, args=(2, _free_port(), out, "owner"), nprocs=2, join=True)
    r0, r1 = _load(out, 2, 0, "owner"), _load(out, 2, 1, "owner")
    assert r0["pool"].shape[0] == 16 and r1["pool"].shape[0] == 16
    for it in range(2):
        b, c = r0["rec"][it], r1["rec"][it]
        assert torch.equal(b["gflat"], c["gflat"]) and torch.equal(b["per"], c["per"]) and torch.equal(b["worst"], c["worst"])
        assert b["per"].numel() == 8 and bool(torch.isfinite(b["gflat"]).all()) and float(b["gflat"].abs().max()) > 0
    assert torch.equal(r0["flat"], r1["flat"])
    assert bool(torch.isfinite(r0["pool"]).all()) and bool(torch.isfinite(r1["pool"]).all())
    # two iterations of 4 local samples each touched at most 8 of the 16 slots of a shard; the others still hold seeds
    for r in (r0, r1):
        untouched = int(((r["pool"][:, :3].abs().sum(dim=(1, 2, 3)) == 0) & (r["pool"][:, 3].sum(dim=(1, 2)) == 1)).sum())
        assert untouched >= 8
