"""world_size-2 gloo test (CPU) of the data-parallel host logic: batch sharding, the SUM all-reduce of the flat
gradient, the all-gather of per-sample losses and the bit-exact global worst-k indices."""
import os
import socket
import tempfile

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from graph_neural_cellular_automata_b200.training.dp import Shard, worst_k_indices
    torch.manual_seed(7)                                  # same seeded host RNG on every rank
    B = 8
    per_global_truth = torch.rand(B)
    grads_per_sample = torch.randn(B, 9169)
    sh = Shard(B)
    assert (sh.rank, sh.world, sh.local_batch) == (rank, world, B // world)
    g_local = sh.take(grads_per_sample).sum(0).contiguous()
    g = sh.allreduce_sum_(g_local.clone())
    per = sh.allgather(sh.take(per_global_truth).contiguous())
    worst = worst_k_indices(per, 0.25)
    states = sh.allgather(sh.take(torch.arange(B * 6, dtype=torch.float32).view(B, 2, 3)).contiguous())
    torch.save((g, per, worst, states, grads_per_sample.sum(0), per_global_truth), os.path.join(out, f"{rank}.pt"))   # files, not a forked Manager
    dist.barrier()
    dist.destroy_process_group()


def test_dp_two_ranks_gloo():
    world = 2
    out = tempfile.mkdtemp(prefix="gnca_dp_")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    g0, per0, w0, st0, gsum, truth = torch.load(os.path.join(out, "0.pt"))
    g1, per1, w1, st1, _, _ = torch.load(os.path.join(out, "1.pt"))
    assert torch.equal(g0, g1)                                        # identical on every rank
    assert torch.allclose(g0, gsum, rtol=1e-6, atol=1e-6)            # == single-process sum over the batch
    assert torch.equal(per0, truth) and torch.equal(per1, truth)
    assert torch.equal(w0, w1) and torch.equal(w0, torch.topk(truth, 2).indices)   # bit-exact global top-k
    assert torch.equal(st0, torch.arange(48, dtype=torch.float32).view(8, 2, 3)) and torch.equal(st0, st1)


def test_shard_rejects_indivisible_batch():
    from graph_neural_cellular_automata_b200.training.dp import Shard
    import pytest
    with pytest.raises(ValueError):
        Shard(7, rank=0, world=2)
    s = Shard(8, rank=1, world=4)
    assert (s.lo, s.hi) == (2, 4)
