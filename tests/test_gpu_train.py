"""GPU parity of the training iteration (trainer.py) against the reference loop body restated on the CPU oracle:
same seeded RNG streams -> same pool indices / step counts / fire rates / offsets / masks; per-sample loss 1e-5,
normalised gradients 1e-4, Adam-updated parameters, worst-k indices bit-exact, pool contents."""
import random

import numpy as np
import pytest
import torch

from conftest import load_params, rel_err
from oracle import nca_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import graph_neural_cellular_automata_b200 as G
    from graph_neural_cellular_automata_b200.training.trainer import (GraphNCATrainer, TrainConfig, premult_loss,
                                                                      scheduled_message_gain)
    from graph_neural_cellular_automata_b200.training.optim import FusedNormalizedAdam
    from graph_neural_cellular_automata_b200.rollout import make_schedule, rollout_fwd_raw, rollout_bwd_raw
    from graph_neural_cellular_automata_b200.utils.nca_init import trainer_seed
    from test_gpu_step import graph_model, T32, DEV
    import os
    from conftest import GOLDEN


def _target():
    return T32(np.load(os.path.join(GOLDEN, "target_gecko_surrogate.npy")))


def _oracle_iteration(params, cfg, pool_cpu, target, epoch, offsets_all):
    """train_graph_augmented_nca.py:289-391 on the CPU oracle, drawing randomness exactly like the reference
    (device draws happen on the CUDA generator, as they would for the reference running on this GPU)."""
    B = cfg.batch_size
    idx = random.sample(range(pool_cpu.shape[0]), B)
    state = pool_cpu[idx].clone()
    if random.random() < cfg.long_rollout_prob:
        lo, hi = cfg.long_rollout_steps_min, cfg.long_rollout_steps_max
    else:
        lo, hi = cfg.nca_steps_min, cfg.nca_steps_max
    nca_steps = torch.randint(lo, hi + 1, (B,), device=DEV).cpu()
    T = int(nca_steps.max())
    base = scheduled_message_gain(epoch, cfg.message_gain)
    p = {k: v.clone().requires_grad_(v.is_floating_point() and "perception" not in k and "gate_mlp" not in k)
         for k, v in params.items()}
    oc = O.StepConfig(update_gain=0.05, alpha_thr=0.12, graph=True, message_gain=base, hidden_only=True,
                      zero_padded_shift=False)
    x = state
    for t in range(T):
        mask = nca_steps > t
        fr = float(torch.empty(1, device=DEV).uniform_(cfg.fire_rate_min, cfg.fire_rate_max).item())
        use_graph = (t % cfg.message_every == 0) if cfg.message_every > 1 else True
        c = O.StepConfig(**{**oc.__dict__, "message_gain": base if use_graph else 0.0})
        chosen = random.sample(offsets_all, 8)
        fu = torch.rand(int(mask.sum()), 1, 40, 40, device=DEV).cpu()
        new = O.nca_step(x[mask], p, c, fr, fu, chosen)
        x = x.clone()
        x[mask] = new
    per = O.loss_premult_rgba(x[:, :4], target.unsqueeze(0).expand(B, -1, -1, -1))
    per.mean().backward()
    grads = {k: v.grad for k, v in p.items() if v.requires_grad}
    return idx, nca_steps, x.detach(), per.detach(), grads


def test_train_step_matches_reference_loop():
    torch.manual_seed(11); random.seed(11)
    m = graph_model(True)
    cfg = TrainConfig(batch_size=4, pool_size=16, nca_steps_min=6, nca_steps_max=9, long_rollout_prob=0.0,
                      fire_rate_min=0.5, fire_rate_max=0.9, learning_rate=2e-4, weight_decay=1e-5,
                      reset_worst_prob=0.25, random_reseed_prob=0.0, message_gain=0.25, message_every=3, damage={},
                      fire="torch", rollout_impl="auto")
    # a pool of grown states (so the rollout has live cells): grow the trainer seeds a little with the model
    tr = GraphNCATrainer(m, _target(), cfg)
    with torch.no_grad():
        x = tr.pool.pool
        for _ in range(24):
            x = m(x, fire_rate=0.7)
        tr.pool.pool.copy_(x)
    params0 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    pool0 = tr.pool.pool.detach().cpu().clone()
    st_py, st_cuda = random.getstate(), torch.cuda.get_rng_state()
    out = tr.train_step(epoch=1)
    params1 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    pool1 = tr.pool.pool.detach().cpu().clone()
    # ---- replay on the oracle ----
    random.setstate(st_py); torch.cuda.set_rng_state(st_cuda)
    idx, nca_steps, xT, per, grads = _oracle_iteration(params0, cfg, pool0, _target(), 1, m.graph.offsets)
    assert np.array_equal(out["steps"], nca_steps.numpy())
    assert rel_err(out["per_sample"].cpu(), per) < 1e-5
    worst = torch.topk(per, int(cfg.reset_worst_prob * cfg.batch_size)).indices
    assert torch.equal(out["worst"].cpu(), worst)                          # bit-exact integer work
    # normalised gradient + Adam on the oracle side
    pr = {k: torch.nn.Parameter(params0[k].clone()) for k in grads}
    for k in grads:
        g = grads[k] if grads[k] is not None else torch.zeros_like(pr[k])
        pr[k].grad = g / (g.norm() + 1e-8)
    torch.optim.Adam(list(pr.values()), lr=cfg.learning_rate, weight_decay=cfg.weight_decay).step()
    for k in grads:
        if any(s in k for s in ("query_proj", "key_proj", "scaling")):
            continue      # torus: true gradient is 0; the reference normalises rounding noise (documented)
        d_ref = pr[k].detach() - params0[k]
        d_ours = params1[k] - params0[k]
        # step 1 of Adam moves every element by ~lr*sign(g): compare where the gradient is not ~0
        g = pr[k].grad
        big = g.abs() > 1e-5
        assert float((d_ours - d_ref)[big].abs().max()) < 0.02 * cfg.learning_rate, k
        assert rel_err(d_ours, d_ref) < 2e-2, (k, rel_err(d_ours, d_ref))
    # pool: sampled slots replaced by the rolled states, worst ones reseeded (alpha 1 at the centre, nothing else alive)
    keep = [i for j, i in enumerate(idx) if j not in worst.tolist()]
    keep_pos = [j for j in range(len(idx)) if j not in worst.tolist()]
    assert rel_err(pool1[keep], xT[keep_pos]) < 1e-5
    for j in worst.tolist():
        s = pool1[idx[j]]
        assert float(s[3, 20, 20]) == 1.0 and float(s[3].sum()) == 1.0
    untouched = [i for i in range(cfg.pool_size) if i not in idx]
    assert torch.equal(pool1[untouched], pool0[untouched])


def test_gradient_is_linear_in_the_batch():
    """DP property: grad(full batch) == sum of grads of its shards (what the SUM all-reduce relies on)."""
    torch.manual_seed(3); random.seed(3)
    m = graph_model(True)
    x0 = T32(__import__("conftest").load_golden("graph_torus_step.npz")["x_in"]).to(DEV).repeat(2, 1, 1, 1)   # B=4
    T = 5
    offs = [m.graph.draw_offsets() for _ in range(T)]
    fu = torch.rand(T, 4, 40, 40, device=DEV)
    tgt = _target().to(DEV)
    desc, packed = m.model_desc(), m.packed_weights()

    def grad_of(sl):
        s = make_schedule(m, sl.stop - sl.start, 40, 40, T, fire_rate=0.6, offsets=offs, fire_u=fu[:, sl].contiguous())
        xT, hist = rollout_fwd_raw(desc, packed, x0[sl].contiguous(), s, history=True)
        per, g = premult_loss(xT, tgt, 1.0 / 4)
        _, gflat = rollout_bwd_raw(desc, packed, hist, s, g)
        return per, gflat
    per_full, g_full = grad_of(slice(0, 4))
    per_a, g_a = grad_of(slice(0, 2))
    per_b, g_b = grad_of(slice(2, 4))
    assert torch.equal(torch.cat([per_a, per_b]), per_full)
    assert rel_err((g_a + g_b).cpu(), g_full.cpu()) < 1e-5


def test_fused_adam_matches_torch():
    torch.manual_seed(0)
    m = graph_model(True)
    ref = {k: torch.nn.Parameter(v.detach().cpu().clone()) for k, v in m.named_parameters()}
    opt = FusedNormalizedAdam(m, lr=1e-3, weight_decay=1e-5, normalize=True)
    topt = None
    names = ["update_net.0.weight", "update_net.0.bias", "update_net.2.weight", "norm.weight", "norm.bias",
             "graph.msg_proj.weight", "graph.msg_proj.bias", "graph.query_proj.weight", "graph.query_proj.bias",
             "graph.key_proj.weight", "graph.key_proj.bias", "graph.scaling"]
    topt = torch.optim.Adam([ref[n] for n in names], lr=1e-3, weight_decay=1e-5)
    for it in range(3):
        gflat = torch.randn(opt.flat.numel(), device=DEV)
        gcpu = gflat.cpu()
        for i, n in enumerate(names):
            g = gcpu[opt.seg[i]:opt.seg[i + 1]].view(ref[n].shape).clone()
            ref[n].grad = g / (g.norm() + 1e-8)
        topt.step()
        opt.step(gflat)
    named = dict(m.named_parameters())
    for n in names:
        assert rel_err(named[n].detach().cpu(), ref[n].detach()) < 1e-6, n
    # parameters are views of the flat buffer and the packed-weight cache was invalidated
    assert named["update_net.0.weight"].data_ptr() == opt.flat.data_ptr()
