"""GPU parity of the classic trainer's iteration WITH its stability phase (training/classic_trainer.py, reference
train_intermediate_loss.py:230-296) against the same loop restated on the CPU oracle: same seeded RNG streams -> same
pool indices / step counts / fire rates / fire masks; per-sample loss, close-mask, stability loss, clipped-gradient Adam
update, worst-k indices and pool contents."""
import os
import random

import numpy as np
import pytest
import torch

from conftest import GOLDEN, load_params, rel_err
from oracle import nca_oracle as O

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from graph_neural_cellular_automata_b200.training.classic_trainer import (ClassicNCATrainer, ClassicTrainConfig,
                                                                              masked_loss)
    from graph_neural_cellular_automata_b200.rollout import make_schedule, rollout
    from graph_neural_cellular_automata_b200.utils.nca_init import trainer_seed
    from test_gpu_step import classic_model, T32, DEV


def _target():
    return T32(np.load(os.path.join(GOLDEN, "target_gecko_surrogate.npy")))


def _oracle_iteration(params, cfg, pool_cpu, target):
    """the reference loop body on the CPU oracle; device draws on the CUDA generator, like the reference on this GPU"""
    B = cfg.batch_size
    idx = random.sample(range(pool_cpu.shape[0]), B)
    state = pool_cpu[idx].clone()
    lo, hi = (cfg.long_min, cfg.long_max) if random.random() < cfg.long_prob else (cfg.nca_steps_min, cfg.nca_steps_max)
    nca_steps = torch.randint(lo, hi + 1, (B,), device=DEV).cpu()
    p = {k: v.clone().requires_grad_(v.is_floating_point() and "perception" not in k) for k, v in params.items()}
    oc = O.StepConfig(update_gain=0.1, alpha_thr=0.1, graph=False)
    x = state
    for t in range(int(nca_steps.max())):
        mask = nca_steps > t
        fr = float(torch.empty(1, device=DEV).uniform_(cfg.fire_rate_min, cfg.fire_rate_max).item())
        fu = torch.rand(int(mask.sum()), 1, 40, 40, device=DEV).cpu()
        new = O.nca_step(x[mask], p, oc, fr, fu, None)
        x = x.clone()
        x[mask] = new
    tb = target.unsqueeze(0).expand(B, -1, -1, -1)
    per = masked_loss(x[:, :4], tb, cfg.loss_alpha_thr, cfg.loss_lam_area)
    loss = per.mean()
    close = per.detach() < cfg.stability_threshold
    stab = None
    if close.any():
        xs = x[close]
        for _ in range(cfg.stability_steps):
            fr = float(torch.empty(1, device=DEV).uniform_(cfg.fire_rate_min, cfg.fire_rate_max).item())
            fu = torch.rand(int(close.sum()), 1, 40, 40, device=DEV).cpu()
            xs = O.nca_step(xs, p, oc, fr, fu, None)
        stab = torch.nn.functional.mse_loss(xs[:, :4], tb[close])
        loss = loss + cfg.stability_weight * stab
    n_reset = int(cfg.reset_worst_prob * B)
    worst = torch.topk(per, n_reset).indices if n_reset > 0 else None
    do_reseed = random.random() < cfg.random_reseed_prob
    rand_idx = int(torch.randint(0, B, (1,), device=DEV).item()) if do_reseed else None
    loss.backward()
    names = [k for k, v in p.items() if v.requires_grad]
    plist = [p[k] for k in names]
    gnorm = torch.nn.utils.clip_grad_norm_(plist, cfg.clip_grad_norm)
    grads = {k: p[k].grad.detach().clone() for k in names}
    opt = torch.optim.Adam(plist, lr=cfg.learning_rate, weight_decay=cfg.weight_decay)
    opt.step()
    new_states = x.detach().clone()
    if worst is not None:
        new_states[worst] = trainer_seed(16, 40, len(worst), DEV).cpu()
    if do_reseed:
        new_states[rand_idx:rand_idx + 1] = trainer_seed(16, 40, 1, DEV).cpu()
    pool_cpu = pool_cpu.clone()
    pool_cpu[idx] = new_states
    return dict(per=per.detach(), close=close, stab=None if stab is None else stab.detach(), loss=loss.detach(),
                params={k: p[k].detach() for k in names}, grads=grads, gnorm=gnorm, worst=worst, pool=pool_cpu, steps=nca_steps)


def _setup(threshold):
    cfg = ClassicTrainConfig(batch_size=6, pool_size=12, nca_steps_min=10, nca_steps_max=16, long_prob=0.0,
                             stability_steps=5, stability_threshold=threshold, reset_worst_prob=0.34,
                             random_reseed_prob=1.0, learning_rate=1e-3)
    m = classic_model()
    tr = ClassicNCATrainer(m, _target(), cfg)
    # pool of grown states (different ages), identical for both runs
    torch.manual_seed(5); random.seed(5)
    with torch.no_grad():
        x = trainer_seed(16, 40, cfg.pool_size, DEV)
        s = make_schedule(m, cfg.pool_size, 40, 40, 40, fire_rate=0.7, seed=3, steps=list(range(16, 40, 2)))
        tr.pool.pool.copy_(rollout(m, x, s, impl="streaming"))
    return cfg, m, tr


@pytest.mark.parametrize("mode", ["all_close", "some_close", "none_close"])
def test_classic_iteration_with_stability_phase_matches_oracle(mode):
    cfg, m, tr = _setup(1e9)
    params0 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    pool0 = tr.pool.pool.detach().cpu().clone()
    if mode != "all_close":                      # pick the threshold from the (seed-determined) per-sample losses
        torch.manual_seed(11); random.seed(11)
        per = tr.train_step()["per_sample"].cpu()
        srt = per.sort().values                     # a threshold BETWEEN two losses (not on one)
        thr = float(0.5 * (srt[2] + srt[3])) if mode == "some_close" else float(srt[0]) * 0.5
        cfg, m, tr = _setup(thr)
    torch.manual_seed(11); random.seed(11)
    out = tr.train_step()
    torch.manual_seed(11); random.seed(11)
    ref = _oracle_iteration(params0, cfg, pool0, _target())
    assert np.array_equal(out["steps"], ref["steps"].numpy())
    assert rel_err(out["per_sample"].cpu(), ref["per"]) < 1e-5
    assert torch.equal(out["close"].cpu(), ref["close"])
    assert int(out["close"].sum()) == {"all_close": 6, "some_close": 3, "none_close": 0}[mode]
    if ref["stab"] is not None:
        assert abs(float(out["stab"]) - float(ref["stab"])) < 1e-5 * max(1.0, abs(float(ref["stab"])))
    assert abs(float(out["loss"]) - float(ref["loss"])) < 1e-5 * max(1.0, abs(float(ref["loss"])))
    assert torch.equal(out["worst"].cpu(), ref["worst"])
    named = dict(m.named_parameters())
    assert abs(float(out["grad_norm"]) - float(ref["gnorm"])) < 1e-4 * float(ref["gnorm"])
    for k, v in ref["grads"].items():              # clipped gradients of the joint backward through both rollouts
        assert rel_err(out["grads"][k].cpu(), v) < 1e-4, (k, rel_err(out["grads"][k].cpu(), v))
    for k, v in ref["params"].items():
        # Adam's first step moves a weight by lr * g / (|g| + 1e-8): weights whose gradient is ~1e-8 may move differently,
        # everything else agrees to fp32 rounding
        assert rel_err(named[k].detach().cpu(), v) < 2e-4, (k, rel_err(named[k].detach().cpu(), v))
    assert rel_err(tr.pool.pool.cpu(), ref["pool"]) < 1e-5
