"""Generate the golden fixtures in this directory by running the REFERENCE's own modules.

Run in the build container only (the reference checkout does not travel to the GPU box):

    python tests/golden/make_golden.py            # needs /root/reference

It imports `/root/reference/src/modules/*`, `utils/damage.py`, `utils/nca_init.py` unmodified, loads two of
the shipped checkpoints (realistic trained weights), records every random draw the modules make
(`torch.rand` fire uniforms, `random.sample` offsets, damage geometry) and stores inputs + outputs as
`.npz`.  `tests/test_oracle_golden.py` pins `oracle/nca_oracle.py` against these files; the `-m gpu` tests
pin the CUDA path against the same files.  Nothing at test time reads /root/reference.
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np
import torch

REF = os.environ.get("GNCA_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REF, "src"))
HERE = os.path.dirname(os.path.abspath(__file__))

from modules.nca import NeuralCA                      # noqa: E402
from modules.ncagraph import NeuralCAGraph            # noqa: E402
from modules.perception import FixedSobelPerception   # noqa: E402
from modules.graph_augmentation import GraphAugmentation  # noqa: E402
from utils import damage as ref_damage                # noqa: E402
from utils.nca_init import make_seed                  # noqa: E402

GRAPH_CKPT = os.path.join(REF, "outputs/graphaug_nca/train_inter_loss/gecko/checkpoints/nca_epoch960.pt")
CLASSIC_CKPT = os.path.join(REF, "outputs/classic_nca/train_inter_loss/gecko/checkpoints/nca_epoch990.pt")


class Recorder:
    """Record torch.rand / random.sample draws made inside the reference modules."""

    def __init__(self):
        self.fire, self.chosen = [], []
        self._rand, self._sample = torch.rand, random.sample

    def __enter__(self):
        def rand(*a, **k):
            r = self._rand(*a, **k)
            self.fire.append(r.clone())
            return r

        def sample(pop, k):
            s = self._sample(pop, k)
            self.chosen.append(list(s))
            return s

        torch.rand, random.sample = rand, sample
        return self

    def __exit__(self, *exc):
        torch.rand, random.sample = self._rand, self._sample


def np_state(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}


def build_graph(torus=True, message_gain=0.25):
    ck = torch.load(GRAPH_CKPT, map_location="cpu", weights_only=False)
    m = NeuralCAGraph(16, update_hidden=128, img_size=40, update_gain=0.05, alpha_thr=0.12,
                      use_groupnorm=True, message_gain=message_gain, hidden_only=True,
                      graph_d_model=16, graph_attention_radius=4, graph_num_neighbors=8,
                      graph_gating_hidden=32, graph_zero_padded_shift=not torus)
    missing, unexpected = m.load_state_dict(ck["model_state"], strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return m


def build_classic():
    ck = torch.load(CLASSIC_CKPT, map_location="cpu", weights_only=False)
    m = NeuralCA(16, update_hidden=128, img_size=40, update_gain=0.1, alpha_thr=0.1, use_groupnorm=True)
    missing, unexpected = m.load_state_dict(ck["model_state"], strict=False)
    assert not missing and not unexpected
    return m


def loss_premult_rgba(pred, target):   # train_graph_augmented_nca.py:52-61 (restated; trainer not importable)
    rgba = torch.cat([pred[:, :3] * pred[:, 3:4], pred[:, 3:4]], dim=1)
    return torch.nn.functional.mse_loss(rgba, target, reduction="none").mean(dim=(1, 2, 3))


def rollout_record(model, x0, T, fire_rate, snaps, graph=True, gains=None, base_gain=None):
    """Run T reference forward calls, recording draws and snapshots of x at the steps in `snaps`."""
    out = {}
    x = x0.clone()
    with Recorder() as rec, torch.no_grad():
        for t in range(T):
            if t in snaps:
                out[f"x_{t}"] = x.numpy().copy()
            if gains is not None:
                model.message_gain = float(gains[t])
            x = model(x, fire_rate=fire_rate)
        out[f"x_{T}"] = x.numpy().copy()
    if base_gain is not None:
        model.message_gain = base_gain
    out["fire_u"] = torch.stack(rec.fire, 0).numpy() if rec.fire else np.zeros((0,), np.float32)
    if graph:
        out["chosen"] = np.asarray(rec.chosen, dtype=np.int32)      # [T,k,2]
    return out, x


def main():
    torch.set_num_threads(4)
    torch.manual_seed(42)
    random.seed(42)
    np.random.seed(42)

    # ---- weights ---------------------------------------------------------------------------------
    g_t = build_graph(torus=True)
    g_z = build_graph(torus=False)
    cl = build_classic()
    np.savez_compressed(os.path.join(HERE, "weights_graph_ep960.npz"), **np_state(g_t.state_dict()))
    np.savez_compressed(os.path.join(HERE, "weights_classic_ep990.npz"), **np_state(cl.state_dict()))

    # ---- known-answer facts -----------------------------------------------------------------------
    facts = {
        "offsets_r4": np.asarray(GraphAugmentation._build_offsets(4), dtype=np.int32),
        "offsets_r2": np.asarray(GraphAugmentation._build_offsets(2), dtype=np.int32),
        "n_params_graph_total": np.int64(sum(p.numel() for p in g_t.parameters())),
        "n_params_graph_trainable": np.int64(sum(p.numel() for p in g_t.parameters() if p.requires_grad)),
        "n_params_classic_total": np.int64(sum(p.numel() for p in cl.parameters())),
        "graph_keys": np.asarray(sorted(g_t.state_dict().keys())),
        "classic_keys": np.asarray(sorted(cl.state_dict().keys())),
    }
    xs = torch.randn(2, 16, 9, 11)
    facts["perc_in"] = xs.numpy()
    facts["perc_out"] = FixedSobelPerception(16)(xs).detach().numpy()
    facts["alive_in"] = torch.rand(3, 16, 12, 12).numpy()
    facts["alive_out"] = g_t._alive_mask(torch.from_numpy(facts["alive_in"])).numpy()
    np.savez_compressed(os.path.join(HERE, "facts.npz"), **facts)

    # ---- surrogate target: 200-step growth of the trained graph model ------------------
    torch.manual_seed(7); random.seed(7)
    x = make_seed(16, 40, 1)
    with torch.no_grad():
        for _ in range(200):
            x = g_t(x, fire_rate=0.5)
    tgt = x[0, :4].clamp(0, 1)
    tgt[:3] = tgt[:3] * tgt[3:4]
    np.save(os.path.join(HERE, "target_gecko_surrogate.npy"), tgt.numpy())

    # ---- case 1: graph torus growth rollout, B=2, 48 steps, fire 0.5 --------------------------------
    torch.manual_seed(1); random.seed(1)
    rec, x48 = rollout_record(g_t, make_seed(16, 40, 2), 48, 0.5, snaps={0, 1, 8, 16, 32, 47})
    np.savez_compressed(os.path.join(HERE, "graph_torus_rollout.npz"), fire_rate=np.float32(0.5),
                        message_gain=np.float32(0.25), **rec)

    # ---- case 2: single steps from an aged state with attention map (torus and zero-pad) ------------
    for name, model in (("torus", g_t), ("zeropad", g_z)):
        torch.manual_seed(2); random.seed(2)
        with Recorder() as r, torch.no_grad():
            y, attn = model(x48, fire_rate=0.7, return_attention=True)
            m_only, attn2 = model.graph(x48, return_attention_map=True)
            y_full = model(x48, fire_rate=1.0)                    # no fire mask, no torch.rand draw
        np.savez_compressed(os.path.join(HERE, f"graph_{name}_step.npz"), x_in=x48.numpy(), x_out=y.numpy(),
                            attn=attn.numpy(), fire_u=r.fire[0].numpy(), fire_rate=np.float32(0.7),
                            chosen=np.asarray(r.chosen[0], np.int32), chosen_graph=np.asarray(r.chosen[1], np.int32),
                            graph_m=m_only.numpy(), graph_attn=attn2.numpy(),
                            chosen_full=np.asarray(r.chosen[2], np.int32), x_out_full=y_full.numpy(),
                            message_gain=np.float32(0.25))

    # ---- case 3: zero-pad rollout (module default), 12 steps from the aged state ---------------------
    torch.manual_seed(3); random.seed(3)
    rec, _ = rollout_record(g_z, x48, 12, 0.5, snaps={0})
    np.savez_compressed(os.path.join(HERE, "graph_zeropad_rollout.npz"), fire_rate=np.float32(0.5),
                        message_gain=np.float32(0.25), **rec)

    # ---- case 4: classic growth rollout B=2, 48 steps ------------------------------------------------
    torch.manual_seed(4); random.seed(4)
    rec, xc48 = rollout_record(cl, make_seed(16, 40, 2), 48, 0.5, snaps={0, 1, 16, 32}, graph=False)
    np.savez_compressed(os.path.join(HERE, "classic_rollout.npz"), fire_rate=np.float32(0.5), **rec)

    # ---- case 5: gradients through a 16-step rollout (graph torus, trainer-style gating msg_every=3) -
    target = torch.from_numpy(np.load(os.path.join(HERE, "target_gecko_surrogate.npy")))

    def grad_case(model, x0, T, fname, gains, fire_rates, steps=None):
        model.zero_grad(set_to_none=True)
        x0 = x0.clone().requires_grad_(True)
        state = x0
        with Recorder() as r:
            for t in range(T):
                model.message_gain = float(gains[t])
                if steps is None:
                    state = model(state, fire_rate=float(fire_rates[t]))
                else:
                    mask = steps > t
                    if not mask.any():
                        continue
                    new = model(state[mask], fire_rate=float(fire_rates[t]))
                    state = state.clone()
                    state[mask] = new
        per = loss_premult_rgba(state[:, :4], target.unsqueeze(0).expand(state.shape[0], -1, -1, -1))
        loss = per.mean()
        loss.backward()
        model.message_gain = 0.25
        out = {"x0": x0.detach().numpy(), "x_T": state.detach().numpy(), "per_sample": per.detach().numpy(),
               "loss": np.float32(loss.item()), "grad_x0": x0.grad.numpy(),
               "gains": np.asarray(gains, np.float32), "fire_rates": np.asarray(fire_rates, np.float32),
               "chosen": np.asarray(r.chosen, np.int32)}
        # fire draws have one row per ACTIVE sample: pad to B rows in sample order for storage
        B = x0.shape[0]
        fu = np.zeros((T, B, 1, 40, 40), np.float32)
        ti = 0
        for t in range(T):
            if steps is None:
                fu[t] = r.fire[ti].numpy(); ti += 1
            else:
                mask = (steps > t).numpy()
                if mask.any():
                    fu[t, mask] = r.fire[ti].numpy(); ti += 1
        out["fire_u"] = fu
        if steps is not None:
            out["steps"] = steps.numpy().astype(np.int32)
        for k, prm in model.named_parameters():
            out["grad:" + k] = (prm.grad.numpy() if prm.grad is not None else np.zeros((0,), np.float32))
        np.savez_compressed(os.path.join(HERE, fname), **out)

    torch.manual_seed(5); random.seed(5)
    T = 16
    gains = [0.25 if t % 3 == 0 else 0.0 for t in range(T)]
    frs = [float(torch.empty(1).uniform_(0.5, 0.9)) for _ in range(T)]
    grad_case(g_t, x48, T, "graph_torus_grads.npz", gains, frs)

    torch.manual_seed(6); random.seed(6)
    T = 6
    grad_case(g_z, x48, T, "graph_zeropad_grads.npz", [0.25] * T, [0.6] * T)

    # per-sample step counts (trainer semantics), B=3 built from the aged states
    torch.manual_seed(8); random.seed(8)
    x3 = torch.cat([x48, make_seed(16, 40, 1)], 0)
    T = 10
    steps = torch.tensor([10, 4, 7])
    gains = [0.25 if t % 3 == 0 else 0.0 for t in range(T)]
    frs = [float(torch.empty(1).uniform_(0.5, 0.9)) for _ in range(T)]
    grad_case(g_t, x3, T, "graph_torus_grads_ragged.npz", gains, frs, steps=steps)

    # classic gradients, 12 steps
    torch.manual_seed(9); random.seed(9)

    class _C(torch.nn.Module):          # give the classic model a message_gain attribute to reuse grad_case
        pass
    cl.message_gain = 0.0
    grad_case(cl, xc48, 12, "classic_grads.npz", [0.0] * 12, [0.5] * 12)

    # ---- case 6: damage operators (record the geometry draws) -----------------------------------------
    dmg = {}
    base = x48.repeat(2, 1, 1, 1)[:3].clone()                 # B=3 aged states
    dmg["state"] = base.numpy()
    cfg = {"alpha_thr": 0.2, "alpha_dropout_p": 0.15, "salt_pepper_p": 0.02, "gaussian_softness": 0.35}
    real_randint, real_rand_like, real_random = torch.randint, torch.rand_like, random.random

    def run(kind, fn):
        ints, rls, rnd = [], [], []

        def ri(*a, **k):
            v = real_randint(*a, **k); ints.append(int(v)); return v

        def rl(*a, **k):
            v = real_rand_like(*a, **k); rls.append(v.clone()); return v

        def rr():
            v = real_random(); rnd.append(v); return v
        torch.randint, torch.rand_like, random.random = ri, rl, rr
        try:
            s = base.clone()
            fn(s)
        finally:
            torch.randint, torch.rand_like, random.random = real_randint, real_rand_like, real_random
        dmg[f"{kind}:out"] = s.numpy()
        dmg[f"{kind}:ints"] = np.asarray(ints, np.int32)
        dmg[f"{kind}:pyrandom"] = np.asarray(rnd, np.float64)
        if rls:
            dmg[f"{kind}:rand"] = rls[0].numpy()

    torch.manual_seed(10); random.seed(10)
    run("square", lambda s: ref_damage.cutout_square_(s, 9))
    run("circle", lambda s: ref_damage.cutout_circle_(s, 5))
    run("stripes", lambda s: ref_damage.stripe_wipe_(s, 6, orientation="auto"))
    run("alpha_drop", lambda s: ref_damage.alpha_dropout_(s, 0.15, alpha_thr=0.2, hard=True))
    run("saltpepper", lambda s: ref_damage.salt_pepper_alpha_(s, 0.02))
    run("gaussian", lambda s: ref_damage.gaussian_hole_(s, radius=6, softness=0.35))
    np.savez_compressed(os.path.join(HERE, "damage.npz"), **dmg)

    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        print(f"  {f:40s} {os.path.getsize(os.path.join(HERE, f)) / 1024:8.1f} KiB")


if __name__ == "__main__":
    main()
