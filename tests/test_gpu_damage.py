"""PRODUCT-path damage parity (graph_neural_cellular_automata_b200/utils/damage.py vs the reference's
utils/damage.py:16-138): the draws the reference made (`torch.randint`, `torch.rand_like`, `random.random`, recorded in
tests/golden/damage.npz by make_golden.py) are fed, in order, through the product's operators by patching the same three
functions; the damaged states must equal the reference's outputs (bit-exact for the {0,1} masks, 1e-6 for the gaussian).
Also: `sample_damage_mask` / `apply_damage_policy_` consume the policy draws in the reference's order (gate ->
kind -> size -> geometry) and `sample_damage_mask` is pure (it never mutates `state`)."""
import random

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from graph_neural_cellular_automata_b200.utils import damage as D

DEV = "cuda"
T32 = lambda a: torch.from_numpy(np.asarray(a)).float()


class Replay:
    """Patch torch.randint / torch.rand_like / random.random to replay recorded draws (and count them)."""

    def __init__(self, ints=(), rand=None, pyrandom=()):
        self.ints, self.rand, self.py = list(int(v) for v in ints), rand, list(float(v) for v in pyrandom)
        self.calls = []

    def __enter__(self):
        self._ri, self._rl, self._rr = torch.randint, torch.rand_like, random.random

        def ri(lo, hi, size, **k):
            v = self.ints.pop(0)
            assert lo <= v < hi, (lo, v, hi)              # the product asks for the same range the reference drew from
            self.calls.append("randint")
            return torch.full(tuple(size), v, dtype=torch.int64, device=k.get("device", "cpu"))

        def rl(t, **k):
            self.calls.append("rand_like")
            assert self.rand is not None and tuple(t.shape) == tuple(self.rand.shape)
            return self.rand.to(t.device)

        def rr():
            self.calls.append("random")
            return self.py.pop(0)
        torch.randint, torch.rand_like, random.random = ri, rl, rr
        return self

    def __exit__(self, *exc):
        torch.randint, torch.rand_like, random.random = self._ri, self._rl, self._rr
        if exc[0] is None:
            assert not self.ints and not self.py, "the product consumed fewer draws than the reference"


@pytest.mark.parametrize("kind", ["square", "circle", "stripes", "alpha_drop", "saltpepper", "gaussian"])
def test_operators_replay_reference_draws(kind):
    d = load_golden("damage.npz")
    s = T32(d["state"]).to(DEV)
    ops = {"square": lambda x: D.cutout_square_(x, 9), "circle": lambda x: D.cutout_circle_(x, 5),
           "stripes": lambda x: D.stripe_wipe_(x, 6, orientation="auto"),
           "alpha_drop": lambda x: D.alpha_dropout_(x, 0.15, alpha_thr=0.2, hard=True),
           "saltpepper": lambda x: D.salt_pepper_alpha_(x, 0.02),
           "gaussian": lambda x: D.gaussian_hole_(x, radius=6, softness=0.35)}
    rand = T32(d[f"{kind}:rand"]) if f"{kind}:rand" in d else None
    with Replay(d[f"{kind}:ints"], rand, d[f"{kind}:pyrandom"]):
        x = s.clone()
        ops[kind](x)
    ref = T32(d[f"{kind}:out"])
    if kind == "gaussian":
        assert rel_err(x.cpu(), ref) < 1e-6
    else:
        assert torch.equal(x.cpu(), ref)
    assert not torch.equal(x.cpu(), s.cpu())              # the damage did something


def test_policy_draw_order_and_purity():
    """apply_damage_policy_ (damage.py:101-138): torch.rand(1) gate, random.choices(kind), random.randint(size), then the
    kind's geometry draws -- checked by running the reference-order draws by hand next to the product."""
    d = load_golden("damage.npz")
    s = T32(d["state"]).to(DEV)
    cfg = {"start_epoch": 100, "prob": 1.0, "kinds": {"square": 0.3, "circle": 0.3, "stripes": 0.2, "gaussian": 0.2},
           "size_min": 6, "size_max": 12, "gaussian_softness": 0.35}
    assert D.sample_damage_mask(s, cfg, epoch=99) is None                       # before start_epoch: no draw at all
    for seed in range(6):
        torch.manual_seed(seed); random.seed(seed)
        x = s.clone()
        keep = x.clone()
        M = D.sample_damage_mask(x, cfg, epoch=150)
        assert torch.equal(x, keep), "sample_damage_mask must not mutate the state"
        # replay by hand in the reference's order
        torch.manual_seed(seed); random.seed(seed)
        gate = torch.rand(1, device=DEV).item()
        assert gate <= 1.0
        names, weights = zip(*cfg["kinds"].items())
        kind = random.choices(names, weights=weights, k=1)[0]
        size = random.randint(6, 12)
        y = s.clone()
        if kind == "square":
            D.cutout_square_(y, size)
        elif kind == "circle":
            D.cutout_circle_(y, size // 2)
        elif kind == "stripes":
            D.stripe_wipe_(y, size, "auto")
        else:
            D.gaussian_hole_(y, max(1, size // 2), 0.35)
        torch.manual_seed(seed); random.seed(seed)
        z = s.clone()
        D.apply_damage_policy_(z, cfg, epoch=150)
        assert torch.equal(z, y), kind
        assert torch.equal(z, s * D.dense_mask(M, s)), kind
