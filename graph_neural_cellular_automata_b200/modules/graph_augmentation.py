"""Drop-in for the reference's `modules/graph_augmentation.py` (GraphAugmentation, :8-169).

Keeps the constructor, the attributes the reference's scripts read (`offsets`, `num_neighbors`,
`zero_padded_shift`, `alive_to_alive`, `scaling`, `query_proj`, `key_proj`, `msg_proj`, `gate_mlp`), the
state-dict keys, and the RNG side effect (exactly one `random.sample(self.offsets, k)` per forward).
The arithmetic runs in libgnca.so: the k shifted copies of K / M / A_send the reference materialises are
replaced by in-place neighbour reads (see csrc/gnca_common.cuh: gather_senders).
"""
from __future__ import annotations

import math
import random
from typing import List, Tuple

import torch
import torch.nn as nn


def build_offsets(radius: int) -> List[Tuple[int, int]]:
    """Mid-range ring: Chebyshev distance in (1, radius], dy-major / dx-minor (graph_augmentation.py:73-83)."""
    span = range(-radius, radius + 1)
    return [(dy, dx) for dy in span for dx in span if max(abs(dy), abs(dx)) >= 2]


def _words_sample(table, n: int, k: int, T: int):
    """T x random.sample(offsets, k) as an int8 [T,k,2] array: a block of raw MT19937 outputs is pulled with ONE
    random.getrandbits call, the C host function gnca_host_sample_offsets_words replays random.sample on it and says
    how many outputs it consumed; the saved state is restored and exactly that many outputs are skipped.  Same
    stream as T reference forwards, no state <-> numpy conversion (that cost more than the sampling itself)."""
    import ctypes as C
    import numpy as np
    from .. import _lib
    lib = _lib.load()
    state = random.getstate()
    out = np.empty((T, k, 2), dtype=np.int8)
    used = C.c_int32(0)
    m = int(2.3 * T * k) + 64                      # ~2 outputs per draw at worst (rejection), + margin
    while True:
        words = random.getrandbits(32 * m).to_bytes(4 * m, "little")
        rc = lib.gnca_host_sample_offsets_words(words, m, table.ctypes.data, n, k, T, out.ctypes.data, C.byref(used))
        random.setstate(state)
        if rc == 0:
            break
        if rc != _lib.GNCA_ERR_UNSUPPORTED or m > (1 << 24):
            raise RuntimeError("gnca_host_sample_offsets_words failed")
        m *= 4                                    # block too short (rejection sampling ran long): retry with more
    if used.value:
        random.getrandbits(32 * used.value)
    return out


_WORDS_SAMPLE_CACHE = {}


def _words_sample_ok(offsets, n: int, k: int) -> bool:
    """One-time self-check per (n, k): the block replay must reproduce random.sample AND leave the same state."""
    import numpy as np
    key = (n, k)
    if key not in _WORDS_SAMPLE_CACHE:
        ok = False
        state = random.getstate()
        try:
            if state[0] == 3 and len(state[1]) == 625:
                ref = [[list(o) for o in random.sample(offsets, k)] for _ in range(5)]
                after_ref = random.getstate()
                random.setstate(state)
                table = np.ascontiguousarray(np.asarray(offsets, dtype=np.int8).reshape(n, 2))
                mine = _words_sample(table, n, k, 5).tolist()
                ok = mine == ref and random.getstate() == after_ref
        except Exception:
            ok = False
        finally:
            random.setstate(state)
        _WORDS_SAMPLE_CACHE[key] = ok
    return _WORDS_SAMPLE_CACHE[key]


class GraphAugmentation(nn.Module):
    def __init__(self, n_channels: int, d_model: int = 16, attention_radius: int = 4, num_neighbors: int = 8,
                 gating_hidden: int = 32, *, alive_to_alive: bool = True, zero_padded_shift: bool = True,
                 alpha_thr: float = 0.1):
        super().__init__()
        self.n_channels = n_channels
        self.d_model = d_model
        self.attention_radius = attention_radius
        self.num_neighbors = num_neighbors
        self.alive_to_alive = bool(alive_to_alive)
        self.zero_padded_shift = bool(zero_padded_shift)
        self.alpha_thr = float(alpha_thr)
        # parameter containers, created in the reference's order (:55-68) so seeded init is identical
        self.query_proj = nn.Conv2d(n_channels, d_model, 1)
        self.key_proj = nn.Conv2d(n_channels, d_model, 1)
        self.msg_proj = nn.Conv2d(n_channels, n_channels, 1)
        self.scaling = nn.Parameter(torch.tensor(math.sqrt(d_model), dtype=torch.float32))
        # never evaluated by the reference's forward ("implemented but disabled"); kept for checkpoint keys
        self.gate_mlp = nn.Sequential(nn.Conv2d(2 * n_channels, gating_hidden, 1), nn.ReLU(inplace=False),
                                      nn.Conv2d(gating_hidden, n_channels, 1), nn.Sigmoid())
        self.offsets = build_offsets(attention_radius)

    _build_offsets = staticmethod(build_offsets)

    def draw_offsets(self) -> List[Tuple[int, int]]:
        """The per-forward draw of graph_augmentation.py:120-121 (Python global RNG, also at message_gain 0)."""
        k = min(self.num_neighbors, len(self.offsets))
        return random.sample(self.offsets, k) if k > 0 else []

    def draw_offsets_array(self, T: int):
        """T consecutive `draw_offsets()` results as one int8 array [T,k,2] -- the SAME python-RNG stream as T forward
        calls.  Fast path: one block of raw MT19937 outputs replayed by the C host function (`_words_sample`, guarded by
        a one-time self-check against `random.sample` including the state it leaves behind); otherwise plain
        `random.sample` per step."""
        import numpy as np
        n, k = len(self.offsets), min(self.num_neighbors, len(self.offsets))
        table = getattr(self, "_offset_table", None)
        if table is None or table.shape[0] != n:
            table = self._offset_table = np.ascontiguousarray(np.asarray(self.offsets, dtype=np.int8).reshape(n, 2))
        if k == 0 or T == 0:
            return np.zeros((T, 0, 2), np.int8)
        if T >= 4 and _words_sample_ok(self.offsets, n, k):
            return _words_sample(table, n, k, T)
        flat = [v for _ in range(T) for o in random.sample(self.offsets, k) for v in o]
        return np.asarray(flat, dtype=np.int8).reshape(T, k, 2)

    def forward(self, x: torch.Tensor, return_attention_map: bool = False):
        from .. import graph_ops
        chosen = self.draw_offsets()
        return graph_ops.graph_message(self, x, chosen, return_attention_map)
