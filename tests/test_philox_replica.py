"""CPU checks of the numpy replica of the in-kernel Philox fire stream (tests/philox_replica.py): the Random123
known-answer vector, and the statistics the kernels rely on (fire fraction within 3 sigma of fire_rate, uniform
marginals).  The GPU tests (test_gpu_rollout.py::test_philox_fire_stream) prove the kernels draw exactly this stream."""
import numpy as np
import pytest

from philox_replica import fire_uniforms, philox4x32_10


def test_known_answer():
    # Random123 kat_vectors: philox4x32-10, counter 0, key 0
    w = philox4x32_10(np.zeros(1, np.uint32), np.zeros(1, np.uint32), 0)
    assert [int(x[0]) for x in w] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]


@pytest.mark.parametrize("seed,offset", [(11, 0), (2 ** 40 + 7, 12345), (4242, 0)])
def test_fire_fraction_and_uniformity(seed, offset):
    T, B, H, W = 16, 8, 40, 40
    u = fire_uniforms(seed, offset, T, B, H, W)
    assert u.dtype == np.float32 and u.min() >= 0.0 and u.max() < 1.0
    n = u.size
    for fr in (0.1, 0.5, 0.7, 0.9):
        frac = float((u <= np.float32(fr)).mean())
        sigma = (fr * (1 - fr) / n) ** 0.5
        assert abs(frac - fr) < 3 * sigma + 2.0 ** -24, (fr, frac, sigma)
    # per (step, sample) the fraction is binomial too (no structure across the block boundaries of 4 cells)
    per = (u <= 0.5).reshape(T * B, -1).mean(1)
    assert np.abs(per - 0.5).max() < 5 * (0.25 / (H * W)) ** 0.5
    hist, _ = np.histogram(u, bins=16, range=(0, 1))
    assert np.abs(hist / n - 1 / 16).max() < 5 * ((1 / 16) * (15 / 16) / n) ** 0.5


def test_stream_layout():
    """step t of a T-step draw == a 1-step draw at t0 = t; an offset of q blocks shifts the stream by 4 q cells."""
    a = fire_uniforms(7, 0, 3, 2, 8, 8)
    assert np.array_equal(a[2], fire_uniforms(7, 0, 1, 2, 8, 8, t0=2)[0])
    b = fire_uniforms(7, 5, 3, 2, 8, 8)
    assert np.array_equal(a.reshape(-1)[20:], b.reshape(-1)[:-20])
