"""numpy replica of the library's in-kernel fire stream (csrc/gnca_common.cuh: philox4x32_10 / philox_uniform;
the replicated-state kernel's fire_task in csrc/gnca_rep.cu draws the same blocks).  TEST INFRASTRUCTURE: it exists so
that the path bench.py times (fire="philox") can be pinned to the oracle -- uniforms from here, fed to the oracle as
recorded `fire_u`, must reproduce what the kernels do with (seed, offset) alone.

Stream layout (not torch's): uniform of (step t, sample b, cell) is word (idx & 3) of the Philox block with
counter = (idx >> 2) + offset, key = seed, idx = (t*B + b)*H*W + cell;  u = (word >> 8) * 2^-24  in [0, 1).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr_lo, ctr_hi, seed):
    """ctr_lo/ctr_hi: uint32 arrays (counter words 0,1; words 2,3 are zero); returns the 4 output words."""
    c0 = ctr_lo.astype(np.uint64); c1 = ctr_hi.astype(np.uint64)
    c2 = np.zeros_like(c0); c3 = np.zeros_like(c0)
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0; p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF; k1 = (k1 + W1) & 0xFFFFFFFF
    return c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32)


def fire_uniforms(seed, offset, T, B, H, W, t0=0):
    """[T,B,H,W] float32 uniforms of steps t0..t0+T-1, exactly the values the kernels compare with fire_rate."""
    HW = H * W
    idx = (np.arange(t0 * B * HW, (t0 + T) * B * HW, dtype=np.uint64))
    blk = (idx >> np.uint64(2)) + np.uint64(offset)
    ublk, inv = np.unique(blk, return_inverse=True)
    w = philox4x32_10((ublk & MASK).astype(np.uint32), (ublk >> np.uint64(32)).astype(np.uint32), int(seed))
    words = np.stack(w, axis=1)                       # [nblk, 4]
    v = words[inv, (idx & np.uint64(3)).astype(np.int64)]
    u = (v >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    return u.reshape(T, B, H, W)
