"""Philox4x32-10 + Box–Muller reference in numpy (TEST INFRASTRUCTURE ONLY).

Restates the published counter-based generator of Salmon et al., "Parallel Random Numbers: As
Easy as 1, 2, 3" (SC'11) as shipped in Random123 (philox.h) and cuRAND
(curand_philox4x32_x.h).  Pinned by Random123's known-answer vectors in
tests/test_oracle_pins.py.  The CUDA SDE kernel's Brownian increments follow the contract in
`normals()` (SURVEY Appendix B "Philox contract"), so the kernel's stream can be regenerated on
the CPU bit-for-bit up to the device's log/sin/cos rounding.
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = np.uint32(0x9E3779B9)
PHILOX_W1 = np.uint32(0xBB67AE85)


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """counter (..., 4) uint32, key (..., 2) uint32 -> (..., 4) uint32, 10 rounds."""
    c = np.array(counter, dtype=np.uint32, copy=True)
    k = np.array(np.broadcast_to(key, c.shape[:-1] + (2,)), dtype=np.uint32, copy=True)
    for _ in range(10):
        p0 = c[..., 0].astype(np.uint64) * PHILOX_M0
        p1 = c[..., 2].astype(np.uint64) * PHILOX_M1
        hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
        lo0 = (p0 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
        lo1 = (p1 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        n0 = hi1 ^ c[..., 1] ^ k[..., 0]
        n1 = lo1
        n2 = hi0 ^ c[..., 3] ^ k[..., 1]
        n3 = lo0
        c = np.stack([n0, n1, n2, n3], axis=-1)
        with np.errstate(over="ignore"):
            k[..., 0] = k[..., 0] + PHILOX_W0
            k[..., 1] = k[..., 1] + PHILOX_W1
    return c


def u01(x: np.ndarray) -> np.ndarray:
    """uint32 -> float32 in (0, 1]: (x + 0.5) * 2^-32 rounded to fp32 would reach 1.0 exactly for the top
    values, which is fine for log(); this is cuRAND's _curand_uniform convention
    (x * 2^-32 + 2^-33)."""
    return (x.astype(np.float32) * np.float32(2.3283064e-10) + np.float32(2.3283064e-10 / 2)).astype(np.float32)


def box_muller(u1: np.ndarray, u2: np.ndarray):
    """Two N(0,1) from two uniforms in (0,1]: r = sqrt(-2 ln u1), (r sin 2πu2, r cos 2πu2)."""
    r = np.sqrt(np.float32(-2.0) * np.log(u1.astype(np.float32)))
    th = np.float32(2.0 * np.pi) * u2.astype(np.float32)
    return (r * np.sin(th)).astype(np.float32), (r * np.cos(th)).astype(np.float32)


def normals(seed: int, traj: np.ndarray, step: int, d_block: int, stream: int = 0) -> np.ndarray:
    """The SDE kernel's contract: key = (seed lo, seed hi); counter = (traj_idx, step_idx, d_block, stream)
    -> 4 uint32 -> 2x Box–Muller -> 4 standard normals ordered (sin01, cos01, sin23, cos23) for state
    components 4*d_block .. 4*d_block+3.  traj: (N,) global trajectory indices.  Returns (N, 4) fp32."""
    traj = np.asarray(traj, dtype=np.uint32)
    ctr = np.stack([traj,
                    np.full_like(traj, step, dtype=np.uint32),
                    np.full_like(traj, d_block, dtype=np.uint32),
                    np.full_like(traj, stream, dtype=np.uint32)], axis=-1)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    r = philox4x32_10(ctr, key)
    a, b = box_muller(u01(r[..., 0]), u01(r[..., 1]))
    c, d = box_muller(u01(r[..., 2]), u01(r[..., 3]))
    return np.stack([a, b, c, d], axis=-1)
