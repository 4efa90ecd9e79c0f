"""Counter-based synthetic gradient histories (numpy twin of ``dxm_synth_gradients``).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The recipe uses integer hashing
(splitmix64) and exactly rounded floating-point operations only, so the device generator and this
file produce identical bits for any (seed, global point index, component, increment).

Recipes:
* ``strain``: ``eps_k = ((k / K) * (amp * u6)) * (2 u_c - 1)``, c = 0..5 Mandel components -- proportional
  loading along a random direction with a random amplitude (cfg2 / cfg4).
* ``defgrad``: ``F_k = I + ((k / K) * (amp * u9)) * (2 u_c - 1)``, c = 0..8 in the reference's
  non-symmetric ordering ``[11,22,33,12,21,13,31,23,32]`` (``dolfinx_materials/utils.py:173-186``) (cfg3).
"""

import numpy as np

_GOLD = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_TWO_M53 = 2.0 ** -53


def uniform(seed, idx, comp):
    """u in [0,1): splitmix64 of ``seed + (16*idx + comp + 1) * GOLD``; idx is a uint64 array."""
    with np.errstate(over="ignore"):
        idx = np.asarray(idx, dtype=np.uint64)
        z = np.uint64(seed) + (idx * np.uint64(16) + np.uint64(comp + 1)) * _GOLD
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * _TWO_M53


def strain(n, seed, amp, k, K, start=0):
    idx = np.arange(start, start + n, dtype=np.uint64)
    a = amp * uniform(seed, idx, 6)
    scale = (float(k) / float(K)) * a
    out = np.empty((n, 6))
    for c in range(6):
        out[:, c] = 0.0 + scale * (2.0 * uniform(seed, idx, c) - 1.0)
    return out


def defgrad(n, seed, amp, k, K, start=0):
    idx = np.arange(start, start + n, dtype=np.uint64)
    a = amp * uniform(seed, idx, 9)
    scale = (float(k) / float(K)) * a
    out = np.empty((n, 9))
    for c in range(9):
        ident = 1.0 if c < 3 else 0.0
        out[:, c] = ident + scale * (2.0 * uniform(seed, idx, c) - 1.0)
    return out
