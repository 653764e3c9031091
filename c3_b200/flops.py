"""Algorithmic flop accounting for the roofline figures.

F(d,K,m,s) = 8 d^3 (M_m + s + 1) + (32/3) d^3 + 4 K d^2 real flops per slice, with (m, s) the
MINIMAL Higham-2005 Pade order / squarings for the slice's 1-norm (not what any particular
implementation happens to evaluate), M_m = matmuls of the order-m evaluation, +1 the
ordered-product matmul, 32/3 d^3 the LU solve, 4 K d^2 the assembly.
"""
from __future__ import annotations

import numpy as np

THETA = (1.495585217958292e-2, 2.539398330063230e-1, 9.504178996162932e-1, 2.097847961257068)
THETA13 = 5.371920351148152
PADE_MATMULS = {3: 2, 5: 3, 7: 4, 9: 5, 13: 6}


def higham_order(norm1: float):
    for m, th in zip((3, 5, 7, 9), THETA):
        if norm1 < th:
            return m, 0
    s = max(int(np.ceil(np.log2(norm1 / THETA13))), 0)
    return 13, s


def flops_closed(d: int, K: int, m: int, s: int) -> float:
    return 8.0 * d ** 3 * (PADE_MATMULS[m] + s + 1) + (32.0 / 3.0) * d ** 3 + 4.0 * K * d * d


def flops_lindblad(d: int, m: int, s: int) -> float:
    D = d * d
    return 8.0 * D ** 3 * (PADE_MATMULS[m] + s + 1) + (32.0 / 3.0) * D ** 3 + 8.0 * d * D


def flops_per_slice_closed(h0, hks, signals_sample, dt: float) -> float:
    """Mean algorithmic flops per slice over a sample signals[b,K,N] of the workload."""
    h0 = np.asarray(h0)
    hks = np.asarray(hks)
    sig = np.asarray(signals_sample)
    d, K = h0.shape[-1], hks.shape[0]
    H = h0[None, None] + np.einsum("bkn,kij->bnij", sig, hks)
    n1 = np.abs(H * dt).sum(axis=-2).max(axis=-1).ravel()
    return float(np.mean([flops_closed(d, K, *higham_order(x)) for x in n1]))
