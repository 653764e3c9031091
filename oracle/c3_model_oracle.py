"""Numpy restatement of the few pieces of the reference's ``Model`` needed to REBUILD inputs
that the golden pickles do not carry (the dressed collapse operators of the two-qubit chip).

TEST INFRASTRUCTURE ONLY (see oracle/c3_oracle.py header).  Citations are relative to the
reference checkout (q-optimize/c3 @ 48b7917e).
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np


def hilbert_space_kron(op: np.ndarray, indx: int, dims: Sequence[int]) -> np.ndarray:
    """Identity on every subsystem except ``indx`` (c3/utils/qt_utils.py:68-96)."""
    out = np.eye(1)
    for j, dj in enumerate(dims):
        out = np.kron(out, op if j == indx else np.identity(dj))
    return out


def annihilators(dims: Sequence[int]):
    """a_j = sum_n sqrt(n) |n-1><n| on subsystem j (c3/model.py:161-171)."""
    return [hilbert_space_kron(np.diag(np.sqrt(np.arange(1, dj)), k=1), j, dims)
            for j, dj in enumerate(dims)]


def resonator(a):            # c3/libraries/hamiltonians.py:16-33
    return a.T.conj() @ a


def duffing(a):              # c3/libraries/hamiltonians.py:36-53
    n = a.T.conj() @ a
    return 0.5 * (n - np.eye(n.shape[0])) @ n


def int_XX(a, b):            # c3/libraries/hamiltonians.py:79-99
    return (a.T.conj() + a) @ (b.T.conj() + b)


def x_drive(a):              # c3/libraries/hamiltonians.py (x_drive): a^dag + a
    return a.T.conj() + a


def reorder_frame(e: np.ndarray, v: np.ndarray, ordered: bool = True):
    """c3/model.py:453-492: assign every eigenvector to the bare state it overlaps most with.  If every
    eigenvector has one component of probability > 0.5 that component decides; otherwise ("overly dressed") the
    assignment is greedy: repeatedly take the largest remaining |v|^2 and strike out its row and column."""
    if not ordered:
        return np.eye(len(e)), np.real(e), v
    v_sq = np.real(v * np.conj(v))
    if v_sq.max(axis=0).min() > 0.5:
        reorder = (v_sq > 0.5).astype(np.float64)
    else:
        vc = v_sq.copy()
        reorder = np.zeros_like(vc)
        for _ in range(vc.shape[1]):
            idx = np.unravel_index(np.argmax(vc), vc.shape)
            vc[idx[0], :] = 0
            vc[:, idx[1]] = 0
            reorder[idx] = 1
    signed = np.sign(np.real(v)) * reorder
    eigenframe = reorder @ np.real(e)
    transform = v @ signed.T
    return reorder, eigenframe, transform


def dressing_transform(drift: np.ndarray, ordered: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """eigh + reorder by largest overlap + sign fix (c3/model.py:453-502)."""
    e, v = np.linalg.eigh(drift)
    _, eigenframe, transform = reorder_frame(e, v, ordered)
    return eigenframe, transform.astype(np.complex128)


def dress(transform: np.ndarray, op: np.ndarray) -> np.ndarray:
    """T^dag op T (c3/model.py:504-534)."""
    return transform.conj().T @ op @ transform


def qubit_collapse_op(a: np.ndarray, t1: float = None, t2star: float = None) -> np.ndarray:
    """One summed collapse operator per qubit, no temperature term
    (c3/libraries/chip.py:191-242; note ``sum(Ls)`` at :242)."""
    L = np.zeros_like(a, dtype=np.complex128)
    if t1 is not None:
        L = L + (1.0 / t1) ** 0.5 * a
    if t2star is not None:
        L = L + (0.5 / t2star) ** 0.5 * 2.0 * (a.T.conj() @ a)
    return L


def z_drive(a):              # c3/libraries/hamiltonians.py:166-182
    return a.T.conj() @ a


def transmon_factor(phi: float, phi_0: float, d: float) -> float:
    """Flux dependence of a tunable transmon (c3/libraries/chip.py:355-372)."""
    x = np.pi * phi / phi_0
    return float(np.sqrt(np.sqrt(np.cos(x) ** 2 + d ** 2 * np.sin(x) ** 2)))


def tunable_coupler_drift(phi: float = 2.3, g1: float = 142e6, g2: float = 116e6):
    """Bare drift of the three-body chip at coupler flux ``phi`` with coupling strengths g1 (Q1-TC), g2 (Q2-TC) in Hz."""
    tp = 2 * np.pi
    a_tc, a_q1, a_q2 = annihilators([3, 3, 3])
    freq_tc, anhar_tc = 8.1e9 * tp, -235e6 * tp
    f_tc = (freq_tc - anhar_tc) * transmon_factor(phi, 10.0, 0.36) + anhar_tc
    drift = f_tc * resonator(a_tc) + anhar_tc * duffing(a_tc)
    drift = drift + 6.189e9 * tp * resonator(a_q1) + (-286e6 * tp) * duffing(a_q1)
    drift = drift + 5.089e9 * tp * resonator(a_q2) + (-310e6 * tp) * duffing(a_q2)
    return drift + g1 * tp * int_XX(a_q1, a_tc) + g2 * tp * int_XX(a_q2, a_tc)


def tunable_coupler_model():
    """The three-body model of test/test_tunable_coupler.py:31-157 (subsystem order [coupler, Q1, Q2],
    :152): drift = sum of subsystem and coupling Hamiltonians (c3/model.py:430-447), tunable transmon
    frequency (freq - anhar) * factor + anhar (chip.py:378-383), dressed with eigh + reorder
    (model.py:453-534).  Returns the dressed drift, the dressed flux-line (z-drive) Hamiltonian and the
    eigenframe."""
    tp = 2 * np.pi
    dims = [3, 3, 3]
    a_tc, a_q1, a_q2 = annihilators(dims)
    freq_tc, anhar_tc = 8.1e9 * tp, -235e6 * tp
    phi_0 = 10.0
    f_tc = (freq_tc - anhar_tc) * transmon_factor(phi_0 * 0.23, phi_0, 0.36) + anhar_tc
    drift = f_tc * resonator(a_tc) + anhar_tc * duffing(a_tc)
    drift = drift + 6.189e9 * tp * resonator(a_q1) + (-286e6 * tp) * duffing(a_q1)
    drift = drift + 5.089e9 * tp * resonator(a_q2) + (-310e6 * tp) * duffing(a_q2)
    drift = drift + 142e6 * tp * int_XX(a_q1, a_tc) + 116e6 * tp * int_XX(a_q2, a_tc)   # Q1-Q2 strength is 0
    eigenframe, T = dressing_transform(drift)
    return {"h0": dress(T, drift), "hk_tc": dress(T, z_drive(a_tc)), "eigenframe": eigenframe, "transform": T,
            "hk_q1": dress(T, x_drive(a_q1)), "hk_q2": dress(T, x_drive(a_q2))}
