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


def dressing_transform(drift: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """eigh + reorder by largest overlap + sign fix (c3/model.py:453-502, ordered=True,
    the ``max_probabilities > 0.5`` branch)."""
    e, v = np.linalg.eigh(drift)
    v_sq = np.real(v * np.conj(v))
    if v_sq.max(axis=0).min() <= 0.5:
        raise RuntimeError("overly dressed states: fallback branch of reorder_frame not restated")
    reorder = (v_sq > 0.5).astype(np.float64)
    signed = np.sign(np.real(v)) * reorder
    eigenframe = reorder @ np.real(e)
    transform = v @ signed.T
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
