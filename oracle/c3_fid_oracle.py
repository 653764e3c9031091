"""CPU oracle for SURVEY.md section 8f row f-3: the goal functions that consume the propagators.

TEST INFRASTRUCTURE ONLY (imported by tests/, __graft_entry__.smoke() and bench.py's CPU leg; the
product path never touches it).  numpy restatement, op for op, of

  c3/utils/qt_utils.py:10-44      pauli_basis
  c3/utils/qt_utils.py:178-193    projector
  c3/utils/tf_utils.py:326-364    tf_unitary_overlap
  c3/utils/tf_utils.py:367-376    tf_superoper_unitary_overlap
  c3/utils/tf_utils.py:380-385    tf_average_fidelity
  c3/utils/tf_utils.py:388-393    tf_superoper_average_fidelity
  c3/utils/tf_utils.py:396-401    tf_super_to_fid
  c3/utils/tf_utils.py:404-413    tf_choi_to_chi
  c3/utils/tf_utils.py:417-427    super_to_choi
  c3/utils/tf_utils.py:430-438    tf_project_to_comp
  c3/libraries/fidelities.py:152-183   unitary_infid          (:186-218 unitary_infid_set)
  c3/libraries/fidelities.py:221-249   lindbladian_unitary_infid
  c3/libraries/fidelities.py:288-311   average_infid          (:314-347 average_infid_set, :350-374 _seq)
  c3/libraries/fidelities.py:377-399   lindbladian_average_infid
  c3/libraries/fidelities.py:753-790   orbit_infid (shots=None, noise=None: the deterministic part)
  c3/experiment.py:603-624             populations

Pinned to the reference's own known answers (test/test_fidelities.py:22-140: X vs X -> 0, X vs Y -> 1
resp. 2/3, projections from 3 levels and from two-qubit spaces) in tests/test_fidelity_oracle.py.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

from . import c3_oracle as orc

Id = np.array([[1, 0], [0, 1]], dtype=np.complex128)
X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)

#: c3/libraries/constants.py:50-57 (the gates the reference's fidelity tests use)
GATES = {
    "id": np.array([[1, 0], [0, 1]], dtype=np.complex128),
    "rx90p": np.array([[1, -1j], [-1j, 1]], dtype=np.complex128) / np.sqrt(2),
    "rxp": np.array([[0, -1j], [-1j, 0]], dtype=np.complex128),
    "ryp": np.array([[0, -1], [1, 0]], dtype=np.complex128),
}


def np_kron_n(mats: Sequence[np.ndarray]) -> np.ndarray:
    """c3/utils/qt_utils.py:100-117."""
    out = np.eye(1, dtype=np.complex128)
    for m in mats:
        out = np.kron(out, m)
    return out


def expand_dims(op: np.ndarray, dim: int) -> np.ndarray:
    out = np.zeros([dim, dim], dtype=op.dtype)
    out[: op.shape[0], : op.shape[1]] = op
    return out


def pauli_basis(dims=(2,)) -> np.ndarray:
    paulis = [[expand_dims(P, dim) for P in (Id, X, Y, Z)] for dim in dims]
    result: List[list] = [[]]
    for pauli_set in paulis:
        result = [x + [y] for x in result for y in pauli_set]
    size = int(np.prod(np.array(dims) ** 2))
    B = np.zeros((size, size), dtype=complex)
    for idx, op_tuple in enumerate(result):
        op = np_kron_n(op_tuple)
        vec = np.reshape(np.transpose(op), [-1, 1])
        B[:, idx] = vec.T.conj()
    return B


def projector(dims, indices, outdims=None) -> np.ndarray:
    if outdims is None:
        outdims = [2] * len(dims)
    ids = []
    for index, dim in enumerate(dims):
        ids.append(np.eye(dim, outdims[index]) if index in indices else np.eye(dim, 1))
    return np_kron_n(ids)


def tf_project_to_comp(A, dims, index=None, to_super=False) -> np.ndarray:
    if not index:
        index = list(range(len(dims)))
    proj = projector(dims, index)
    if to_super:
        proj = np.kron(proj, proj)
    P = proj.astype(np.complex128)
    return P.T @ np.asarray(A, dtype=np.complex128) @ P


def tf_unitary_overlap(A, B, lvls=None) -> float:
    if lvls is None:
        lvls = B.shape[0]
    t = np.trace(A @ B.conj().T) / lvls
    return float(np.real(np.conj(t) * t))


def tf_superoper_unitary_overlap(A, B, lvls=None) -> float:
    if lvls is None:
        lvls = np.sqrt(B.shape[0])
    return float(np.abs(np.sqrt(np.trace(A @ B.conj().T).astype(complex)) / lvls) ** 2)


def super_to_choi(A) -> np.ndarray:
    s = int(np.sqrt(A.shape[0]))
    return np.reshape(np.transpose(np.reshape(A, [s] * 4), (3, 1, 2, 0)), A.shape)


def tf_choi_to_chi(U, dims=None) -> np.ndarray:
    if dims is None:
        dims = [np.sqrt(U.shape[0])]
    B = pauli_basis([2] * len(dims))
    return B.conj().T @ U @ B


def tf_super_to_fid(err, lvls) -> float:
    lambda_chi = tf_choi_to_chi(super_to_choi(err), dims=lvls)
    d = 2 ** len(lvls)
    return float(np.abs((lambda_chi[0, 0] / d + 1) / (d + 1)))


def tf_average_fidelity(A, B, lvls=None) -> float:
    if lvls is None:
        lvls = [B.shape[0]]
    Lambda = A.conj().T @ B
    return tf_super_to_fid(orc.tf_super(Lambda), lvls)


def tf_superoper_average_fidelity(A, B, lvls=None) -> float:
    if lvls is None:
        lvls = np.sqrt(B.shape[0])
    lambda_super = tf_project_to_comp(A, lvls, to_super=True).conj().T @ B
    return tf_super_to_fid(lambda_super, lvls)


def unitary_infid(ideal, actual, index=None, dims=None) -> float:
    if index is None:
        index = list(range(len(dims)))
    actual_comp = tf_project_to_comp(actual, dims=dims, index=index)
    return 1 - tf_unitary_overlap(actual_comp, np.asarray(ideal, dtype=np.complex128), lvls=2 ** len(index))


def average_infid(ideal, actual, index=(0,), dims=(2,)) -> float:
    actual_comp = tf_project_to_comp(actual, dims=list(dims), index=list(index))
    return 1 - tf_average_fidelity(actual_comp, np.asarray(ideal, dtype=np.complex128), lvls=[2] * len(index))


def lindbladian_unitary_infid(ideal, actual, index=(0,), dims=(2,)) -> float:
    U_ideal = orc.tf_super(np.asarray(ideal, dtype=np.complex128))
    actual_comp = tf_project_to_comp(actual, dims=list(dims), index=list(index), to_super=True)
    return 1 - tf_superoper_unitary_overlap(actual_comp, U_ideal, lvls=2 ** len(index))


def lindbladian_average_infid(ideal, actual, index=(0,), dims=(2,)) -> float:
    U_ideal = orc.tf_super(np.asarray(ideal, dtype=np.complex128))
    actual_comp = tf_project_to_comp(actual, dims=list(dims), index=list(index), to_super=True)
    return 1 - tf_superoper_average_fidelity(actual_comp, U_ideal, lvls=list(dims))


def _set_mean(fn, propagators: Dict, ideals: Dict, index, dims) -> float:
    return float(np.mean([fn(ideals[g], U, index, dims) for g, U in propagators.items()]))


def unitary_infid_set(propagators: Dict, ideals: Dict, index, dims) -> float:
    """ideals[gate] plays instructions[gate].get_ideal_gate(dims, index)."""
    return _set_mean(unitary_infid, propagators, ideals, index, dims)


def average_infid_set(propagators: Dict, ideals: Dict, index, dims) -> float:
    return _set_mean(average_infid, propagators, ideals, index, dims)


def average_infid_seq(propagators: Dict, ideals: Dict, index, dims) -> float:
    fid = 1.0
    for g, U in propagators.items():
        fid *= 1 - average_infid(ideals[g], U, index, dims)
    return 1 - fid


def populations(state: np.ndarray, lindbladian: bool) -> np.ndarray:
    """c3/experiment.py:603-624; tf_vec_to_dm = transpose(reshape(vec, [d, d])) (tf_utils.py:305-308)."""
    state = np.asarray(state)
    if lindbladian:
        d = int(round(np.sqrt(state.shape[0])))
        rho = np.reshape(state, [d, d]).T
        return np.real(np.diag(rho)).reshape(-1, 1)
    return np.abs(state) ** 2


def orbit_infid(propagators: Dict, seqs: list, lindbladian: bool = False) -> float:
    """Deterministic part of orbit_infid (shots=None, noise=None): mean over sequences of
    1 - |<0| U_seq |0>|^2 with psi_init = basis(dim, 0)."""
    Us = orc.evaluate_sequences(propagators, seqs)
    infids = []
    for U in Us:
        dim = U.shape[0]
        psi = np.zeros((dim, 1), dtype=np.complex128)
        psi[0, 0] = 1.0
        psi_actual = U @ psi
        infids.append(1 - np.abs(psi_actual[0, 0]) ** 2)
    return float(np.mean(infids))
