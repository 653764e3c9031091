"""Deterministic synthetic models and control envelopes for benchmarks and tests
(SURVEY.md section 8d).  Host-side numpy; none of this is on the timed path.

The models follow the reference's own test chips (test/test_model.py:129-187 two transmons,
test/one_qubit.hjson single qubit, test/test_tunable_coupler.py:31-151 three-body tunable
coupler): Duffing oscillators with XX couplings, dressed by diagonalising the drift and
re-ordering eigenvectors by largest overlap (c3/model.py:453-534).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

TWO_PI = 2.0 * np.pi


@dataclass
class SynthModel:
    name: str
    dims: List[int]
    h0: np.ndarray            # [d,d] dressed drift (rad/s)
    hks: np.ndarray           # [K,d,d] dressed control Hamiltonians
    col_ops: np.ndarray       # [C,d,d] dressed collapse operators (may be empty)
    drive_freqs: np.ndarray   # [K] carrier angular frequencies for the synthetic envelopes

    @property
    def d(self) -> int:
        return int(self.h0.shape[0])

    @property
    def K(self) -> int:
        return int(self.hks.shape[0])


def _lift(op: np.ndarray, idx: int, dims: Sequence[int]) -> np.ndarray:
    out = np.eye(1)
    for j, dj in enumerate(dims):
        out = np.kron(out, op if j == idx else np.eye(dj))
    return out


def _lowering_ops(dims: Sequence[int]) -> List[np.ndarray]:
    return [_lift(np.diag(np.sqrt(np.arange(1, dj)), k=1), j, dims) for j, dj in enumerate(dims)]


def _dress(drift: np.ndarray):
    e, v = np.linalg.eigh(drift)
    prob = np.abs(v) ** 2
    if prob.max(axis=0).min() > 0.5:
        order = (prob > 0.5).astype(float)
    else:  # greedy assignment for strongly dressed states
        order = np.zeros_like(prob)
        pc = prob.copy()
        for _ in range(pc.shape[1]):
            i, j = np.unravel_index(np.argmax(pc), pc.shape)
            pc[i, :] = 0
            pc[:, j] = 0
            order[i, j] = 1
    sgn = np.sign(np.real(v))
    sgn[sgn == 0] = 1.0
    T = (v @ (sgn * order).T).astype(np.complex128)
    return T


def duffing_model(name: str, dims: Sequence[int], freqs: Sequence[float], anhars: Sequence[float],
                  couplings: Sequence[tuple], drives: Sequence[int], t1: Optional[Sequence[float]] = None,
                  t2star: Optional[Sequence[float]] = None, lo_detuning: float = 50e6) -> SynthModel:
    """freqs/anhars in Hz; couplings = [(i, j, g_Hz)]; drives = subsystem index per control line."""
    dims = list(dims)
    a = _lowering_ops(dims)
    d = int(np.prod(dims))
    drift = np.zeros((d, d), dtype=np.complex128)
    for j in range(len(dims)):
        n = a[j].conj().T @ a[j]
        drift += TWO_PI * freqs[j] * n
        if dims[j] > 2:
            drift += TWO_PI * anhars[j] * 0.5 * (n - np.eye(d)) @ n
    for (i, j, g) in couplings:
        drift += TWO_PI * g * (a[i].conj().T + a[i]) @ (a[j].conj().T + a[j])
    T = _dress(drift)
    dress = lambda x: T.conj().T @ x @ T
    h0 = dress(drift)
    hks = np.stack([dress(a[j].conj().T + a[j]) for j in drives]) if len(drives) else np.zeros((0, d, d), complex)
    cols = []
    if t1 is not None:
        for j in range(len(dims)):
            L = (1.0 / t1[j]) ** 0.5 * a[j]
            if t2star is not None:
                L = L + (0.5 / t2star[j]) ** 0.5 * 2.0 * (a[j].conj().T @ a[j])
            cols.append(dress(L.astype(np.complex128)))
    col_ops = np.stack(cols) if cols else np.zeros((0, d, d), complex)
    drive_freqs = np.array([TWO_PI * (freqs[j] + lo_detuning) for j in drives])
    return SynthModel(name, dims, h0, hks, col_ops, drive_freqs)


def two_transmon(levels: int = 3) -> SynthModel:
    """d = levels^2 (9 for the headline config): the reference's two-transmon test chip."""
    return duffing_model("two_transmon", [levels, levels], [5.0e9, 5.6e9], [-210e6, -240e6], [(0, 1, 20e6)],
                         drives=[0, 1], t1=[27e-6, 23e-6], t2star=[39e-6, 31e-6])


def one_qubit(levels: int = 3) -> SynthModel:
    """test/one_qubit.hjson: freq 3.82 GHz, anharmonicity -229 MHz, one drive line."""
    return duffing_model("one_qubit", [levels], [3.82e9], [-229e6], [], drives=[0], t1=[27e-6], t2star=[39e-6])


def tunable_coupler(levels: int = 3) -> SynthModel:
    """Three-body d = 27 chip with the parameters of test/test_tunable_coupler.py:31-55."""
    return duffing_model("tunable_coupler", [levels] * 3, [6.189e9, 5.089e9, 8.1e9], [-286e6, -310e6, -235e6],
                         [(0, 2, 142e6), (1, 2, 116e6)], drives=[0, 1, 2], t1=[23e-6, 70e-6, 15e-6],
                         t2star=[27e-6, 50e-6, 7e-6])


def controls(model: SynthModel, B: int, N: int, dt: float = 1e-11, seed: int = 1234,
             single_line: bool = False, b_offset: int = 0) -> np.ndarray:
    """signals[B,K,N] (float64): Gaussian-windowed carriers
    c_k[b,n] = 2 pi 1e9 A_b exp(-(t-T/2)^2 / (2 (T/4)^2)) cos(w_k t + phi_b),  t = (n + 1/2) dt,
    A_b ~ U[0.2, 0.5], phi_b ~ U[0, 2 pi) from default_rng(seed + b).  ``b_offset`` shifts the
    batch index so that ranks of a sharded run draw disjoint, reproducible rows."""
    K = model.K
    t = (np.arange(N) + 0.5) * dt
    T = N * dt
    env = np.exp(-((t - T / 2) ** 2) / (2 * (T / 4) ** 2))
    out = np.empty((B, K, N), dtype=np.float64)
    for b in range(B):
        rng = np.random.default_rng(seed + b + b_offset)
        A = rng.uniform(0.2, 0.5, size=K)
        phi = rng.uniform(0.0, TWO_PI, size=K)
        for k in range(K):
            out[b, k] = TWO_PI * 1e9 * A[k] * env * np.cos(model.drive_freqs[k] * t + phi[k])
        if single_line and K > 1:
            out[b, 1:] = 0.0
    return out


def controls_fast(model: SynthModel, B: int, N: int, dt: float = 1e-11, seed: int = 1234,
                  b_offset: int = 0) -> np.ndarray:
    """Vectorised variant of :func:`controls` for large B (one generator seeded with
    ``seed + b_offset``; same distribution, different draws)."""
    K = model.K
    rng = np.random.default_rng(seed + b_offset)
    t = (np.arange(N) + 0.5) * dt
    T = N * dt
    env = np.exp(-((t - T / 2) ** 2) / (2 * (T / 4) ** 2))
    A = rng.uniform(0.2, 0.5, size=(B, K, 1))
    phi = rng.uniform(0.0, TWO_PI, size=(B, K, 1))
    w = model.drive_freqs.reshape(1, K, 1)
    return TWO_PI * 1e9 * A * env.reshape(1, 1, N) * np.cos(w * t.reshape(1, 1, N) + phi)


def rb_sequences(S: int, length: int, n_gates: int, seed: int = 0, mean_native: float = 2.25):
    """Index form of random gate sequences for the ORBIT-shaped workload (config 4):
    each of ``length`` Cliffords expands to 1-4 native gates (mean 2.25,
    c3/utils/qt_utils.py:586-589).  Returns (seq_idx [S,Lmax] int32, seq_len [S] int32)."""
    rng = np.random.default_rng(seed)
    per = rng.choice([1, 2, 3, 4], size=(S, length), p=[0.25, 0.40, 0.20, 0.15])
    lens = per.sum(axis=1).astype(np.int32)
    Lmax = int(lens.max())
    idx = rng.integers(0, n_gates, size=(S, Lmax)).astype(np.int32)
    return idx, lens


# ---- randomized-benchmarking sequences (c3/utils/qt_utils.py:448-498, 528-553) -------------------------
_RB_NATIVE = {"X": "rx90p", "x": "rx90m", "Y": "ry90p", "y": "ry90m"}
#: native-gate decomposition of the 24 single-qubit Cliffords C1..C24, applied left to right
#: (upper case = +90 degrees, lower case = -90 degrees); same table as qt_utils.cliffords_decomp
_RB_DECOMP = ("Xx YX xy YXX x Xyx XX yx Xy y X XYX YY yX XY yXX XYY XyX XXYY Yx xY Y xYY XYx").split()


def _rb_native_matrix(code: str) -> np.ndarray:
    s = 1.0 if code.isupper() else -1.0
    if code.upper() == "X":
        return np.array([[1, -1j * s], [-1j * s, 1]], dtype=np.complex128) / np.sqrt(2)
    return np.array([[1, -s], [s, 1]], dtype=np.complex128) / np.sqrt(2)


def clifford_decomp(n: int) -> List[str]:
    """Native gate names of Clifford C_n, n = 1..24."""
    return [_RB_NATIVE[c] for c in _RB_DECOMP[n - 1]]


def clifford_matrix(n: int) -> np.ndarray:
    """2x2 unitary of C_n: the product of its native gates, later gates on the left
    (equals c3/libraries/constants.py:96-121 CLIFFORDS["C<n>"])."""
    u = np.eye(2, dtype=np.complex128)
    for c in _RB_DECOMP[n - 1]:
        u = _rb_native_matrix(c) @ u
    return u


def inverse_clifford(seq: Sequence[int]) -> int:
    """The Clifford that returns the sequence to the identity up to a phase (qt_utils.inverseC, :480-491)."""
    op = np.eye(2, dtype=np.complex128)
    for n in seq:
        op = clifford_matrix(int(n)) @ op
    for i in range(1, 25):
        if abs(2 - abs(np.trace(clifford_matrix(i) @ op))) < 1e-4:
            return i
    raise RuntimeError("C3:ERROR: no inverting Clifford found")


def single_length_RB(RB_number: int, RB_length: int, target: int = 0, rng=None) -> List[List[str]]:
    """``RB_number`` sequences of ``RB_length`` Cliffords (the last one inverts the rest), as lists of
    native gate names ``"<gate>[<target>]"`` (qt_utils.single_length_RB, :448-477).  ``rng``: a
    ``numpy.random.Generator``/``RandomState``; default is numpy's global state like the reference."""
    draw = (lambda size: np.random.choice(24, size=size)) if rng is None else (
        (lambda size: rng.choice(24, size=size)))
    S = []
    for _ in range(RB_number):
        seq = np.asarray(draw(RB_length - 1)) + 1
        seq = np.append(seq, inverse_clifford(seq))
        gates: List[str] = []
        for n in seq:
            gates.extend(f"{g}[{target}]" for g in clifford_decomp(int(n)))
        S.append(gates)
    return S
