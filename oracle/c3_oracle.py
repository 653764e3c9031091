"""CPU oracle for the C3 piecewise-constant (PWC) propagator path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may
import it.  The shipped package (``c3_b200``) never imports anything from ``oracle/`` and
raises when its CUDA library is missing.

It restates, op for op in numpy, what the reference (q-optimize/c3 @ 48b7917e, pure Python on
TensorFlow) computes on the hot path.  All ``file:line`` citations are relative to the
reference checkout.  TensorFlow itself is an un-vendored dependency (``requirements.txt:17``,
``tensorflow>=2.15.0``) that is not installable in this image, so ``tf.linalg.expm`` is restated
from its published algorithm (Higham 2005 scaling-and-squaring Pade, the batched
"evaluate every order, select by norm" formulation) in :func:`expm_tf`.

Parity pin: the restatement is checked against the reference's own golden vectors
(``test/two_qubit_data.pickle``: closed 4x4 and Lindblad 16x16 propagators,
``test/transmon_expanded.pickle``: 20 partial propagators 24x24 with excitation cut,
``test/test_tf_utils.pickle``: Kronecker / superoperator helpers), re-exported as
``tests/golden/*.npz`` by ``tests/golden/make_golden.py``.  There is no golden vector for
d=9, D=81 or for norms in the squaring regime (||A||_1 > 5.37): there the oracle is
"parity unpinned" against TensorFlow itself and is cross-checked against scipy instead.
"""
from __future__ import annotations

import itertools
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np

# ----------------------------------------------------------------------------------------
# tf.linalg.expm  (external; call sites c3/libraries/propagation.py:440 and :584)
# ----------------------------------------------------------------------------------------

#: Higham-2005 1-norm thresholds for Pade orders 3, 5, 7, 9 and the order-13 scaling norm.
THETA = (1.495585217958292e-2, 2.539398330063230e-1, 9.504178996162932e-1, 2.097847961257068)
THETA13 = 5.371920351148152

PADE_B = {
    3: (120.0, 60.0, 12.0, 1.0),
    5: (30240.0, 15120.0, 3360.0, 420.0, 30.0, 1.0),
    7: (17297280.0, 8648640.0, 1995840.0, 277200.0, 25200.0, 1512.0, 56.0, 1.0),
    9: (17643225600.0, 8821612800.0, 2075673600.0, 302702400.0, 30270240.0, 2162160.0,
        110880.0, 3960.0, 90.0, 1.0),
    13: (64764752532480000.0, 32382376266240000.0, 7771770303897600.0, 1187353796428800.0,
         129060195264000.0, 10559470521600.0, 670442572800.0, 33522128640.0, 1323241920.0,
         40840800.0, 960960.0, 16380.0, 182.0, 1.0),
}


def _eye_like(a: np.ndarray) -> np.ndarray:
    return np.broadcast_to(np.eye(a.shape[-1], dtype=a.dtype), a.shape)


def _pade_uv_low(a: np.ndarray, m: int):
    """U (odd part) and V (even part) of the [m/m] Pade numerator, m in {3,5,7,9}."""
    b = PADE_B[m]
    ident = _eye_like(a)
    a2 = a @ a
    powers = [ident, a2]
    for _ in range(m // 2 - 1):
        powers.append(powers[-1] @ a2)
    inner = sum(b[2 * k + 1] * powers[k] for k in range(len(powers)))
    u = a @ inner
    v = sum(b[2 * k] * powers[k] for k in range(len(powers)))
    return u, v


def _pade_uv_13(a: np.ndarray):
    b = PADE_B[13]
    ident = _eye_like(a)
    a2 = a @ a
    a4 = a2 @ a2
    a6 = a4 @ a2
    inner_u = a6 @ (b[13] * a6 + b[11] * a4 + b[9] * a2) + b[7] * a6 + b[5] * a4 + b[3] * a2 + b[1] * ident
    u = a @ inner_u
    v = a6 @ (b[12] * a6 + b[10] * a4 + b[8] * a2) + b[6] * a6 + b[4] * a4 + b[2] * a2 + b[0] * ident
    return u, v


def expm_tf(matrix: np.ndarray) -> np.ndarray:
    """Batched matrix exponential the way ``tf.linalg.expm`` evaluates it for f64/c128.

    Every Pade order 3/5/7/9/13 is evaluated for the whole batch and the one matching each
    matrix's 1-norm is selected; order 13 is applied to ``A / 2**s`` with
    ``s = max(floor(log2(||A||_1 / theta13)), 0)`` and the result is squared ``s`` times.
    ``R = solve(V - U, V + U)``.  (used by
    c3/libraries/propagation.py:426-440 and :551-585.)
    """
    a = np.asarray(matrix)
    if a.dtype not in (np.complex128, np.float64):
        a = a.astype(np.complex128)
    shape = a.shape
    a = a.reshape((-1,) + shape[-2:])
    if a.shape[0] == 0 or a.shape[-1] == 0:
        return a.reshape(shape).copy()
    l1 = np.abs(a).sum(axis=-2).max(axis=-1)  # max column sum
    with np.errstate(divide="ignore"):
        squarings = np.maximum(np.floor(np.log(l1 / THETA13) / np.log(2.0)), 0.0)
    squarings = np.where(np.isfinite(squarings), squarings, 0.0)
    u3, v3 = _pade_uv_low(a, 3)
    u5, v5 = _pade_uv_low(a, 5)
    u7, v7 = _pade_uv_low(a, 7)
    u9, v9 = _pade_uv_low(a, 9)
    u13, v13 = _pade_uv_13(a / (2.0 ** squarings)[:, None, None])
    sel = l1[:, None, None]
    u = np.where(sel < THETA[0], u3, np.where(sel < THETA[1], u5, np.where(sel < THETA[2], u7,
        np.where(sel < THETA[3], u9, u13))))
    v = np.where(sel < THETA[0], v3, np.where(sel < THETA[1], v5, np.where(sel < THETA[2], v7,
        np.where(sel < THETA[3], v9, v13))))
    if not np.isfinite(l1.max()):
        return np.full(shape, np.nan, dtype=a.dtype)
    r = np.linalg.solve(v - u, v + u)
    max_sq = int(squarings.max())
    for i in range(max_sq):
        todo = (i < squarings)[:, None, None]
        r = np.where(todo, r @ r, r)
    return r.reshape(shape)


def pade_order_and_squarings(norm1: float):
    """Minimal Higham-2005 (order, squarings) for a 1-norm; the flop-count convention of
    SURVEY.md section 8(d)."""
    for m, th in zip((3, 5, 7, 9), THETA):
        if norm1 < th:
            return m, 0
    s = max(int(np.ceil(np.log2(norm1 / THETA13))), 0) if norm1 > 0 else 0
    return 13, s


PADE_MATMULS = {3: 2, 5: 3, 7: 4, 9: 5, 13: 6}


def algorithmic_flops_per_slice(d: int, K: int, m: int, s: int, lindblad_d: Optional[int] = None) -> float:
    """F(d,K,m,s) = 8 d^3 (M_m + s + 1) + 32/3 d^3 + assembly."""
    assembly = 4.0 * K * d * d if lindblad_d is None else 8.0 * lindblad_d * d
    return 8.0 * d ** 3 * (PADE_MATMULS[m] + s + 1) + (32.0 / 3.0) * d ** 3 + assembly


# ----------------------------------------------------------------------------------------
# c3/utils/tf_utils.py
# ----------------------------------------------------------------------------------------

def tf_matmul_left(dUs: np.ndarray) -> np.ndarray:
    """``tf.foldr(lambda a, x: matmul(a, x))`` over [dU_0 ... dU_{N-1}] -> dU_{N-1} ... dU_1 dU_0.

    c3/utils/tf_utils.py:120-129.  ``tf.foldr`` walks the list from the back with the last
    element as the initial accumulator ``a`` and calls ``fn(a, x)`` with ``x`` the next
    (earlier) element, so every earlier slice is multiplied on from the right: later slices
    end up on the left.  The golden propagators pin this order.
    """
    dUs = np.asarray(dUs)
    acc = dUs[-1]
    for i in range(dUs.shape[0] - 2, -1, -1):
        acc = acc @ dUs[i]
    return acc


def tf_matmul_right(dUs: np.ndarray) -> np.ndarray:
    """``tf.foldl(matmul)``: ((dU_0 dU_1) dU_2) ...  (c3/utils/tf_utils.py:132-141)."""
    dUs = np.asarray(dUs)
    acc = dUs[0]
    for i in range(1, dUs.shape[0]):
        acc = acc @ dUs[i]
    return acc


def _tf_matmul_n_even(odd: np.ndarray, even: np.ndarray) -> np.ndarray:
    """c3/utils/tf_utils.py:166-178."""
    return odd @ even


def _tf_matmul_n_odd(odd: np.ndarray, even: np.ndarray) -> np.ndarray:
    """c3/utils/tf_utils.py:181-193: the unpaired last ``even`` element is carried."""
    return np.concatenate([odd @ even[:-1], even[-1:]], axis=0)


def compute_folding_stack(n_steps: int) -> List[Callable]:
    """Per-level even/odd function list, c3/experiment.py:93-107."""
    stack = []
    n = n_steps
    while n > 1:
        stack.append(_tf_matmul_n_even if n % 2 == 0 else _tf_matmul_n_odd)
        n = int(np.ceil(n / 2))
    return stack


def tf_matmul_n(tensor_list: np.ndarray, folding_stack: Sequence[Callable]) -> np.ndarray:
    """Pairwise tree product, each level ``odd @ even`` (c3/utils/tf_utils.py:144-163)."""
    t = np.asarray(tensor_list)
    for func in folding_stack:
        t = func(t[1::2], t[0::2])
    return t[0]


def Id_like(A: np.ndarray) -> np.ndarray:
    """c3/utils/tf_utils.py:240-245."""
    A = np.asarray(A)
    return np.broadcast_to(np.eye(A.shape[-1], dtype=A.dtype), A.shape).copy()


def tf_kron(A: np.ndarray, B: np.ndarray) -> np.ndarray:
    """(Batched) Kronecker product, c3/utils/tf_utils.py:257-267."""
    A = np.asarray(A)
    B = np.asarray(B)
    res = A[..., :, None, :, None] * B[..., None, :, None, :]
    return res.reshape(res.shape[:-4] + (A.shape[-2] * B.shape[-2], A.shape[-1] * B.shape[-1]))


def tf_spre(A: np.ndarray) -> np.ndarray:
    """A (x) I, c3/utils/tf_utils.py:271-274."""
    return tf_kron(A, Id_like(A))


def tf_spost(A: np.ndarray) -> np.ndarray:
    """I (x) A^T, c3/utils/tf_utils.py:277-280."""
    A = np.asarray(A)
    return tf_kron(Id_like(A), np.swapaxes(A, -1, -2))


def tf_super(A: np.ndarray) -> np.ndarray:
    """spre(A) @ spost(A^dagger), c3/utils/tf_utils.py:284-289."""
    A = np.asarray(A)
    return tf_spre(A) @ tf_spost(np.conj(np.swapaxes(A, -1, -2)))


# ----------------------------------------------------------------------------------------
# c3/model.py: excitation cutter
# ----------------------------------------------------------------------------------------

def make_ex_cutter(dims: Sequence[int], max_excitations: int) -> np.ndarray:
    """Projector rows for product-state labels with sum <= max_excitations
    (c3/model.py:141-159 labels, :198-216 cutter)."""
    labels = list(itertools.product(*[range(d) for d in dims]))
    rows = []
    for i, lab in enumerate(labels):
        if sum(lab) <= max_excitations:
            line = np.zeros(len(labels))
            line[i] = 1
            rows.append(line)
    return np.array(rows).astype(np.complex128)


def cut_excitations(cutter: np.ndarray, op: np.ndarray) -> np.ndarray:
    """P op P^T (c3/model.py:218-220)."""
    return cutter @ op @ cutter.T


def blowup_excitations(cutter: np.ndarray, op: np.ndarray) -> np.ndarray:
    """P^T op P (c3/model.py:222-224)."""
    return cutter.T @ op @ cutter


# ----------------------------------------------------------------------------------------
# c3/libraries/propagation.py
# ----------------------------------------------------------------------------------------

def tf_propagation_vectorized(h0, hks, cflds_t, dt, expm=expm_tf) -> np.ndarray:
    """dU_n = expm(-i (h0 + sum_k c_k[n] H_k) dt); c3/libraries/propagation.py:426-440."""
    if hks is not None and cflds_t is not None:
        cflds = np.asarray(cflds_t).astype(np.complex128)[:, :, None, None]  # [K,n,1,1]
        hks_ = np.asarray(hks).astype(np.complex128)[:, None]               # [K,1,d,d]
        h0 = np.asarray(h0).astype(np.complex128)
        if h0.ndim < 3:
            h0 = h0[None]
        h = h0 + (cflds * hks_).sum(axis=0)
    else:
        h = np.asarray(h0).astype(np.complex128)
    dh = -1.0j * h * complex(dt)
    return expm(dh)


def lindblad_superop(h: np.ndarray, col_ops: np.ndarray) -> np.ndarray:
    """L = -i (H (x) I - I (x) H^T) + sum_c (L (x) L* - 1/2 L^dag L (x) I - 1/2 I (x) (L^dag L)^T).

    Built exactly like c3/libraries/propagation.py:563-582 (Kronecker temporaries, then matmuls
    with adjoint flags).  ``h`` is [n,d,d]; ``col_ops`` is [C,d,d].
    """
    h = np.asarray(h).astype(np.complex128)
    col_ops = np.asarray(col_ops).astype(np.complex128)
    d = h.shape[-1]
    h_id = np.broadcast_to(np.eye(d, dtype=np.complex128), h.shape)
    l_s = tf_kron(h, h_id)
    r_s = tf_kron(h_id, np.swapaxes(h, -1, -2))
    lind_op = -1j * (l_s - r_s)
    c_id = np.broadcast_to(np.eye(d, dtype=np.complex128), col_ops.shape)
    l_col = tf_kron(col_ops, c_id)
    r_col = tf_kron(c_id, np.swapaxes(col_ops, -1, -2))
    adj = lambda x: np.conj(np.swapaxes(x, -1, -2))
    super_clp = l_col @ adj(r_col)
    anticom_l = 0.5 * (adj(l_col) @ l_col)
    anticom_r = 0.5 * (r_col @ adj(r_col))
    clp = (super_clp - anticom_l - anticom_r).sum(axis=0)[None]
    return lind_op + clp


def tf_propagation_lind(h0, hks, col_ops, cflds_t, dt, expm=expm_tf) -> np.ndarray:
    """dU_n = expm(L_n dt) with L_n the Lindblad superoperator; propagation.py:551-585."""
    if hks is not None and cflds_t is not None:
        cflds = np.asarray(cflds_t).astype(np.complex128)[:, :, None, None]
        hks_ = np.asarray(hks).astype(np.complex128)[:, None]
        h = np.asarray(h0).astype(np.complex128)[None] + (cflds * hks_).sum(axis=0)
    else:
        h = np.asarray(h0).astype(np.complex128)
    return expm(lindblad_superop(h, col_ops) * complex(dt))


def tf_batch_propagate(hamiltonian, hks, signals, dt, batch_size, col_ops=None, lindbladian=False,
                       expm=expm_tf) -> np.ndarray:
    """Chunk the TIME axis into ceil(N / batch_size) pieces, propagate each, concatenate
    (c3/libraries/propagation.py:460-515)."""
    batch_size = int(batch_size)
    out = []
    if signals is not None:
        signals = np.asarray(signals)
        n = signals.shape[1]
        for i in range(int(np.ceil(n / batch_size))):
            x = signals[:, i * batch_size:(i + 1) * batch_size]
            if lindbladian:
                out.append(tf_propagation_lind(hamiltonian, hks, col_ops, x, dt, expm=expm))
            else:
                out.append(tf_propagation_vectorized(hamiltonian, hks, x, dt, expm=expm))
    else:
        hamiltonian = np.asarray(hamiltonian)
        n = hamiltonian.shape[0]
        for i in range(int(np.ceil(n / batch_size))):
            x = hamiltonian[i * batch_size:(i + 1) * batch_size]
            if lindbladian:
                # dead in the reference (tf.cast(None) at propagation.py:552); kept meaningful here
                out.append(tf_propagation_lind(x, None, col_ops, None, dt, expm=expm))
            else:
                out.append(tf_propagation_vectorized(x, None, None, dt, expm=expm))
    return np.concatenate(out, axis=0)


def pwc(model, gen, instr, folding_stack, batch_size=None, expm=expm_tf) -> Dict:
    """One gate: signals -> (h0, hks, signals, dt) -> dUs -> U  (propagation.py:258-341).

    ``model`` / ``gen`` are duck-typed exactly as the reference uses them.
    """
    signal = gen.generate_signals(instr)
    ts = []
    if model.controllability:
        h0, hctrls = model.get_Hamiltonians()
        signals, hks = [], []
        for key in signal:
            signals.append(np.asarray(signal[key]["values"]))
            ts = np.asarray(signal[key]["ts"])
            hks.append(np.asarray(hctrls[key]))
        signals = np.asarray(signals).astype(np.complex128)
        hks = np.asarray(hks).astype(np.complex128)
    else:
        h0 = np.asarray(model.get_Hamiltonian(signal))
        ts_list = np.asarray([np.asarray(sig["ts"])[1:] for sig in signal.values()])
        ts = ts_list.mean(axis=0)
        hks = None
        signals = None
        if not np.all(ts_list.var(axis=0) < 1e-5 * (ts[1] - ts[0])):
            raise Exception("C3Error:Something with the times happend.")
        if not np.all(np.var(ts[1:] - ts[:-1]) < 1e-5 * (ts[1] - ts[0])):
            raise Exception("C3Error:Something with the times happend.")
    dt = ts[1] - ts[0]
    if batch_size is None:
        batch_size = len(ts)
    if model.lindbladian:
        col_ops = [np.asarray(c) for c in model.get_Lindbladians()]
        if model.max_excitations:
            cutter = np.asarray(model.ex_cutter)
            col_ops = [cutter @ c @ cutter.T for c in col_ops]
        dUs = tf_batch_propagate(h0, hks, signals, dt, batch_size, col_ops=np.asarray(col_ops),
                                 lindbladian=True, expm=expm)
    else:
        dUs = tf_batch_propagate(h0, hks, signals, dt, batch_size, expm=expm)
    U = tf_matmul_n(dUs, folding_stack)
    if model.max_excitations:
        cutter = np.asarray(model.ex_cutter)
        U = blowup_excitations(cutter, tf_matmul_left(dUs))
        dUs = np.stack([blowup_excitations(cutter, x) for x in dUs])
    return {"U": U, "dUs": dUs, "ts": ts}


def evaluate_sequences(propagators: Dict, sequences: list) -> list:
    """U_seq per gate-name sequence via tf_matmul_left; empty -> identity
    (c3/libraries/propagation.py:588-627)."""
    first = np.asarray(list(propagators.values())[0])
    dim = first.shape[0]
    out = []
    for seq in sequences:
        if len(seq) == 0:
            out.append(np.eye(dim, dtype=first.dtype))
        else:
            us = np.asarray([np.asarray(propagators[g]) for g in seq]).astype(np.complex128)
            out.append(tf_matmul_left(us))
    return out


# ----------------------------------------------------------------------------------------
# Batched convenience (what B serial reference calls compute) -- used by tests and bench
# ----------------------------------------------------------------------------------------

def propagate_batch(h0, hks, signals_bkn, dt, col_ops=None, lindbladian=False, expm=expm_tf,
                    return_dUs=False):
    """B independent reference calls: for each b, tf_batch_propagate + tf_matmul_n.

    ``signals_bkn`` is [B,K,N].  Returns U[B,D,D] (and dUs[B,N,D,D] if asked).
    """
    signals_bkn = np.asarray(signals_bkn)
    B, K, N = signals_bkn.shape
    stack = compute_folding_stack(N)
    Us, dUs_all = [], []
    for b in range(B):
        dUs = tf_batch_propagate(h0, hks, signals_bkn[b], dt, N, col_ops=col_ops,
                                 lindbladian=lindbladian, expm=expm)
        Us.append(tf_matmul_n(dUs, stack))
        if return_dUs:
            dUs_all.append(dUs)
    if return_dUs:
        return np.stack(Us), np.stack(dUs_all)
    return np.stack(Us)


# ----------------------------------------------------------------------------------------
# Frame rotation, dephasing channel and the gate loop around pwc
# (c3/model.py:536-578, 597-639; c3/experiment.py:440-534)
# ----------------------------------------------------------------------------------------

def frame_rotation(ann_opers: Sequence[np.ndarray], line_to_index: Dict[str, int], t_final, freqs: Dict, framechanges: Dict,
                   expm=None) -> np.ndarray:
    """FR = expm( sum_line 1j a_q^dag a_q (freq_line t_final + framechange_line) )  (c3/model.py:536-578).
    ``line_to_index[line]`` is the subsystem the line drives (``couplings[line].connected[0]`` in the reference)."""
    expm = expm or (lambda a: expm_tf(a[None])[0])
    dim = ann_opers[0].shape[0]
    exponent = np.zeros((dim, dim), dtype=np.complex128)
    if len(freqs) == 0:
        return np.eye(dim, dtype=np.complex128)
    for line in freqs:
        a = np.asarray(ann_opers[line_to_index[line]])
        num = a.T.conj() @ a
        exponent = exponent + 1.0j * num * (freqs[line] * t_final + framechanges[line])
    return expm(exponent)


def dephasing_channel(ann_opers: Sequence[np.ndarray], line_to_index: Dict[str, int], t_final, amps: Dict,
                      dephasing_strength: float, expm=None) -> np.ndarray:
    """prod_line ((1 - p) Id + p Z_line) with the ELEMENTWISE product the reference uses (c3/model.py:597-639;
    ``deph_ch * (...)`` on tf tensors is elementwise), p = t_final amp strength, Z = super(expm(1j pi a^dag a))."""
    expm = expm or (lambda a: expm_tf(a[None])[0])
    dim = ann_opers[0].shape[0]
    Id = tf_super(np.eye(dim, dtype=np.complex128))
    ch = Id
    for line in amps:
        a = np.asarray(ann_opers[line_to_index[line]])
        num = a.T.conj() @ a
        Z = tf_super(expm(1.0j * num * np.pi))
        p = t_final * amps[line] * dephasing_strength
        if np.real(p) > 1 or np.real(p) < 0:
            raise ValueError(f"Dephasing channel strength {p} is outside [0,1] range")
        ch = ch * ((1 - p) * Id + p * Z)
    return ch


def compute_propagators(model, generator, instructions: Dict, sim_res: float, gate_ids=None, batch_size=None,
                        use_control_fields: bool = True, expm=expm_tf):
    """The gate loop of Experiment.compute_propagators (c3/experiment.py:440-534): pwc per gate, then FR U (or SFR U for
    Lindbladian models) and the dephasing channel from the left.  Returns (propagators, partial_propagators)."""
    propagators, partial = {}, {}
    for gate in (gate_ids if gate_ids is not None else instructions.keys()):
        if gate not in instructions:
            raise Exception(f"C3:Error: Gate '{gate}' is not defined. Available gates are:\n {list(instructions.keys())}.")
        instr = instructions[gate]
        model.controllability = use_control_fields
        steps = int((instr.t_end - instr.t_start) * sim_res)
        res = pwc(model, generator, instr, compute_folding_stack(steps), batch_size, expm=expm)
        U, dUs = res["U"], res["dUs"]
        if model.use_FR:
            freqs, framechanges = {}, {}
            for line, ctrls in instr.comps.items():
                offset = 0.0
                for ctrl in ctrls.values():
                    if "freq_offset" in ctrl.params.keys():
                        if np.asarray(ctrl.params["amp"].get_value()) != 0.0:
                            offset = float(np.asarray(ctrl.params["freq_offset"].get_value()))
                freqs[line] = complex(float(np.asarray(ctrls["carrier"].params["freq"].get_value())) + offset)
                framechanges[line] = complex(float(np.asarray(ctrls["carrier"].params["framechange"].get_value())))
            t_final = complex(instr.t_end - instr.t_start)
            FR = np.asarray(model.get_Frame_Rotation(t_final, freqs, framechanges))
            U = (tf_super(FR) if model.lindbladian else FR) @ U
        if model.dephasing_strength != 0.0:
            if not model.lindbladian:
                raise ValueError("Dephasing can only be added when lindblad is on.")
            amps = {}
            for line in instr.comps:
                amp, _ = generator.devices["awg"].get_average_amp()
                amps[line] = complex(float(np.asarray(amp)))
            U = np.asarray(model.get_dephasing_channel(complex(instr.t_end - instr.t_start), amps)) @ U
        propagators[gate] = U
        partial[gate] = dUs
    return propagators, partial
