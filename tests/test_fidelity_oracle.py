"""CPU: the fidelity oracle (oracle/c3_fid_oracle.py) against the reference's own known answers
(test/test_fidelities.py:22-140), the closed forms the CUDA kernels use against the op-for-op
restatement, and the host-side helpers of c3_b200.fidelities / synth (no GPU needed)."""
import numpy as np
import pytest

from oracle import c3_fid_oracle as fo
from oracle import c3_oracle as orc

X, Y, Id = fo.GATES["rxp"], fo.GATES["ryp"], fo.GATES["id"]
LEAKY = np.array([[0 + 0j, 1, 0], [1, 0, 0], [0, 0, 34345j]])


def test_unitary_infid_known_answers():
    """test/test_fidelities.py:22-81."""
    assert abs(fo.unitary_infid(X, X, dims=[2])) < 1e-12
    assert fo.unitary_infid(X, Y, dims=[2]) == 1
    a = np.kron(X, Id)
    assert abs(fo.unitary_infid(a, a, index=[0, 1], dims=[2, 2])) < 1e-12
    assert abs(fo.unitary_infid(X, a, index=[0], dims=[2, 2])) < 1e-12
    assert abs(fo.unitary_infid(X, np.kron(Id, X), index=[1], dims=[2, 2])) < 1e-12
    assert abs(fo.unitary_infid(X, LEAKY[:, :] * np.array([1, 1, 0]), index=[0], dims=[3])) < 1e-12
    assert abs(fo.unitary_infid(X, LEAKY, index=[0], dims=[3])) < 1e-12
    assert abs(fo.unitary_infid(X, np.kron(LEAKY, Id), index=[0], dims=[3, 2])) < 1e-12


def test_average_infid_known_answers():
    """test/test_fidelities.py:84-140."""
    assert abs(fo.average_infid(X, X)) < 1e-12
    assert abs(fo.average_infid(X, Y) - 2.0 / 3) < 1e-12
    assert abs(fo.average_infid(X, np.kron(X, Id), index=[0], dims=[2, 2])) < 1e-12
    assert abs(fo.average_infid(X, np.kron(Id, X), index=[1], dims=[2, 2])) < 1e-12
    assert abs(fo.average_infid(X, LEAKY, index=[0], dims=[3])) < 1e-12
    assert abs(fo.average_infid(X, np.kron(LEAKY, Id), index=[0], dims=[3, 2])) < 1e-12


def test_set_known_answer():
    """test/test_fidelities.py:143-154: every gate against its own ideal -> 0."""
    assert abs(fo.unitary_infid_set({"rxp": X, "ryp": Y}, {"rxp": X, "ryp": Y}, [0], [2])) < 1e-12


@pytest.mark.parametrize("dims,index", [([3], [0]), ([3, 3], [0, 1]), ([3, 3], [1]), ([2, 3, 2], [0, 2]), ([2], [0])])
def test_closed_forms_match_op_for_op(dims, index):
    """The kernels evaluate every gate fidelity from ONE gathered overlap t; check those closed forms
    (c3_b200/csrc/fidelity.cuh) against the literal restatement of the reference's operator chain."""
    from c3_b200.fidelities import comp_indices, _super_sel
    rng = np.random.default_rng(5)
    d = int(np.prod(dims))
    c = 2 ** len(index)
    sel = comp_indices(dims, index)
    P = fo.projector(dims, index)
    assert np.array_equal(sel, np.argmax(P, axis=0)) and P.sum() == c
    A = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
    G = rng.normal(size=(c, c)) + 1j * rng.normal(size=(c, c))
    t = np.sum(A[np.ix_(sel, sel)] * G.conj())
    assert abs(fo.unitary_infid(G, A, index, dims) - (1 - abs(t) ** 2 / c ** 2)) < 1e-12
    assert abs(fo.average_infid(G, A, index, dims) - (1 - (abs(t) ** 2 / c + 1) / (c + 1))) < 1e-12
    S = rng.normal(size=(d * d, d * d)) + 1j * rng.normal(size=(d * d, d * d))
    sel2 = _super_sel(sel, d)
    ts = np.sum(S[np.ix_(sel2, sel2)] * np.kron(G, G.conj()).conj())
    assert abs(fo.lindbladian_unitary_infid(G, S, index, dims) - (1 - abs(ts) / c ** 2)) < 1e-12
    if all(x == 2 for x in dims) and len(index) == len(dims):
        assert abs(fo.lindbladian_average_infid(G, S, index, dims) - (1 - abs(np.conj(ts) / c + 1) / (c + 1))) < 1e-12


def test_clifford_tables_and_rb_sequences():
    """single_length_RB restatement (c3/utils/qt_utils.py:448-498): every sequence is the identity up to
    a phase; the 24 Clifford matrices form a group of order 24 modulo phases."""
    from c3_b200 import synth
    nat = {"rx90p": "X", "rx90m": "x", "ry90p": "Y", "ry90m": "y"}
    rng = np.random.default_rng(0)
    seqs = synth.single_length_RB(8, 20, rng=rng)
    assert len(seqs) == 8
    for s in seqs:
        u = np.eye(2, dtype=complex)
        for g in s:
            assert g.endswith("[0]")
            u = synth._rb_native_matrix(nat[g[:-3]]) @ u
        assert abs(abs(np.trace(u)) - 2) < 1e-9
    mats = [synth.clifford_matrix(n) for n in range(1, 25)]
    for i, a in enumerate(mats):
        for j, b in enumerate(mats):
            if i < j:
                assert abs(abs(np.trace(a.conj().T @ b)) - 2) > 1e-6   # pairwise distinct modulo phase
    assert np.allclose(mats[10], fo.GATES["rx90p"])                     # C11 = rx90p (constants.py:107)


def test_orbit_and_populations_oracle():
    rng = np.random.default_rng(1)
    q, _ = np.linalg.qr(rng.normal(size=(3, 3)) + 1j * rng.normal(size=(3, 3)))
    props = {"a": q, "b": q.conj().T}
    assert abs(fo.orbit_infid(props, [["a", "b"], []])) < 1e-12          # U^dag U = 1, empty = identity
    assert fo.orbit_infid(props, [["a"]]) == pytest.approx(1 - abs(q[0, 0]) ** 2)
    rho = rng.normal(size=(3, 3))
    assert np.allclose(fo.populations(rho.T.reshape(-1, 1), True).ravel(), np.diag(rho))
    assert np.allclose(fo.populations(q[:, :1], False).ravel(), np.abs(q[:, 0]) ** 2)
    assert orc.evaluate_sequences(props, [["a", "b"]])[0].shape == (3, 3)
