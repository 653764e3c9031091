"""GPU parity of the goal-function kernels against the CPU oracle
(oracle/c3_fid_oracle.py), through the reference-shaped API of c3_b200.fidelities and the C ABI.
Scalar goal values: absolute tolerance 1e-12 (they are O(1) sums of <= 256 products)."""
import numpy as np
import pytest
import torch

from oracle import c3_fid_oracle as fo
from oracle import c3_oracle as orc

pytestmark = pytest.mark.gpu
ATOL = 1e-12

X, Y, Id = fo.GATES["rxp"], fo.GATES["ryp"], fo.GATES["id"]
LEAKY = np.array([[0 + 0j, 1, 0], [1, 0, 0], [0, 0, 34345j]])


@pytest.fixture(scope="module")
def fid():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from c3_b200 import fidelities
    return fidelities


class _Instr:
    """Duck-typed Instruction: only get_ideal_gate is used by the *_set goal functions."""

    def __init__(self, gate):
        self.gate = gate

    def get_ideal_gate(self, dims, index=None):
        return self.gate


def test_reference_known_answers(fid):
    """test/test_fidelities.py:22-140 of the reference, run through the CUDA path."""
    f = lambda x: float(x)
    assert abs(f(fid.unitary_infid(X, X, dims=[2]))) < ATOL
    assert abs(f(fid.unitary_infid(X, Y, dims=[2])) - 1) < ATOL
    a = np.kron(X, Id)
    assert abs(f(fid.unitary_infid(a, a, index=[0, 1], dims=[2, 2]))) < ATOL
    assert abs(f(fid.unitary_infid(X, a, index=[0], dims=[2, 2]))) < ATOL
    assert abs(f(fid.unitary_infid(X, np.kron(Id, X), index=[1], dims=[2, 2]))) < ATOL
    assert abs(f(fid.unitary_infid(ideal=X, actual=LEAKY, index=[0], dims=[3]))) < ATOL
    assert abs(f(fid.unitary_infid(ideal=X, actual=np.kron(LEAKY, Id), index=[0], dims=[3, 2]))) < ATOL
    assert abs(f(fid.average_infid(X, X))) < ATOL
    assert abs(f(fid.average_infid(X, Y)) - 2.0 / 3) < ATOL
    assert abs(f(fid.average_infid(X, a, index=[0], dims=[2, 2]))) < ATOL
    assert abs(f(fid.average_infid(X, np.kron(Id, X), index=[1], dims=[2, 2]))) < ATOL
    assert abs(f(fid.average_infid(ideal=X, actual=LEAKY, index=[0], dims=[3]))) < ATOL
    assert abs(f(fid.average_infid(ideal=X, actual=np.kron(LEAKY, Id), index=[0], dims=[3, 2]))) < ATOL
    props = {"rxp": X, "ryp": Y}
    instrs = {"rxp": _Instr(X), "ryp": _Instr(Y)}
    assert abs(f(fid.unitary_infid_set(props, instrs, index=[0], dims=[2], n_eval=136))) < ATOL
    assert abs(f(fid.average_infid_set(props, instrs, index=[0], dims=[2]))) < ATOL


@pytest.mark.parametrize("dims,index", [([3], [0]), ([3, 3], [0, 1]), ([3, 3], [1]), ([3, 3, 3], [0, 2]), ([2, 2], [0, 1])])
def test_random_batched_parity(fid, dims, index):
    rng = np.random.default_rng(11)
    d, c, B = int(np.prod(dims)), 2 ** len(index), 37
    A = rng.normal(size=(B, d, d)) + 1j * rng.normal(size=(B, d, d))
    G = rng.normal(size=(c, c)) + 1j * rng.normal(size=(c, c))
    G /= np.linalg.norm(G)
    A /= np.linalg.norm(A, axis=(1, 2), keepdims=True)
    got_u = fid.unitary_infid(G, A, index, dims).cpu().numpy()
    got_a = fid.average_infid(G, A, index, dims).cpu().numpy()
    for b in range(B):
        assert abs(got_u[b] - fo.unitary_infid(G, A[b], index, dims)) < ATOL
        assert abs(got_a[b] - fo.average_infid(G, A[b], index, dims)) < ATOL
    if d <= 9:
        S = rng.normal(size=(5, d * d, d * d)) + 1j * rng.normal(size=(5, d * d, d * d))
        S /= np.linalg.norm(S, axis=(1, 2), keepdims=True)
        got = fid.lindbladian_unitary_infid(G, S, index, dims).cpu().numpy()
        for b in range(5):
            assert abs(got[b] - fo.lindbladian_unitary_infid(G, S[b], index, dims)) < ATOL
        if all(x == 2 for x in dims):
            got = fid.lindbladian_average_infid(G, S, index, dims).cpu().numpy()
            for b in range(5):
                assert abs(got[b] - fo.lindbladian_average_infid(G, S[b], index, dims)) < ATOL
        else:
            with pytest.raises(ValueError):
                fid.lindbladian_average_infid(G, S, index, dims)


def test_propagator_to_infid_chain(fid):
    """pwc_batch -> unitary_infid on the device against oracle propagators + oracle goal function."""
    from c3_b200 import propagation as prop, synth
    m = synth.two_transmon()
    sig = synth.controls(m, 6, 120)
    U = prop.pwc_batch(m.h0, m.hks, torch.as_tensor(sig).cuda(), 1e-11)
    want_U = orc.propagate_batch(m.h0, m.hks, sig, 1e-11)
    G = np.kron(fo.GATES["rx90p"], Id)
    got = fid.unitary_infid(G, U, index=[0, 1], dims=[3, 3]).cpu().numpy()
    for b in range(6):
        assert abs(got[b] - fo.unitary_infid(G, want_U[b], [0, 1], [3, 3])) < 1e-10
    seq = fid.average_infid_seq({"a": U[0], "b": U[1]}, {"a": _Instr(G), "b": _Instr(G)}, [0, 1], [3, 3])
    assert abs(float(seq) - fo.average_infid_seq({"a": want_U[0], "b": want_U[1]}, {"a": G, "b": G}, [0, 1], [3, 3])) < 1e-10


@pytest.mark.parametrize("d", [2, 9])
def test_orbit_infid_and_populations(fid, d):
    from c3_b200 import synth
    rng = np.random.default_rng(3)
    names = ["rx90p[0]", "rx90m[0]", "ry90p[0]", "ry90m[0]"]
    props = {}
    for i, n in enumerate(names):
        h = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
        w, v = np.linalg.eigh(h + h.conj().T)
        props[n] = (v * np.exp(-1j * 0.3 * w)) @ v.conj().T        # some unitary per native gate
    seqs = synth.single_length_RB(33, 7, rng=np.random.default_rng(0)) + [[]]
    got = float(fid.orbit_infid(props, seqs=seqs))
    assert abs(got - fo.orbit_infid(props, seqs)) < ATOL
    pops = fid.sequence_populations(props, seqs).cpu().numpy()
    Us = orc.evaluate_sequences(props, seqs)
    for s, U in enumerate(Us):
        assert np.allclose(pops[s], np.abs(U[:, 0]) ** 2, atol=ATOL)
    psi0 = rng.normal(size=d) + 1j * rng.normal(size=d)
    pops = fid.sequence_populations(props, seqs, psi_init=psi0.reshape(-1, 1)).cpu().numpy()
    for s, U in enumerate(Us):
        assert np.allclose(pops[s], fo.populations(U @ psi0, False), atol=1e-11)
    # default sequences come from numpy's global state like the reference
    np.random.seed(4)
    a = float(fid.orbit_infid(props, RB_number=5, RB_length=4))
    np.random.seed(4)
    b = fo.orbit_infid(props, synth.single_length_RB(5, 4))
    assert abs(a - b) < ATOL
    # shots / noise only perturb the deterministic value statistically
    torch.manual_seed(0)
    assert abs(float(fid.orbit_infid(props, seqs=seqs, shots=100000)) - got) < 0.02


def test_lindblad_sequence_populations(fid):
    rng = np.random.default_rng(8)
    d = 3
    sup = {}
    for n in ("a", "b"):
        h = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
        w, v = np.linalg.eigh(h + h.conj().T)
        u = (v * np.exp(-1j * w)) @ v.conj().T
        sup[n] = orc.tf_super(u)
    seqs = [["a"], ["a", "b", "a"], []]
    rho0 = np.zeros((d, d), complex)
    rho0[0, 0] = 1
    vec0 = rho0.T.reshape(-1, 1)
    pops = fid.sequence_populations(sup, seqs, psi_init=vec0, lindbladian=True).cpu().numpy()
    Us = orc.evaluate_sequences(sup, seqs)
    for s, U in enumerate(Us):
        assert np.allclose(pops[s], fo.populations(U @ vec0, True).ravel(), atol=ATOL)
        assert abs(pops[s].sum() - 1) < 1e-10


@pytest.mark.parametrize("average", [False, True])
def test_infid_gradient(fid, average):
    """Analytic cotangent vs torch autograd of the same closed form, then the whole GRAPE step
    signals -> propagators -> infidelity -> backward on the device vs the torch-CPU autograd oracle."""
    from c3_b200 import propagation as prop, synth
    from oracle import c3_grad_oracle as gorc
    rng = np.random.default_rng(2)
    dims, index = [3, 3], [0, 1]
    G = np.kron(fo.GATES["rx90p"], Id)
    U0 = rng.normal(size=(4, 9, 9)) + 1j * rng.normal(size=(4, 9, 9))
    w = rng.normal(size=4)
    Ut = torch.tensor(U0, device="cuda", requires_grad=True)
    L = (fid.unitary_infid_autograd(G, Ut, index, dims, average=average) * torch.as_tensor(w, device="cuda")).sum()
    L.backward()
    sel = torch.as_tensor(fid.comp_indices(dims, index).astype(np.int64))
    Uc = torch.tensor(U0, requires_grad=True)
    t = (Uc[:, sel][:, :, sel] * torch.as_tensor(G).conj()).sum(dim=(1, 2))
    inf = 1 - (t.abs() ** 2 / 4 + 1) / 5 if average else 1 - t.abs() ** 2 / 16
    (inf * torch.as_tensor(w)).sum().backward()
    assert abs(float(L) - float((inf * torch.as_tensor(w)).sum())) < 1e-12
    assert np.allclose(Ut.grad.cpu().numpy(), Uc.grad.numpy(), atol=1e-13)

    m = synth.two_transmon()
    sig = synth.controls(m, 3, 40)
    s_ref = torch.tensor(sig, dtype=torch.float64, requires_grad=True)
    U_ref = gorc.propagate_torch(m.h0, m.hks, s_ref, 1e-11)
    t = (U_ref[:, sel][:, :, sel] * torch.as_tensor(G).conj()).sum(dim=(1, 2))
    L_ref = (1 - (t.abs() ** 2 / 4 + 1) / 5 if average else 1 - t.abs() ** 2 / 16).mean()
    L_ref.backward()
    s = torch.tensor(sig, device="cuda", requires_grad=True)
    Ud = prop.pwc_batch_autograd(m.h0, m.hks, s, 1e-11)
    Ld = fid.unitary_infid_autograd(G, Ud, index, dims, average=average).mean()
    Ld.backward()
    assert abs(float(Ld) - float(L_ref)) < 1e-10
    g, g_ref = s.grad.cpu().numpy(), s_ref.grad.numpy()
    assert np.linalg.norm(g - g_ref) / np.linalg.norm(g_ref) < 1e-8
