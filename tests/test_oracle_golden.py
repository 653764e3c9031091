"""Pin the CPU oracle to the reference's own golden vectors (no GPU needed).

Mirrors test/test_two_qubits.py:46-62,193-213, test/test_transmon_expanded.py:252-283,
test/test_tf_utils.py:81-111 and test/test_exp.py:10-32 of the reference, but with a much
tighter tolerance than the reference's 6 decimals: the stored arrays carry full fp64.
"""
import numpy as np
import pytest
import scipy.linalg

from oracle import c3_oracle as orc
from conftest import rel_fro

TOL = 1e-12


def test_closed_propagator_matches_reference_pickle(golden_two_qubit):
    g = golden_two_qubit
    dt = g["ts"][1] - g["ts"][0]
    N = g["signals"].shape[1]
    dUs = orc.tf_batch_propagate(g["hdrift"], g["hks"], g["signals"], dt, N)
    U_tree = orc.tf_matmul_n(dUs, orc.compute_folding_stack(N))
    U_seq = orc.tf_matmul_left(dUs)
    assert rel_fro(U_tree, g["propagator"]) < TOL
    assert rel_fro(U_seq, g["propagator"]) < TOL
    # product order matters: the reversed product must NOT match
    assert rel_fro(orc.tf_matmul_right(dUs), g["propagator"]) > 1e-3


def test_lindblad_propagator_matches_reference_pickle(golden_two_qubit):
    g = golden_two_qubit
    dt = g["ts"][1] - g["ts"][0]
    N = g["signals"].shape[1]
    # propagate_batch_size = 360 as in test/test_two_qubits.py:200 (exercises time chunking)
    dUs = orc.tf_batch_propagate(g["hdrift"], g["hks"], g["signals"], dt, 360,
                                 col_ops=g["col_ops"], lindbladian=True)
    assert dUs.shape == (N, 16, 16)
    U = orc.tf_matmul_n(dUs, orc.compute_folding_stack(N))
    assert rel_fro(U, g["lindblad_propagator"]) < TOL


def test_rebuilt_two_qubit_model_matches_pickle(golden_two_qubit):
    g = golden_two_qubit
    assert rel_fro(g["h0_rebuilt"], g["hdrift"]) < 1e-14
    assert rel_fro(g["hks_rebuilt"], g["hks"]) < 1e-12


@pytest.mark.parametrize("q", ["q1", "q2"])
def test_transmon_expanded_partial_propagators(golden_transmon, q):
    g = golden_transmon
    cutter = orc.make_ex_cutter(g["dims"], int(g["max_excitations"]))
    assert cutter.shape == (14, 24)
    hs = np.stack([orc.cut_excitations(cutter, h) for h in g[f"hamiltonians_{q}"]])
    ts = g[f"ts_{q}"][1:]
    dt = ts[1] - ts[0]
    dUs = orc.tf_batch_propagate(hs, None, None, dt, hs.shape[0])
    dUs_big = np.stack([orc.blowup_excitations(cutter, x) for x in dUs])
    assert rel_fro(dUs_big, g[f"partial_propagators_{q}"]) < 1e-12
    U = orc.blowup_excitations(cutter, orc.tf_matmul_left(dUs))
    # frame rotation is the identity for this test (carrier freq 0, framechange 0)
    assert rel_fro(U, g[f"propagators_{q}"]) < 1e-11


def test_pwc_duck_typed_hlist_mode(golden_transmon):
    """pwc() with use_control_fields=False and max_excitations (propagation.py:294-339)."""
    g = golden_transmon
    cutter = orc.make_ex_cutter(g["dims"], 4)

    class M:
        controllability = False
        lindbladian = False
        max_excitations = 4
        ex_cutter = cutter

        def get_Hamiltonian(self, signal):
            return np.stack([orc.cut_excitations(cutter, h) for h in g["hamiltonians_q1"]])

    class G:
        def generate_signals(self, instr):
            # the reference's generator returns one more sample than there are slices
            ts = g["ts_q1"]
            return {"Qubit1": {"values": np.zeros_like(ts), "ts": ts}}

    n = g["hamiltonians_q1"].shape[0]
    res = orc.pwc(M(), G(), None, orc.compute_folding_stack(n))
    assert rel_fro(res["dUs"], g["partial_propagators_q1"]) < 1e-12
    assert rel_fro(res["U"], g["propagators_q1"]) < 1e-11


def test_superoperator_helpers(golden_tf_utils):
    g = golden_tf_utils
    for i in (0, 1):
        np.testing.assert_allclose(orc.tf_kron(g[f"tf_kron_{i}_inA"], g[f"tf_kron_{i}_inB"]),
                                   g[f"tf_kron_{i}_desired"], rtol=1e-13)
        np.testing.assert_allclose(orc.tf_spre(g[f"tf_spre_{i}_in"]), g[f"tf_spre_{i}_desired"], rtol=1e-13)
        np.testing.assert_allclose(orc.tf_spost(g[f"tf_spost_{i}_in"]), g[f"tf_spost_{i}_desired"], rtol=1e-13)
        np.testing.assert_allclose(orc.Id_like(g[f"Id_like_{i}_in"]), g[f"Id_like_{i}_desired"])
    np.testing.assert_allclose(orc.tf_super(g["tf_super_0_in"]), g["tf_super_0_desired"], rtol=1e-7)


def test_expm_closed_form_pauli():
    """exp(i theta n.sigma) = cos(theta) I + i sin(theta) n.sigma  (test/conftest.py:41-60,
    test/test_exp.py:10-32)."""
    rng = np.random.default_rng(0)
    sx = np.array([[0, 1], [1, 0]], dtype=complex)
    sy = np.array([[0, -1j], [1j, 0]])
    sz = np.array([[1, 0], [0, -1]], dtype=complex)
    for _ in range(20):
        theta = 2 * np.pi * rng.random()
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        ns = n[0] * sx + n[1] * sy + n[2] * sz
        want = np.cos(theta) * np.eye(2) + 1j * np.sin(theta) * ns
        assert rel_fro(orc.expm_tf(1j * theta * ns), want) < 1e-14


@pytest.mark.parametrize("d", [3, 9, 27])
@pytest.mark.parametrize("scale", [1e-3, 0.1, 0.6, 1.5, 3.0, 7.0, 20.0, 50.0])
def test_expm_tf_vs_scipy(d, scale):
    """Cross-check of the restated tf.linalg.expm against scipy over every Pade branch,
    including the squaring regime where TF under-scales (s = floor, not ceil)."""
    rng = np.random.default_rng(d)
    h = rng.normal(size=(4, d, d)) + 1j * rng.normal(size=(4, d, d))
    h = h + np.conj(np.swapaxes(h, -1, -2))
    a = -1j * h
    a *= (scale / np.abs(a).sum(axis=-2).max(axis=-1))[:, None, None]
    got = orc.expm_tf(a)
    # in the squaring regime TF scales by floor(log2(norm/theta13)), i.e. up to 2x less than
    # Higham's ceil: its own truncation error (Pade-13 applied at up to 2*theta13) reaches ~1e-10 there
    tol = 2e-13 if scale < orc.THETA13 else 1e-8
    for i in range(4):
        assert rel_fro(got[i], scipy.linalg.expm(a[i])) < tol


def test_tree_product_equals_sequential_for_all_lengths():
    rng = np.random.default_rng(1)
    for n in [1, 2, 3, 5, 8, 13, 50, 97]:
        x = rng.normal(size=(n, 3, 3)) + 1j * rng.normal(size=(n, 3, 3))
        t = orc.tf_matmul_n(x, orc.compute_folding_stack(n))
        s = orc.tf_matmul_left(x)
        assert rel_fro(t, s) < 1e-12


def test_evaluate_sequences():
    rng = np.random.default_rng(2)
    gates = {k: rng.normal(size=(3, 3)) + 1j * rng.normal(size=(3, 3)) for k in "abc"}
    out = orc.evaluate_sequences(gates, [["a", "b", "c"], [], ["c"]])
    assert rel_fro(out[0], gates["c"] @ gates["b"] @ gates["a"]) < 1e-14
    np.testing.assert_array_equal(out[1], np.eye(3))
    np.testing.assert_array_equal(out[2], gates["c"])


def test_tunable_coupler_d27_partial_propagators(golden_tunable_coupler):
    """test/test_tunable_coupler.py:399-409 of the reference (d = 27, 10 000 slices, every 50th stored):
    the oracle's TF-style expm on the rebuilt model reproduces the pickled slice propagators."""
    g = golden_tunable_coupler
    dt = g["tc_ts"][1] - g["tc_ts"][0]
    idx = g["dUs_index"]
    got = orc.tf_propagation_vectorized(g["h0"], g["hk_tc"][None], g["tc_signal"][None, idx], dt)
    assert rel_fro(got, g["dUs"]) < 1e-12
    assert abs(g["tc_ts"][0] - 0.5 * dt) < 1e-20 and len(g["tc_signal"]) == 10000


def test_rebuilt_tunable_coupler_model():
    """Dressed three-body model: Hermitian drift, diagonal in its own eigenbasis, coupler 0-1 and 1-2
    transition frequencies as pickled by the reference (checked when the fixture is made, repeated here
    on the stored numbers 45.100 and 43.785 Grad/s)."""
    from oracle import c3_model_oracle as mo
    m = mo.tunable_coupler_model()
    h0 = m["h0"]
    assert np.abs(h0 - np.diag(np.diag(h0))).max() < 1e-3 * np.abs(h0).max() * 1e-6
    e = m["eigenframe"]
    assert abs(abs(abs(e[0]) - abs(e[9])) - 45100118139.44866) / 45100118139.44866 < 1e-12
    assert abs(abs(abs(e[9]) - abs(e[18])) - 43784918348.34318) / 43784918348.34318 < 1e-12


def _tc_level_sweep(golden, dress):
    """Replay test/test_tunable_coupler.py:315-383 with ``dress(drift, ordered) -> eigenframe``."""
    from oracle import c3_model_oracle as mo
    e0 = dress(mo.tunable_coupler_drift(phi=0.0), True)
    order = np.argsort(np.abs(e0) / 2 / np.pi / 1e9)
    prod, ordd, dres, fallback = [], [], [], []
    for r in golden["flux_ratio"]:
        phi = r * 10.0
        drift = mo.tunable_coupler_drift(phi)
        prod.append(dress(mo.tunable_coupler_drift(phi, 0.0, 0.0), True)[order] / 2 / np.pi / 1e9)
        ordd.append(dress(drift, True)[order] / 2 / np.pi / 1e9)
        dres.append(dress(drift, False) / 2 / np.pi / 1e9)
        v = np.linalg.eigh(drift)[1]
        fallback.append((np.abs(v) ** 2).max(axis=0).min() <= 0.5)
    return np.array(prod), np.array(ordd), np.array(dres), np.array(fallback)


def test_dressing_oracle_against_pickled_energy_levels(golden_tc_levels):
    """f-4 oracle (eigh + reorder_frame, c3/model.py:453-502) against the reference's pickled level sweeps (d = 27,
    101 flux points).  Where every eigenvector has a component above 0.5 (70 points) all three tables are reproduced to
    1e-11 GHz; of the 31 "overly dressed" points 18 agree as well and at 13 avoided crossings one or two levels swap:
    the pickle predates the current greedy fallback, and the reference's own test only asks for |diff| < 1 GHz
    (test_tunable_coupler.py:376-381)."""
    from oracle import c3_model_oracle as mo
    g = golden_tc_levels
    prod, ordd, dres, fallback = _tc_level_sweep(g, lambda h, o: mo.dressing_transform(h, ordered=o)[0])
    assert np.abs(prod - g["product_basis"]).max() < 1e-12
    assert np.abs(dres - g["dressed_basis"]).max() < 1e-11
    assert np.abs(ordd[~fallback] - g["ordered_basis"][~fallback]).max() < 1e-11
    differs = (np.abs(ordd - g["ordered_basis"]) > 1e-9).sum(axis=1)
    assert fallback.sum() == 31 and (differs > 0).sum() == 13 and not np.any(differs[~fallback]) and differs.max() <= 2
    assert np.abs(ordd - g["ordered_basis"]).max() < 1.0
