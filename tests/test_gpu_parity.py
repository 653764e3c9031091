"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and against the
reference's golden vectors.  Tolerance is the one BASELINE.json states:
||U_gpu - U_ref||_F / ||U_ref||_F < 1e-10 (complex128)."""
import numpy as np
import pytest
import torch

from conftest import rel_fro
from oracle import c3_oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from c3_b200 import engine
    return engine


def _rand_model(rng, d, K, scale, hermitian=True):
    def herm():
        h = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
        return h + h.conj().T if hermitian else h
    h0 = herm()
    hks = np.stack([herm() for _ in range(K)]) if K else np.zeros((0, d, d), complex)
    h0 *= scale / np.abs(h0).sum(axis=0).max()
    for k in range(K):
        hks[k] *= 0.2 * scale / np.abs(hks[k]).sum(axis=0).max()
    return h0, hks


# ---------------------------------------------------------------------------------------------
# golden vectors of the reference
# ---------------------------------------------------------------------------------------------

def test_closed_two_qubit_golden(eng, golden_two_qubit):
    """test/test_two_qubits.py:46-62 of the reference."""
    g = golden_two_qubit
    dt = g["ts"][1] - g["ts"][0]
    U, dUs = eng.pwc_closed(g["hdrift"], g["hks"], g["signals"][None], dt, return_dUs=True)
    assert rel_fro(U[0].cpu().numpy(), g["propagator"]) < TOL
    want = orc.tf_batch_propagate(g["hdrift"], g["hks"], g["signals"], dt, 700)
    assert rel_fro(dUs[0].cpu().numpy(), want) < TOL


def test_lindblad_two_qubit_golden(eng, golden_two_qubit):
    """test/test_two_qubits.py:193-213 of the reference (16x16 superoperator)."""
    g = golden_two_qubit
    dt = g["ts"][1] - g["ts"][0]
    U = eng.pwc_lindblad(g["hdrift"], g["hks"], g["col_ops"], g["signals"][None], dt)
    assert rel_fro(U[0].cpu().numpy(), g["lindblad_propagator"]) < TOL


@pytest.mark.parametrize("q", ["q1", "q2"])
def test_transmon_expanded_golden(eng, golden_transmon, q):
    """test/test_transmon_expanded.py:252-283: H-list mode, excitation cut 24 -> 14, blow-up."""
    from c3_b200 import propagation as prop
    g = golden_transmon
    cutter = orc.make_ex_cutter(g["dims"], 4)
    hs = np.stack([orc.cut_excitations(cutter, h) for h in g[f"hamiltonians_{q}"]])
    ts = g[f"ts_{q}"][1:]
    dt = ts[1] - ts[0]
    U, dUs = eng.pwc_closed_hlist(hs[None], dt, return_dUs=True)
    dUs_big = prop.blowup_excitations(cutter, dUs[0]).cpu().numpy()
    U_big = prop.blowup_excitations(cutter, U[0]).cpu().numpy()
    assert rel_fro(dUs_big, g[f"partial_propagators_{q}"]) < TOL
    assert rel_fro(U_big, g[f"propagators_{q}"]) < TOL


def test_superoperator_helpers_golden(eng, golden_tf_utils):
    """test/test_tf_utils.py:81-111 of the reference."""
    from c3_b200 import tf_utils as tu
    g = golden_tf_utils
    for i in (0, 1):
        np.testing.assert_allclose(tu.tf_kron(g[f"tf_kron_{i}_inA"], g[f"tf_kron_{i}_inB"]).cpu().numpy(),
                                   g[f"tf_kron_{i}_desired"], rtol=1e-13)
        np.testing.assert_allclose(tu.tf_spre(g[f"tf_spre_{i}_in"]).cpu().numpy(), g[f"tf_spre_{i}_desired"], rtol=1e-13)
        np.testing.assert_allclose(tu.tf_spost(g[f"tf_spost_{i}_in"]).cpu().numpy(), g[f"tf_spost_{i}_desired"], rtol=1e-13)
        np.testing.assert_allclose(tu.Id_like(g[f"Id_like_{i}_in"]).cpu().numpy(), g[f"Id_like_{i}_desired"])
    np.testing.assert_allclose(tu.tf_super(g["tf_super_0_in"]).cpu().numpy(), g["tf_super_0_desired"], rtol=1e-7)


# ---------------------------------------------------------------------------------------------
# oracle parity on seeded inputs
# ---------------------------------------------------------------------------------------------

def test_headline_shape_d9(eng):
    """BASELINE config 2 shape (d=9, K=2, N=1000) at a batch the oracle finishes in seconds."""
    from c3_b200 import synth
    m = synth.two_transmon()
    sig = synth.controls(m, 6, 1000)
    U, dUs = eng.pwc_closed(m.h0, m.hks, sig, 1e-11, return_dUs=True)
    wantU, want_dUs = orc.propagate_batch(m.h0, m.hks, sig, 1e-11, return_dUs=True)
    for b in range(6):
        assert rel_fro(U[b].cpu().numpy(), wantU[b]) < TOL
    assert rel_fro(dUs.cpu().numpy(), want_dUs) < TOL
    # without the dUs store the result must be the same bits
    U2 = eng.pwc_closed(m.h0, m.hks, sig, 1e-11)
    assert torch.equal(U, U2)


@pytest.mark.parametrize("N", [1, 2, 3, 7, 50, 800])
def test_config1_single_qubit(eng, N):
    """BASELINE config 1: one 3-level qubit, B=1, N=50 (and the 800 the hjson actually yields),
    plus ragged tiny N."""
    from c3_b200 import synth
    m = synth.one_qubit()
    sig = synth.controls(m, 1, N)
    U, dUs = eng.pwc_closed(m.h0, m.hks, sig, 1e-11, return_dUs=True)
    wantU, want_dUs = orc.propagate_batch(m.h0, m.hks, sig, 1e-11, return_dUs=True)
    assert rel_fro(U.cpu().numpy(), wantU) < TOL
    assert rel_fro(dUs.cpu().numpy(), want_dUs) < TOL


@pytest.mark.parametrize("force_cta", [0, 1])
@pytest.mark.parametrize("d", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 16, 20, 27, 33])
def test_all_dimensions_and_paths(eng, d, force_cta):
    """Every register-kernel instantiation (incl. zero-padded d=7, 11), the shared-memory CTA
    kernel and the global-workspace CTA kernel (d=33), on the same seeded inputs."""
    rng = np.random.default_rng(100 + d)
    K, B, N = 2, 3, 37
    h0, hks = _rand_model(rng, d, K, 1.2)
    sig = rng.uniform(-1, 1, size=(B, K, N))
    eng.set_tuning("force_cta", force_cta)
    try:
        U, dUs = eng.pwc_closed(h0, hks, sig, 1.0, return_dUs=True)
    finally:
        eng.set_tuning("force_cta", 0)
    wantU, want_dUs = orc.propagate_batch(h0, hks, sig, 1.0, return_dUs=True)
    assert rel_fro(dUs.cpu().numpy(), want_dUs) < TOL
    assert rel_fro(U.cpu().numpy(), wantU) < TOL


@pytest.mark.parametrize("force_cta", [0, 1])
@pytest.mark.parametrize("scale", [1e-3, 0.1, 0.5, 1.3, 1.45, 2.5, 4.0])
def test_pade_orders_and_squarings(eng, scale, force_cta):
    """Every Pade order / squaring count.  Norms stay below theta_13 = 5.37: above it
    tf.linalg.expm under-scales and parity would mean matching TensorFlow's own truncation
    error; that regime is covered against scipy below."""
    rng = np.random.default_rng(7)
    d, K, B, N = 9, 2, 2, 21
    h0, hks = _rand_model(rng, d, K, scale)
    sig = rng.uniform(-1, 1, size=(B, K, N))
    eng.set_tuning("force_cta", force_cta)
    try:
        U, dUs = eng.pwc_closed(h0, hks, sig, 1.0, return_dUs=True)
    finally:
        eng.set_tuning("force_cta", 0)
    wantU, want_dUs = orc.propagate_batch(h0, hks, sig, 1.0, return_dUs=True)
    assert rel_fro(dUs.cpu().numpy(), want_dUs) < TOL
    assert rel_fro(U.cpu().numpy(), wantU) < TOL


@pytest.mark.parametrize("force_cta", [0, 1])
@pytest.mark.parametrize("scale", [8.0, 30.0, 200.0])
def test_large_norm_against_scipy(eng, scale, force_cta):
    import scipy.linalg
    rng = np.random.default_rng(8)
    d, K, B, N = 6, 1, 1, 5
    h0, hks = _rand_model(rng, d, K, scale)
    sig = rng.uniform(-1, 1, size=(B, K, N))
    eng.set_tuning("force_cta", force_cta)
    try:
        _, dUs = eng.pwc_closed(h0, hks, sig, 1.0, return_dUs=True)
    finally:
        eng.set_tuning("force_cta", 0)
    for n in range(N):
        want = scipy.linalg.expm(-1j * (h0 + sig[0, 0, n] * hks[0]))
        assert rel_fro(dUs[0, n].cpu().numpy(), want) < 1e-11


def test_ragged_lengths_and_segments(eng):
    """N not divisible by the lane-group count / segment length, with segmentation forced."""
    rng = np.random.default_rng(3)
    d, K = 9, 2
    h0, hks = _rand_model(rng, d, K, 1.0)
    for N in (5, 31, 64, 101):
        sig = rng.uniform(-1, 1, size=(2, K, N))
        want = orc.propagate_batch(h0, hks, sig, 1.0)
        for target in (1, 32768, 10 ** 7):
            eng.set_tuning("target_units", target)
            eng.set_tuning("min_chunk", 2)
            try:
                U = eng.pwc_closed(h0, hks, sig, 1.0)
            finally:
                eng.set_tuning("target_units", 0)
                eng.set_tuning("min_chunk", 8)
            assert rel_fro(U.cpu().numpy(), want) < TOL


def test_lindblad_small_dims(eng):
    """Lindblad with d=2 (D=4, register kernel on a NON-Hermitian generator), d=3 (D=9) and
    d=5 (D=25, CTA kernel)."""
    for d in (2, 3, 5):
        rng = np.random.default_rng(20 + d)
        K, B, N = 2, 2, 25
        h0, hks = _rand_model(rng, d, K, 0.8)
        col = 0.3 * (rng.normal(size=(2, d, d)) + 1j * rng.normal(size=(2, d, d)))
        sig = rng.uniform(-1, 1, size=(B, K, N))
        U, dUs = eng.pwc_lindblad(h0, hks, col, sig, 0.7, return_dUs=True)
        for b in range(B):
            want = orc.tf_batch_propagate(h0, hks, sig[b], 0.7, N, col_ops=col, lindbladian=True)
            assert rel_fro(dUs[b].cpu().numpy(), want) < TOL
            assert rel_fro(U[b].cpu().numpy(), orc.tf_matmul_n(want, orc.compute_folding_stack(N))) < TOL


@pytest.mark.parametrize("cta_variant", [0, 1])
@pytest.mark.parametrize("d", [13, 16, 24, 27, 33, 40])
def test_cta_kernels_pade_and_taylor(eng, d, cta_variant):
    """Both CTA kernels on the same inputs: 0 = Higham Pade + pivoted Gauss-Jordan (the literal
    restatement of tf.linalg.expm), 1 = degree-18 Taylor on DMMA tiles with trace shift (default)."""
    rng = np.random.default_rng(200 + d)
    K, B, N = 2, 2, 9
    h0, hks = _rand_model(rng, d, K, 2.2)
    h0 = h0 + 0.7 * np.eye(d)          # non-zero trace: exercises the shift / phase re-application
    sig = rng.uniform(-1, 1, size=(B, K, N))
    eng.set_tuning("cta_variant", cta_variant)
    try:
        U, dUs = eng.pwc_closed(h0, hks, sig, 1.0, return_dUs=True)
    finally:
        eng.set_tuning("cta_variant", 1)
    wantU, want_dUs = orc.propagate_batch(h0, hks, sig, 1.0, return_dUs=True)
    assert rel_fro(dUs.cpu().numpy(), want_dUs) < TOL
    assert rel_fro(U.cpu().numpy(), wantU) < TOL


D9_VARIANTS = [0, 1, 2, 3]


def _default_d9_variant():
    import os
    return int(os.environ.get("C3B_D9_VARIANT", 1))


@pytest.mark.parametrize("variant", D9_VARIANTS)
def test_d9_kernel_variants_headline_model(eng, variant):
    """The d = 9 kernels on the headline model: generic 3x3-block kernel (0), own-block products from the annealed
    shared-memory layout (1), shuffle-exchange kernel (2)."""
    from c3_b200 import synth
    m = synth.two_transmon()
    sig = synth.controls(m, 3, 203)
    eng.set_tuning("d9_variant", variant)
    try:
        U, dUs = eng.pwc_closed(m.h0, m.hks, sig, 1e-11, return_dUs=True)
    finally:
        eng.set_tuning("d9_variant", _default_d9_variant())
    wantU, want_dUs = orc.propagate_batch(m.h0, m.hks, sig, 1e-11, return_dUs=True)
    assert rel_fro(dUs.cpu().numpy(), want_dUs) < TOL
    assert rel_fro(U.cpu().numpy(), wantU) < TOL


@pytest.mark.parametrize("variant", D9_VARIANTS)
@pytest.mark.parametrize("scale,N,B,d", [(6.0, 37, 2, 9), (30.0, 20, 3, 9), (0.5, 1, 2, 9), (2.0, 1000, 5, 9), (3.0, 29, 4, 7)])
def test_d9_kernel_variants_random_models(eng, variant, scale, N, B, d):
    """d = 9 kernels (and zero-padded d = 7) on random complex Hermitian models: squarings (norm up to ~30), a single
    slice, ragged segments (B * N forces several segments per batch element), the H-list entry and the partial propagators."""
    rng = np.random.default_rng(int(scale * 10) + N)
    K = 2
    h0, hks = _rand_model(rng, d, K, scale)
    sig = rng.uniform(-1, 1, size=(B, K, N))
    eng.set_tuning("d9_variant", variant)
    try:
        U, dUs = eng.pwc_closed(h0, hks, sig, 1.0, return_dUs=True)
        Hs = h0[None, None] + np.einsum("bkn,kij->bnij", sig, hks)
        U2 = eng.pwc_closed_hlist(Hs, 1.0)
    finally:
        eng.set_tuning("d9_variant", _default_d9_variant())
    wantU, want_dUs = orc.propagate_batch(h0, hks, sig, 1.0, return_dUs=True)
    assert rel_fro(dUs.cpu().numpy(), want_dUs) < TOL
    assert rel_fro(U.cpu().numpy(), wantU) < TOL
    assert rel_fro(U2.cpu().numpy(), wantU) < TOL


@pytest.mark.parametrize("norm_bound", [0, 1])
@pytest.mark.parametrize("d,scale", [(14, 0.8), (27, 2.5), (27, 20.0), (33, 6.0)])
def test_cta_kernel_scaling_from_row_sum_bound(eng, d, scale, norm_bound):
    """DMMA CTA kernel: squarings chosen from the row-sum bound of the generators (default) or from the exact inf-norm of
    every assembled slice; shared and per-sample models."""
    rng = np.random.default_rng(d + int(scale * 10))
    K, B, N = 2, 3, 7
    h0, hks = _rand_model(rng, d, K, scale)
    sig = rng.uniform(-1, 1, size=(B, K, N))
    eng.set_tuning("norm_bound", norm_bound)
    try:
        U, dUs = eng.pwc_closed(h0, hks, sig, 1.0, return_dUs=True)
        Ub = eng.pwc_closed(np.stack([h0 * (1 + 0.1 * b) for b in range(B)]), np.stack([hks] * B), sig, 1.0)
    finally:
        eng.set_tuning("norm_bound", 1)
    wantU, want_dUs = orc.propagate_batch(h0, hks, sig, 1.0, return_dUs=True)
    assert rel_fro(dUs.cpu().numpy(), want_dUs) < TOL
    assert rel_fro(U.cpu().numpy(), wantU) < TOL
    for b in range(B):
        want = orc.propagate_batch(h0 * (1 + 0.1 * b), hks, sig[b:b + 1], 1.0)[0]
        assert rel_fro(Ub[b].cpu().numpy(), want) < TOL


@pytest.mark.parametrize("gemm_big", [0, 1, 2, 3, 4, 5])
def test_lindblad_config3_shape(eng, gemm_big):
    """BASELINE config 3 shape (two 3-level transmons, D=81) on a small batch / few slices, with every instantiation of the
    global-workspace DMMA kernel: macro-tile shapes / CTA sizes (0 default, 1..3) and the cp.async-staged block-row product
    (4: 22 warps, 5: 12 warps)."""
    from c3_b200 import synth
    m = synth.two_transmon()
    B, N = 2, 12
    sig = synth.controls(m, B, 1000)[:, :, 494:494 + N].copy()
    eng.set_tuning("gemm_big", gemm_big)
    try:
        U, dUs = eng.pwc_lindblad(m.h0, m.hks, m.col_ops, sig, 1e-11, return_dUs=True)
    finally:
        eng.set_tuning("gemm_big", 0)
    for b in range(B):
        want = orc.tf_batch_propagate(m.h0, m.hks, sig[b], 1e-11, N, col_ops=m.col_ops, lindbladian=True)
        assert rel_fro(dUs[b].cpu().numpy(), want) < TOL
        assert rel_fro(U[b].cpu().numpy(), orc.tf_matmul_n(want, orc.compute_folding_stack(N))) < TOL


def test_config5_shape_d27(eng):
    """BASELINE config 5 shape (tunable coupler d=27, K=3) on a small batch."""
    from c3_b200 import synth
    m = synth.tunable_coupler()
    sig = synth.controls(m, 3, 64)
    U, dUs = eng.pwc_closed(m.h0, m.hks, sig, 1e-11, return_dUs=True)
    wantU, want_dUs = orc.propagate_batch(m.h0, m.hks, sig, 1e-11, return_dUs=True)
    assert rel_fro(dUs.cpu().numpy(), want_dUs) < TOL
    assert rel_fro(U.cpu().numpy(), wantU) < TOL


def test_tunable_coupler_golden_d27(eng, golden_tunable_coupler):
    """test/test_tunable_coupler.py:399-409 of the reference: d = 27 slice propagators of the CPHASE flux
    pulse (10 000 slices, stored every 100th here) through the DMMA CTA kernel, and the full product
    against the oracle."""
    g = golden_tunable_coupler
    dt = float(g["tc_ts"][1] - g["tc_ts"][0])
    U, dUs = eng.pwc_closed(g["h0"], g["hk_tc"][None], g["tc_signal"][None, None, :], dt, return_dUs=True)
    got = dUs[0].cpu().numpy()
    assert rel_fro(got[g["dUs_index"]], g["dUs"]) < TOL
    # three control lines as in the reference (the two qubit drives carry no_drive = zeros)
    from oracle import c3_model_oracle as mo
    m = mo.tunable_coupler_model()
    hks = np.stack([m["hk_q1"], m["hk_q2"], g["hk_tc"]])
    sig3 = np.zeros((1, 3, 10000))
    sig3[0, 2] = g["tc_signal"]
    U3 = eng.pwc_closed(g["h0"], hks, sig3, dt)
    assert rel_fro(U3.cpu().numpy(), U.cpu().numpy()) < 1e-12
    want = np.eye(27, dtype=complex)
    for n in range(10000):
        want = got[n] @ want
    assert rel_fro(U[0].cpu().numpy(), want) < TOL
    assert rel_fro(U[0].cpu().numpy().conj().T @ U[0].cpu().numpy(), np.eye(27)) < 1e-10


def test_batched_model(eng):
    """Per-sample models h0[B,d,d], hks[B,K,d,d] (optimiser samples that change the model)."""
    rng = np.random.default_rng(11)
    d, K, B, N = 9, 2, 4, 19
    models = [_rand_model(rng, d, K, 1.0) for _ in range(B)]
    h0 = np.stack([m[0] for m in models])
    hks = np.stack([m[1] for m in models])
    sig = rng.uniform(-1, 1, size=(B, K, N))
    U = eng.pwc_closed(h0, hks, sig, 1.0)
    for b in range(B):
        want = orc.propagate_batch(h0[b], hks[b], sig[b:b + 1], 1.0)[0]
        assert rel_fro(U[b].cpu().numpy(), want) < TOL


@pytest.mark.parametrize("D", [2, 9, 16, 27, 70])
@pytest.mark.parametrize("M", [1, 2, 9, 64, 301])
def test_ordered_product(eng, D, M):
    """tf_matmul_left / tf_matmul_n semantics (c3/utils/tf_utils.py:120-193)."""
    from c3_b200 import tf_utils as tu
    rng = np.random.default_rng(D * 1000 + M)
    x = (rng.normal(size=(M, D, D)) + 1j * rng.normal(size=(M, D, D))) / np.sqrt(2 * D)
    want = orc.tf_matmul_n(x, orc.compute_folding_stack(M))
    got = tu.tf_matmul_n(x, tu.compute_folding_stack(M)).cpu().numpy()
    assert rel_fro(got, want) < 1e-11
    assert rel_fro(tu.tf_matmul_left(x).cpu().numpy(), orc.tf_matmul_left(x)) < 1e-11
    if M > 1:
        assert rel_fro(tu.tf_matmul_right(x).cpu().numpy(), orc.tf_matmul_right(x)) < 1e-11


@pytest.mark.parametrize("d,seq_variant", [(9, 1), (9, 0), (3, 1), (4, 1), (7, 1), (12, 1), (27, 1)])
def test_evaluate_sequences(eng, d, seq_variant):
    """c3/libraries/propagation.py:588-627 incl. the empty-sequence identity: the lane-group kernel (small d, gate table
    in shared memory; d = 7 zero-padded to 8) and the CTA-per-sequence kernel (seq_variant 0, and any d > 12)."""
    from c3_b200 import propagation as prop
    rng = np.random.default_rng(5 + d)
    names = ["rx90p[0]", "ry90p[0]", "rx90m[0]", "ry90m[0]", "id[0]"]
    gates = {}
    for n in names:
        q, _ = np.linalg.qr(rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)))
        gates[n] = q
    seqs = [[], ["id[0]"], ["rx90p[0]", "ry90p[0]"]]
    for _ in range(47):
        L = int(rng.integers(1, 60))
        seqs.append([names[i] for i in rng.integers(0, len(names), size=L)])
    eng.set_tuning("seq_variant", seq_variant)
    try:
        got = prop.evaluate_sequences({k: torch.as_tensor(v) for k, v in gates.items()}, seqs)
    finally:
        eng.set_tuning("seq_variant", 1)
    want = orc.evaluate_sequences(gates, seqs)
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert rel_fro(a.cpu().numpy(), b) < 1e-11
    np.testing.assert_array_equal(got[0].cpu().numpy(), np.eye(d))


def test_full_size_properties(eng):
    """BASELINE config 2 at full size (d=9, N=1000, B=256): size-independent properties.
    U must be unitary; splitting the time axis must compose (U = U_second_half U_first_half);
    and a sample of batch rows is checked against the oracle."""
    from c3_b200 import synth
    m = synth.two_transmon()
    B, N = 256, 1000
    sig = synth.controls_fast(m, B, N)
    U = eng.pwc_closed(m.h0, m.hks, sig, 1e-11)
    eye = torch.eye(9, dtype=torch.complex128, device=U.device)
    err = (U.conj().transpose(-1, -2) @ U - eye).abs().amax().item()
    assert err < 1e-11
    U1 = eng.pwc_closed(m.h0, m.hks, sig[:, :, :500].copy(), 1e-11)
    U2 = eng.pwc_closed(m.h0, m.hks, sig[:, :, 500:].copy(), 1e-11)
    comp = torch.matmul(U2, U1)
    assert rel_fro(comp.cpu().numpy(), U.cpu().numpy()) < 1e-11
    for b in (0, 17, 255):
        want = orc.propagate_batch(m.h0, m.hks, sig[b:b + 1], 1e-11)[0]
        assert rel_fro(U[b].cpu().numpy(), want) < TOL


def test_host_pipelined_path_matches_device_path(eng):
    """propagation.pwc_batch with HOST signals (chunked H2D/compute pipeline, the e2e path of
    bench.py) gives the same bits as the device-resident call."""
    from c3_b200 import synth, propagation as prop
    m = synth.two_transmon()
    B, N = 2500, 120                     # not a multiple of the 1024-row chunk
    sig = synth.controls_fast(m, B, N)
    host = torch.as_tensor(sig).pin_memory()
    U_host = prop.pwc_batch(m.h0, m.hks, host, 1e-11)
    U_np = prop.pwc_batch(m.h0, m.hks, sig, 1e-11)          # pageable numpy input
    U_dev = eng.pwc_closed(m.h0, m.hks, torch.as_tensor(sig).cuda(), 1e-11)
    torch.cuda.synchronize()
    assert U_host.shape == (B, 9, 9)
    assert rel_fro(U_host.cpu().numpy(), U_dev.cpu().numpy()) < 1e-13
    assert rel_fro(U_np.cpu().numpy(), U_dev.cpu().numpy()) < 1e-13
    want = orc.propagate_batch(m.h0, m.hks, sig[[0, 1023, 1024, 2499]], 1e-11)
    assert rel_fro(U_host[[0, 1023, 1024, 2499]].cpu().numpy(), want) < TOL


@pytest.mark.parametrize("d,K,N,chunk,first", [(9, 2, 40, 3, 1), (9, 2, 40, 64, 256), (6, 2, 40, 5, 2), (27, 2, 40, 4, 1),
                                               (9, 1, 37, 1, 1), (9, 2, 37, 2, 1), (9, 1, 37, 3, 2), (9, 3, 13, 1, 3), (7, 1, 5, 2, 1)])
def test_gated_single_launch_host_path(eng, d, K, N, chunk, first):
    """Host-resident control fields through the gated single launch (c3b_pwc_closed_gated: d = 9) and through the
    chunked two-stream fallback (d = 6, 27): many small chunks, repeated calls on reused buffers, bit-identical to the
    device-resident call.  K * N not a multiple of 4 makes batch rows share 32-byte sectors with their neighbours: the
    rows in flight must not be read through the non-coherent path."""
    rng = np.random.default_rng(d + N)
    B = 23
    h0, hks = _rand_model(rng, d, K, 1.5)
    for rep in range(4):
        sig = rng.uniform(-1, 1, size=(B, K, N))
        host = torch.as_tensor(sig).pin_memory()
        U_host = eng.pwc_closed_from_host(h0, hks, host, 1.0, chunk=chunk, first_chunk=first)
        U_dev = eng.pwc_closed(h0, hks, torch.as_tensor(sig).cuda(), 1.0)
        torch.cuda.synchronize()
        eng.check_gated_launches()
        assert not torch.isnan(U_host.real).any()
        assert rel_fro(U_host.cpu().numpy(), U_dev.cpu().numpy()) < 1e-13
    want = orc.propagate_batch(h0, hks, sig, 1.0)
    assert rel_fro(U_host.cpu().numpy(), want) < TOL


def test_gated_launch_reports_rows_that_never_arrive(eng):
    """A gate that stops at row 3 of 8: the kernel gives up after its patience runs out (no hang), raises gate[1], leaves
    the rows it never saw as NaN, and the host turns the flag into an error at its next check."""
    from c3_b200 import _lib, synth
    lib = _lib.load()
    if not lib.c3b_pwc_gated_supported(9):
        pytest.skip("the selected d = 9 kernel has no gated entry")
    m = synth.two_transmon()
    B, K, N = 8, 2, 16
    dev = eng.default_device()
    sig = torch.as_tensor(synth.controls(m, B, N)).to(dev)
    h0 = torch.as_tensor(m.h0).to(dev)
    hks = torch.as_tensor(m.hks).to(dev)
    U = torch.full((B, 9, 9), float("nan"), dtype=torch.complex128, device=dev)
    gate = torch.tensor([3, 0], dtype=torch.int32, device=dev)
    ws = torch.empty(lib.c3b_pwc_workspace_bytes(B, K, N, 9, 0, 0), dtype=torch.uint8, device=dev)
    rc = lib.c3b_pwc_closed_gated(h0.data_ptr(), hks.data_ptr(), sig.data_ptr(), 1e-11, B, K, N, 9, U.data_ptr(),
                                  gate.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    eng._watch_gate(gate, torch.cuda.current_stream())
    with pytest.raises(_lib.C3BError, match="gated launch timed out"):
        eng.check_gated_launches()
    assert int(gate[1].item()) == 1
    done = ~torch.isnan(U.real).any(dim=(1, 2))
    assert done[:3].all() and not done[3:].any()
    want = orc.propagate_batch(m.h0, m.hks, sig[:3].cpu().numpy(), 1e-11)
    assert rel_fro(U[:3].cpu().numpy(), want) < TOL


def test_error_reporting(eng):
    from c3_b200 import _lib
    lib = _lib.load()
    rc = lib.c3b_pwc_closed(None, None, None, 1.0, 1, 0, 1, 3, 0, None, None, None, 0, None)
    assert rc != 0
    assert lib.c3b_last_error().decode().startswith("C3:ERROR:")
    with pytest.raises(ValueError):
        eng.pwc_closed(np.eye(3), np.zeros((1, 4, 4)), np.zeros((1, 1, 5)), 1.0)


# ---------------------------------------------------------------------------------------------
# gradients
# ---------------------------------------------------------------------------------------------

@pytest.mark.parametrize("levels,N,B", [(2, 40, 3), (3, 60, 4)])
def test_gradient_matches_autograd_oracle(eng, levels, N, B):
    """dL/dsignals through the propagator vs torch-CPU autograd through matrix_exp (the role
    tf.GradientTape plays in the reference), for a unitary-overlap infidelity."""
    from c3_b200 import synth, propagation as prop
    from oracle import c3_grad_oracle as gorc
    m = synth.two_transmon(levels=levels)
    d = m.d
    rng = np.random.default_rng(levels)
    sig = synth.controls(m, B, N)
    q = np.stack([np.linalg.qr(rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)))[0] for _ in range(B)])
    L_ref, g_ref, U_ref = gorc.loss_and_grad(m.h0, m.hks, sig, 1e-11, q)

    s = torch.tensor(sig, device="cuda", requires_grad=True)
    U = prop.pwc_batch_autograd(m.h0, m.hks, s, 1e-11)
    T = torch.as_tensor(q, device="cuda")
    ov = torch.einsum("bij,bij->b", T.conj(), U)
    L = (1.0 - (ov.abs() ** 2) / d ** 2).sum()
    L.backward()
    assert rel_fro(U.detach().cpu().numpy(), U_ref) < TOL
    assert abs(float(L) - L_ref) < 1e-10
    g = s.grad.cpu().numpy()
    assert g.shape == g_ref.shape
    assert rel_fro(g, g_ref) < 1e-8


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("scale,d", [(1.0, 9), (6.0, 9), (40.0, 9), (3.0, 5), (1.0, 12), (1.0, 16)])
def test_gradient_variants_and_squarings(eng, variant, scale, d):
    """The gradient implementations (0: augmented exponential, 1: default -- the fused lockstep kernel at d = 9, the
    Frechet kernels elsewhere, 2: the stored-propagator Frechet kernels everywhere) against torch-CPU autograd, including
    slices that need 1 .. 6 squarings and every chunk width of the shared-memory product (d divisible by 3, by 2, by
    neither)."""
    from oracle import c3_grad_oracle as gorc
    rng = np.random.default_rng(int(scale) + d)
    K, B, N = 2, 3, 17
    h0, hks = _rand_model(rng, d, K, 0.9 * scale)
    sig = rng.uniform(-1, 1, size=(B, K, N))
    q = np.stack([np.linalg.qr(rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)))[0] for _ in range(B)])
    L_ref, g_ref, U_ref = gorc.loss_and_grad(h0, hks, sig, 1.0, q)
    Uc = torch.tensor(U_ref, requires_grad=True)
    T = torch.as_tensor(q)
    ((1.0 - (torch.einsum("bij,bij->b", T.conj(), Uc).abs() ** 2) / d ** 2).sum()).backward()
    eng.set_tuning("grad_variant", variant)
    try:
        U, g = eng.pwc_closed_grad(h0, hks, sig, 1.0, Uc.grad.numpy())
    finally:
        eng.set_tuning("grad_variant", 1)
    assert rel_fro(U.cpu().numpy(), U_ref) < TOL
    assert rel_fro(g.cpu().numpy(), g_ref) < 1e-8


@pytest.mark.parametrize("levels", [2, 3])
def test_lindblad_gradient_matches_autograd_oracle(eng, levels):
    """Open-system gradient: dL/dsignals through the Lindblad superoperator propagators (D = 4: one qubit pair would be
    16; here one/two-level systems of the synthetic chip) vs torch-CPU autograd through the same operator."""
    from c3_b200 import synth, propagation as prop
    from oracle import c3_grad_oracle as gorc
    m = synth.one_qubit(levels=3) if levels == 3 else synth.two_transmon(levels=2)     # D = 9 resp. 16
    D = m.d * m.d
    B, N = 3, 25
    rng = np.random.default_rng(levels)
    sig = synth.controls(m, B, N)
    cols = np.asarray(m.col_ops) * 3e3          # strong damping so that the dissipator matters over 25 slices
    T = rng.normal(size=(B, D, D)) + 1j * rng.normal(size=(B, D, D))
    s_ref = torch.tensor(sig, dtype=torch.float64, requires_grad=True)
    U_ref = gorc.propagate_lindblad_torch(m.h0, m.hks, cols, s_ref, 1e-11)
    L_ref = (torch.einsum("bij,bij->b", torch.as_tensor(T).conj(), U_ref).abs() ** 2).sum()
    L_ref.backward()
    want_U = orc.propagate_batch(m.h0, m.hks, sig, 1e-11, col_ops=cols, lindbladian=True)
    assert rel_fro(U_ref.detach().numpy(), want_U) < 1e-10           # the torch restatement agrees with the oracle
    s = torch.tensor(sig, device="cuda", requires_grad=True)
    U = prop.pwc_batch_autograd(m.h0, m.hks, s, 1e-11, col_ops=list(cols), lindbladian=True)
    L = (torch.einsum("bij,bij->b", torch.as_tensor(T, device="cuda").conj(), U).abs() ** 2).sum()
    L.backward()
    assert rel_fro(U.detach().cpu().numpy(), want_U) < TOL
    assert abs(float(L) - float(L_ref)) < 1e-9 * abs(float(L_ref))
    assert rel_fro(s.grad.cpu().numpy(), s_ref.grad.numpy()) < 1e-8


@pytest.mark.parametrize("scale,d,K", [(1.0, 20, 2), (1.0, 27, 3), (5.0, 27, 3), (25.0, 24, 1), (1.0, 33, 2), (2.0, 48, 1)])
def test_gradient_cta_path_closed(eng, scale, d, K):
    """Closed-system gradient for d > 16: CTA sweeps + Frechet derivative of the Taylor scheme on the DMMA product (shared
    memory matrices up to d = 32, per-CTA global workspace above), with 0 .. 5 squarings, against torch-CPU autograd."""
    from oracle import c3_grad_oracle as gorc
    rng = np.random.default_rng(int(scale) + d)
    B, N = 3, 9
    h0, hks = _rand_model(rng, d, K, 0.9 * scale)
    sig = rng.uniform(-1, 1, size=(B, K, N))
    q = np.stack([np.linalg.qr(rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)))[0] for _ in range(B)])
    L_ref, g_ref, U_ref = gorc.loss_and_grad(h0, hks, sig, 1.0, q)
    Uc = torch.tensor(U_ref, requires_grad=True)
    ((1.0 - (torch.einsum("bij,bij->b", torch.as_tensor(q).conj(), Uc).abs() ** 2) / d ** 2).sum()).backward()
    U, g = eng.pwc_closed_grad(h0, hks, sig, 1.0, Uc.grad.numpy())
    assert rel_fro(U.cpu().numpy(), U_ref) < TOL
    assert rel_fro(g.cpu().numpy(), g_ref) < 1e-8
    if d <= 32:            # the augmented-exponential cross-check agrees
        eng.set_tuning("grad_variant", 0)
        try:
            _, g0 = eng.pwc_closed_grad(h0, hks, sig, 1.0, Uc.grad.numpy())
        finally:
            eng.set_tuning("grad_variant", 1)
        assert rel_fro(g0.cpu().numpy(), g_ref) < 1e-8


@pytest.mark.parametrize("model,N", [("qutrit_pair_d5", 7), ("two_transmon", 6)])
def test_lindblad_gradient_cta_path(eng, model, N):
    """Open-system gradient above D = 16: D = 25 (d = 5, shared-memory CTA kernels) and the BASELINE config-3 shape D = 81
    (two 3-level transmons, global workspace) against torch-CPU autograd through the same superoperator."""
    from c3_b200 import synth, propagation as prop
    from oracle import c3_grad_oracle as gorc
    if model == "two_transmon":
        m = synth.two_transmon()
        h0, hks, cols = m.h0, m.hks, np.asarray(m.col_ops) * 3e3
        dt = 1e-11
        sig = synth.controls(m, 2, N)
    else:
        rng0 = np.random.default_rng(5)
        h0, hks = _rand_model(rng0, 5, 2, 1.2)
        cols = np.stack([0.3 * np.diag(np.sqrt(np.arange(1, 5)), k=1).astype(complex), 0.2 * np.diag(np.arange(5)).astype(complex)])
        dt = 1.0
        sig = rng0.uniform(-1, 1, size=(2, 2, N))
    D = h0.shape[0] ** 2
    B = sig.shape[0]
    rng = np.random.default_rng(D)
    T = rng.normal(size=(B, D, D)) + 1j * rng.normal(size=(B, D, D))
    s_ref = torch.tensor(sig, dtype=torch.float64, requires_grad=True)
    U_ref = gorc.propagate_lindblad_torch(h0, hks, cols, s_ref, dt)
    L_ref = (torch.einsum("bij,bij->b", torch.as_tensor(T).conj(), U_ref).abs() ** 2).sum()
    L_ref.backward()
    s = torch.tensor(sig, device="cuda", requires_grad=True)
    U = prop.pwc_batch_autograd(h0, hks, s, dt, col_ops=list(cols), lindbladian=True)
    L = (torch.einsum("bij,bij->b", torch.as_tensor(T, device="cuda").conj(), U).abs() ** 2).sum()
    L.backward()
    assert rel_fro(U.detach().cpu().numpy(), U_ref.detach().numpy()) < TOL
    assert abs(float(L.detach()) - float(L_ref.detach())) < 1e-9 * abs(float(L_ref.detach()))
    assert rel_fro(s.grad.cpu().numpy(), s_ref.grad.numpy()) < 1e-8


@pytest.mark.parametrize("d,K,B,N,scale", [(9, 2, 4, 131, 1.0), (9, 1, 1, 5, 1.0), (9, 3, 2, 50, 6.0), (8, 2, 3, 33, 1.0),
                                           (7, 2, 5, 64, 3.0), (9, 2, 1, 700, 0.5), (9, 5, 2, 40, 1.0)])
def test_gradient_fused_d9_kernel(eng, d, K, B, N, scale):
    """The fused d = 9 gradient kernel (grad_blk9.cuh: Y_n recurrence by unitarity, lockstep Frechet derivative, no stored
    propagators) against torch-CPU autograd and against the stored-propagator kernels: zero-padded d = 7, 8, ragged last
    chunks, a single slice chunk, several chunks per row, K = 1 .. 5, slices that need squarings."""
    from oracle import c3_grad_oracle as gorc
    rng = np.random.default_rng(d * 100 + K * 10 + N)
    h0, hks = _rand_model(rng, d, K, 0.9 * scale)
    sig = rng.uniform(-1, 1, size=(B, K, N))
    Ubar = rng.normal(size=(B, d, d)) + 1j * rng.normal(size=(B, d, d))
    U1, g1 = eng.pwc_closed_grad(h0, hks, sig, 1.0, Ubar)
    eng.set_tuning("grad_variant", 2)
    try:
        U2, g2 = eng.pwc_closed_grad(h0, hks, sig, 1.0, Ubar)
    finally:
        eng.set_tuning("grad_variant", 1)
    assert rel_fro(U1.cpu().numpy(), U2.cpu().numpy()) < 1e-12
    assert rel_fro(g1.cpu().numpy(), g2.cpu().numpy()) < 1e-9
    if N <= 131:
        q = np.stack([np.linalg.qr(rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)))[0] for _ in range(B)])
        L_ref, g_ref, U_ref = gorc.loss_and_grad(h0, hks, sig, 1.0, q)
        Uc = torch.tensor(U_ref, requires_grad=True)
        ((1.0 - (torch.einsum("bij,bij->b", torch.as_tensor(q).conj(), Uc).abs() ** 2) / d ** 2).sum()).backward()
        U, g = eng.pwc_closed_grad(h0, hks, sig, 1.0, Uc.grad.numpy())
        assert rel_fro(U.cpu().numpy(), U_ref) < TOL
        assert rel_fro(g.cpu().numpy(), g_ref) < 1e-8


@pytest.mark.parametrize("d,K,B,N,scale", [(27, 3, 3, 37, 1.0), (27, 3, 2, 9, 5.0), (17, 1, 2, 21, 1.0), (24, 2, 2, 30, 3.0), (32, 2, 1, 12, 1.0),
                                           (20, 2, 5, 3, 25.0), (29, 4, 2, 70, 1.0)])
def test_gradient_fused_cta_kernel(eng, d, K, B, N, scale):
    """The fused unitary-recurrence kernel on the DMMA product (grad_ucta.cuh, closed 16 < d <= 32: the swizzled DP = 32
    instances and the padded ones, several chunks per row, ragged last chunks, 1 .. 5 squarings) against the
    stored-propagator kernels and torch-CPU autograd."""
    from oracle import c3_grad_oracle as gorc
    rng = np.random.default_rng(d * 100 + K * 10 + N)
    h0, hks = _rand_model(rng, d, K, 0.9 * scale)
    sig = rng.uniform(-1, 1, size=(B, K, N))
    Ubar = rng.normal(size=(B, d, d)) + 1j * rng.normal(size=(B, d, d))
    U1, g1 = eng.pwc_closed_grad(h0, hks, sig, 1.0, Ubar)
    eng.set_tuning("grad_variant", 2)
    try:
        U2, g2 = eng.pwc_closed_grad(h0, hks, sig, 1.0, Ubar)
    finally:
        eng.set_tuning("grad_variant", 1)
    assert rel_fro(U1.cpu().numpy(), U2.cpu().numpy()) < 1e-12
    assert rel_fro(g1.cpu().numpy(), g2.cpu().numpy()) < 1e-9
    if N <= 40:
        q = np.stack([np.linalg.qr(rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)))[0] for _ in range(B)])
        L_ref, g_ref, U_ref = gorc.loss_and_grad(h0, hks, sig, 1.0, q)
        Uc = torch.tensor(U_ref, requires_grad=True)
        ((1.0 - (torch.einsum("bij,bij->b", torch.as_tensor(q).conj(), Uc).abs() ** 2) / d ** 2).sum()).backward()
        U, g = eng.pwc_closed_grad(h0, hks, sig, 1.0, Uc.grad.numpy())
        assert rel_fro(U.cpu().numpy(), U_ref) < TOL
        assert rel_fro(g.cpu().numpy(), g_ref) < 1e-8


@pytest.mark.parametrize("d,K,B,N", [(9, 2, 3, 61), (27, 3, 2, 29), (9, 2, 1, 7), (5, 2, 2, 11)])
def test_gradient_forward_saved_backward(eng, d, K, B, N):
    """Forward that keeps its chunk products + backward from the saved state (c3b_pwc_closed_fwd_saved / _bwd_saved, what the
    autograd node uses) against the one-call gradient; shapes without a fused kernel and non-Hermitian models return no state."""
    from c3_b200 import propagation as prop
    rng = np.random.default_rng(d + N)
    h0, hks = _rand_model(rng, d, K, 0.9)
    sig = rng.uniform(-1, 1, size=(B, K, N))
    Ubar = rng.normal(size=(B, d, d)) + 1j * rng.normal(size=(B, d, d))
    U_ref, g_ref = eng.pwc_closed_grad(h0, hks, sig, 1.0, Ubar)
    U, saved = eng.pwc_closed_saving(h0, hks, sig, 1.0)
    assert rel_fro(U.cpu().numpy(), U_ref.cpu().numpy()) < 1e-13
    if d == 5:
        assert saved is None
    else:
        assert saved is not None and saved.Q * saved.CL >= N
        g = eng.pwc_closed_grad_saved(sig, Ubar, saved)
        assert rel_fro(g.cpu().numpy(), g_ref.cpu().numpy()) < 1e-12
        g2 = eng.pwc_closed_grad_saved(sig, 2.0 * Ubar, saved)            # the state serves any number of cotangents
        assert rel_fro(g2.cpu().numpy(), 2.0 * g_ref.cpu().numpy()) < 1e-12
        hn, hkn = _rand_model(rng, d, K, 0.9, hermitian=False)
        assert eng.pwc_closed_saving(hn, hkn, sig, 1.0)[1] is None
        assert eng.pwc_closed_saving(torch.as_tensor(hn).cuda(), torch.as_tensor(hkn).cuda(), sig, 1.0)[1] is None   # found on the device
    # the autograd node takes whichever path exists
    s_t = torch.tensor(sig, device="cuda", requires_grad=True)
    Ua = prop.pwc_batch_autograd(h0, hks, s_t, 1.0)
    (torch.view_as_real(Ua) * torch.view_as_real(torch.as_tensor(Ubar, device="cuda"))).sum().backward()
    assert rel_fro(s_t.grad.cpu().numpy(), g_ref.cpu().numpy()) < 1e-11


def test_gradient_non_hermitian_falls_back(eng):
    """The fused kernel assumes unitary slice propagators; a non-Hermitian 'Hamiltonian' must be detected on the device and
    served by the stored-propagator kernels (same result as forcing them), and the caller's assertion 'grad_unitary' = 0
    / 1 skips the check."""
    rng = np.random.default_rng(77)
    d, K, B, N = 9, 2, 3, 21
    h0, hks = _rand_model(rng, d, K, 0.9, hermitian=False)
    sig = rng.uniform(-1, 1, size=(B, K, N))
    Ubar = rng.normal(size=(B, d, d)) + 1j * rng.normal(size=(B, d, d))
    U1, g1 = eng.pwc_closed_grad(h0, hks, sig, 1.0, Ubar)
    eng.set_tuning("grad_variant", 2)
    try:
        U2, g2 = eng.pwc_closed_grad(h0, hks, sig, 1.0, Ubar)
    finally:
        eng.set_tuning("grad_variant", 1)
    assert torch.equal(g1, g2) and torch.equal(U1, U2)
    h0h, hksh = _rand_model(rng, d, K, 0.9)
    ref = eng.pwc_closed_grad(h0h, hksh, sig, 1.0, Ubar)[1]
    for flag in (1, 0):
        eng.set_tuning("grad_unitary", flag)
        try:
            g = eng.pwc_closed_grad(h0h, hksh, sig, 1.0, Ubar)[1]
        finally:
            eng.set_tuning("grad_unitary", -1)
        assert rel_fro(g.cpu().numpy(), ref.cpu().numpy()) < 1e-9


def test_gradient_chunking_and_finite_difference(eng):
    """Chunked passes give the same gradient; a central finite difference agrees to 1e-6."""
    from c3_b200 import synth
    m = synth.two_transmon()
    B, N = 5, 30
    sig = synth.controls(m, B, N)
    rng = np.random.default_rng(9)
    Ubar = rng.normal(size=(B, 9, 9)) + 1j * rng.normal(size=(B, 9, 9))
    U1, g1 = eng.pwc_closed_grad(m.h0, m.hks, sig, 1e-11, Ubar)
    U2, g2 = eng.pwc_closed_grad(m.h0, m.hks, sig, 1e-11, Ubar, max_workspace_bytes=1 << 20)   # forces small chunks
    assert rel_fro(U2.cpu().numpy(), U1.cpu().numpy()) < 1e-13     # the chunk length follows the batch chunk: other rounding
    assert rel_fro(g2.cpu().numpy(), g1.cpu().numpy()) < 1e-11
    b, k, n = 2, 1, 17
    eps = 1e3          # signals are ~1e9 rad/s
    sp, sm = sig.copy(), sig.copy()
    sp[b, k, n] += eps
    sm[b, k, n] -= eps
    Up = eng.pwc_closed(m.h0, m.hks, sp, 1e-11)[b].cpu().numpy()
    Um = eng.pwc_closed(m.h0, m.hks, sm, 1e-11)[b].cpu().numpy()
    fd = np.real(np.sum(np.conj(Ubar[b]) * (Up - Um))) / (2 * eps)
    assert abs(fd - float(g1[b, k, n])) < 1e-6 * max(abs(fd), 1e-12) + 1e-18
