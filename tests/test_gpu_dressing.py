"""GPU parity of the batched dressing kernel against the oracle
(oracle/c3_model_oracle.py: eigh + reorder_frame + T^dag X T, c3/model.py:453-534) and the reference's pickled
energy-level sweeps of the tunable-coupler chip."""
import numpy as np
import pytest
import torch

from conftest import rel_fro
from oracle import c3_model_oracle as mo
from test_oracle_golden import _tc_level_sweep

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from c3_b200 import engine
    return engine


def test_pickled_energy_level_sweeps(eng, golden_tc_levels):
    """test/test_tunable_coupler.py:315-383 through the CUDA kernel (d = 27)."""
    g = golden_tc_levels

    def dress(h, ordered):
        return eng.dress_models(h, ordered=ordered)["eigenframe"].cpu().numpy()

    prod, ordd, dres, fallback = _tc_level_sweep(g, dress)
    assert np.abs(prod - g["product_basis"]).max() < 1e-9
    assert np.abs(dres - g["dressed_basis"]).max() < 1e-9
    assert np.abs(ordd[~fallback] - g["ordered_basis"][~fallback]).max() < 1e-9
    assert np.abs(ordd - g["ordered_basis"]).max() < 1.0


@pytest.mark.parametrize("d,dims", [(9, [3, 3]), (4, [2, 2]), (27, [3, 3, 3]), (3, [3])])
def test_batched_dressing_matches_oracle(eng, d, dims):
    """Per-sample model parameters (frequencies, anharmonicities, couplings drawn per batch element): eigenframe,
    transform, dressed drift and dressed control / collapse operators against the oracle, sample by sample."""
    rng = np.random.default_rng(d)
    B = 17
    a = mo.annihilators(dims)
    drifts = []
    for b in range(B):
        h = np.zeros((d, d), complex)
        for i, ai in enumerate(a):
            h = h + 2 * np.pi * rng.uniform(4e9, 6e9) * mo.resonator(ai) + 2 * np.pi * rng.uniform(-3e8, -2e8) * mo.duffing(ai)
        for i in range(len(a) - 1):
            h = h + 2 * np.pi * rng.uniform(10e6, 40e6) * mo.int_XX(a[i], a[i + 1])
        drifts.append(h)
    drifts = np.stack(drifts)
    ops = np.stack([mo.x_drive(ai) for ai in a] + [mo.qubit_collapse_op(ai, 27e-6, 39e-6) for ai in a])
    out = eng.dress_models(drifts, ops, ordered=True)
    assert int(out["sweeps"].max()) < 30
    for b in range(B):
        ef, T = mo.dressing_transform(drifts[b])
        scale = np.abs(ef).max()
        assert np.abs(out["eigenframe"][b].cpu().numpy() - ef).max() < 1e-12 * scale
        assert rel_fro(out["transform"][b].cpu().numpy(), T) < 1e-10
        assert rel_fro(out["drift"][b].cpu().numpy(), mo.dress(T, drifts[b])) < 1e-10
        for m in range(ops.shape[0]):
            assert rel_fro(out["ops"][b, m].cpu().numpy(), mo.dress(T, ops[m])) < 1e-10
    un = eng.dress_models(drifts, ordered=False)
    for b in range(B):
        assert np.abs(un["eigenframe"][b].cpu().numpy() - np.linalg.eigvalsh(drifts[b])).max() < 1e-12 * np.abs(drifts[b]).max()


def test_complex_hermitian_and_degenerate(eng):
    """Complex Hermitian drifts (no reference phase convention exists): T is unitary and diagonalises the drift;
    exactly degenerate and already-diagonal inputs converge."""
    rng = np.random.default_rng(0)
    d, B = 12, 9
    h = rng.normal(size=(B, d, d)) + 1j * rng.normal(size=(B, d, d))
    h = h + np.conj(np.swapaxes(h, 1, 2)) + 40 * np.diag(np.arange(d))[None]
    h[0] = np.diag(np.arange(d) // 2).astype(complex)          # pairwise degenerate, diagonal
    h[1] = np.eye(d)
    out = eng.dress_models(h, ordered=True)
    T = out["transform"].cpu().numpy()
    for b in range(B):
        assert np.abs(T[b].conj().T @ T[b] - np.eye(d)).max() < 1e-12
        D = T[b].conj().T @ h[b] @ T[b]
        assert np.abs(D - np.diag(np.diag(D))).max() < 1e-11 * np.abs(h[b]).max()
        assert np.allclose(np.sort(out["eigenframe"][b].cpu().numpy()), np.linalg.eigvalsh(h[b]), atol=1e-10)
        assert np.abs(np.real(np.diag(D)) - out["eigenframe"][b].cpu().numpy()).max() < 1e-10


def test_dressed_samples_feed_the_propagator(eng):
    """Model samples -> dressed h0 / hks per sample -> batched-model propagators, against the oracle chain."""
    from oracle import c3_oracle as orc
    from c3_b200 import synth
    rng = np.random.default_rng(3)
    dims, B, N = [3, 3], 4, 30
    a = mo.annihilators(dims)
    drifts = np.stack([2 * np.pi * (5e9 + 1e7 * b) * mo.resonator(a[0]) + 2 * np.pi * -2.1e8 * mo.duffing(a[0])
                       + 2 * np.pi * 5.6e9 * mo.resonator(a[1]) + 2 * np.pi * -2.4e8 * mo.duffing(a[1])
                       + 2 * np.pi * 2e7 * mo.int_XX(a[0], a[1]) for b in range(B)])
    ops = np.stack([mo.x_drive(ai) for ai in a])
    out = eng.dress_models(drifts, ops)
    sig = synth.controls(synth.two_transmon(), B, N)
    U = eng.pwc_closed(out["drift"], out["ops"], sig, 1e-11)
    for b in range(B):
        _, T = mo.dressing_transform(drifts[b])
        want = orc.propagate_batch(mo.dress(T, drifts[b]), np.stack([mo.dress(T, o) for o in ops]), sig[b:b + 1], 1e-11)[0]
        assert rel_fro(U[b].cpu().numpy(), want) < 1e-10
