"""The BASELINE.json configurations AT THEIR STATED SIZE on the GPU, checked row by row against the CPU oracle on a seeded
sample of batch rows (all rows for the small configs, >= 16 random rows for the large ones), plus the
size-independent properties the domain offers (trace preservation of Lindblad maps, unitarity)."""
import numpy as np
import pytest
import scipy.linalg
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


def _expm_each(a):
    """Slice-wise scipy expm (Higham 2005/2009): the oracle's product chain with a faster exponential for the big rows.
    tests/test_oracle_golden.py pins expm_tf against scipy to 1e-13 in this norm range."""
    return np.stack([scipy.linalg.expm(x) for x in a])


def test_config2_d9_b256_all_rows(eng):
    """two-qubit d = 9, 1000 slices, batch = 256: every row against the oracle."""
    from c3_b200 import synth
    m = synth.two_transmon()
    sig = synth.controls_fast(m, 256, 1000)
    U = eng.pwc_closed(m.h0, m.hks, sig, 1e-11).cpu().numpy()
    want = orc.propagate_batch(m.h0, m.hks, sig, 1e-11, expm=_expm_each)
    errs = [rel_fro(U[b], want[b]) for b in range(256)]
    assert max(errs) < TOL, max(errs)


def test_config3_lindblad_d81_full_size_sampled_rows(eng):
    """Lindblad D = 81, 1000 slices, batch = 1024 in ONE call; 16 seeded rows against the oracle over all 1000 slices,
    every row trace preserving."""
    from c3_b200 import synth
    m = synth.two_transmon()
    B, N = 1024, 1000
    sig = synth.controls_fast(m, B, N)
    U = eng.pwc_lindblad(m.h0, m.hks, m.col_ops, sig, 1e-11)
    torch.cuda.synchronize()
    rows = np.random.default_rng(81).choice(B, size=16, replace=False)
    Uh = U[torch.as_tensor(rows, device=U.device)].cpu().numpy()
    for i, b in enumerate(rows):
        want = orc.propagate_batch(m.h0, m.hks, sig[b:b + 1], 1e-11, col_ops=m.col_ops, lindbladian=True, expm=_expm_each)[0]
        assert rel_fro(Uh[i], want) < TOL, (int(b), rel_fro(Uh[i], want))
    # tr(rho) is conserved: vec(I)^T S = vec(I)^T for every superoperator of the batch
    d = 9
    vec_id = torch.eye(d, dtype=torch.complex128, device=U.device).reshape(-1)
    lhs = torch.einsum("i,bij->bj", vec_id, U)
    assert float((lhs - vec_id).abs().max()) < 1e-9


def test_config5_d27_shard_sampled_rows(eng):
    """tunable coupler d = 27, K = 3, 2000 slices: one GPU's shard of the 8192-sample batch (1024 rows) in one call;
    16 seeded rows against the oracle, every row unitary."""
    from c3_b200 import synth
    m = synth.tunable_coupler()
    B, N = 1024, 2000
    sig = synth.controls_fast(m, B, N)
    U = eng.pwc_closed(m.h0, m.hks, sig, 1e-11)
    torch.cuda.synchronize()
    rows = np.random.default_rng(27).choice(B, size=16, replace=False)
    Uh = U[torch.as_tensor(rows, device=U.device)].cpu().numpy()
    for i, b in enumerate(rows):
        want = orc.propagate_batch(m.h0, m.hks, sig[b:b + 1], 1e-11, expm=_expm_each)[0]
        assert rel_fro(Uh[i], want) < TOL, (int(b), rel_fro(Uh[i], want))
    eye = torch.eye(27, dtype=torch.complex128, device=U.device)
    dev = (U.conj().transpose(1, 2) @ U - eye).abs().amax(dim=(1, 2))
    assert float(dev.max()) < 1e-9


def test_config4_orbit_sequences_full_size(eng):
    """ORBIT: 4096 random Clifford sequences x 20 Cliffords (~45 native gates each) at d = 9 in one launch; 64 seeded
    sequences against the oracle's evaluate_sequences, identity for the empty sequence."""
    from c3_b200 import synth
    m = synth.two_transmon()
    gate_sig = synth.controls(m, 5, 70)
    gates = eng.pwc_closed(m.h0, m.hks, gate_sig, 1e-11)
    idx, lens = synth.rb_sequences(4096, 20, 5, seed=0)
    lens = lens.copy()
    lens[17] = 0
    out = eng.seq_product(gates, idx, lens)
    gh = gates.cpu().numpy()
    names = [f"g{i}" for i in range(5)]
    table = {n: gh[i] for i, n in enumerate(names)}
    rows = np.random.default_rng(4).choice(4096, size=63, replace=False).tolist() + [17]
    seqs = [[names[j] for j in idx[s, :lens[s]]] for s in rows]
    want = orc.evaluate_sequences(table, seqs)
    got = out[torch.as_tensor(rows, device=out.device)].cpu().numpy()
    for i in range(len(rows)):
        assert rel_fro(got[i], want[i]) < TOL
    assert np.allclose(got[-1], np.eye(9))


def test_config1_single_qubit_latency_path(eng):
    """single-qubit 3-level rx90p, batch = 1, 50 slices (and the 800 the hjson yields): prepared model + captured graph."""
    from c3_b200 import synth
    m = synth.one_qubit()
    pm = eng.prepare_model(m.h0, m.hks, 1e-11)
    for N in (50, 800):
        sig = synth.controls(m, 1, N)
        gp = eng.GraphedPwc(pm, 1, N)
        U = gp.run(torch.as_tensor(sig).cuda())
        want = orc.propagate_batch(m.h0, m.hks, sig, 1e-11)
        assert rel_fro(U.cpu().numpy(), want) < TOL
