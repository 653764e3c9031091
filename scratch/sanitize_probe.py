"""Small invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck): sizes chosen so that a
tool pass finishes in seconds.  Usage on a GPU box:  compute-sanitizer --tool racecheck python scratch/sanitize_probe.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from c3_b200 import engine, synth

rng = np.random.default_rng(0)
def model(d, K, scale=0.9):
    def herm():
        h = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)); return h + h.conj().T
    h0 = herm(); h0 *= scale / np.abs(h0).sum(axis=0).max()
    hks = np.stack([herm() for _ in range(K)])
    for k in range(K): hks[k] *= 0.2 * scale / np.abs(hks[k]).sum(axis=0).max()
    return h0, hks

which = sys.argv[1:] or ["d9", "small", "cta27", "lind", "grad9", "gradu", "gradcta", "gated", "misc", "signals"]
if "d9" in which:
    h0, hks = model(9, 2); sig = rng.uniform(-1, 1, (5, 2, 40))
    for v in (0, 1):
        engine.set_tuning("d9_variant", v); engine.pwc_closed(h0, hks, sig, 1.0, return_dUs=True)
    engine.set_tuning("d9_variant", 1)
    engine.pwc_closed(h0 * 8, hks, sig, 1.0)          # squarings
if "small" in which:
    for d in (3, 6, 12):
        h0, hks = model(d, 1); engine.pwc_closed(h0, hks, rng.uniform(-1, 1, (3, 1, 33)), 1.0)
if "cta27" in which:
    h0, hks = model(27, 3, 2.0); engine.pwc_closed(h0, hks, rng.uniform(-1, 1, (2, 3, 10)), 1.0, return_dUs=True)
    h0, hks = model(20, 1); engine.pwc_closed(h0, hks, rng.uniform(-1, 1, (2, 1, 9)), 1.0)
if "lind" in which:
    m = synth.two_transmon(); sig = synth.controls(m, 1, 1000)[:, :, 500:504].copy()
    engine.pwc_lindblad(m.h0, m.hks, m.col_ops, sig, 1e-11)
if "grad9" in which:
    h0, hks = model(9, 2); sig = rng.uniform(-1, 1, (4, 2, 21))
    engine.pwc_closed_grad(h0, hks, sig, 1.0, rng.normal(size=(4, 9, 9)) + 0j)
    engine.pwc_closed_grad(h0 * 6, hks, sig, 1.0, rng.normal(size=(4, 9, 9)) + 0j)
if "gradu" in which:
    h0, hks = model(27, 3, 2.0); sig = rng.uniform(-1, 1, (2, 3, 9))
    engine.pwc_closed_grad(h0, hks, sig, 1.0, rng.normal(size=(2, 27, 27)) + 0j)
if "gradcta" in which:
    engine.set_tuning("grad_variant", 2)
    h0, hks = model(20, 2); sig = rng.uniform(-1, 1, (2, 2, 7))
    engine.pwc_closed_grad(h0, hks, sig, 1.0, rng.normal(size=(2, 20, 20)) + 0j)
    h0, hks = model(9, 2); sig = rng.uniform(-1, 1, (2, 2, 7))
    engine.pwc_closed_grad(h0, hks, sig, 1.0, rng.normal(size=(2, 9, 9)) + 0j)
    engine.set_tuning("grad_variant", 1)
if "gated" in which:
    h0, hks = model(9, 2)
    host = torch.as_tensor(rng.uniform(-1, 1, (7, 2, 37))).pin_memory()
    engine.pwc_closed_from_host(h0, hks, host, 1.0, chunk=2, first_chunk=1)
    engine.check_gated_launches()
if "misc" in which:
    mats = rng.normal(size=(3, 21, 9, 9)) + 1j * rng.normal(size=(3, 21, 9, 9))
    engine.ordered_product(mats)
    engine.ordered_product(rng.normal(size=(2, 5, 70, 70)) + 0j)
    gates = rng.normal(size=(4, 9, 9)) + 0j
    idx = np.array([[0, 1, 2, 3], [3, 3, 0, 0], [1, 0, 0, 0]], dtype=np.int32); ln = np.array([4, 2, 0], dtype=np.int32)
    engine.seq_product(gates, idx, ln)
    engine.seq_populations(gates, idx, ln)
    U = torch.as_tensor(rng.normal(size=(5, 9, 9)) + 1j * rng.normal(size=(5, 9, 9))).cuda()
    engine.gate_infid(U, np.eye(4), [0, 1, 3, 4], "average")
    engine.frame_dephase(U.clone(), np.array([[0, 0, 0, 1, 1, 1, 2, 2, 2], [0, 1, 2, 0, 1, 2, 0, 1, 2]]), rng.normal(size=(5, 2)))
    engine.kron(rng.normal(size=(3, 3)) + 0j, rng.normal(size=(4, 4)) + 0j)
    d = rng.normal(size=(3, 9, 9)); engine.dress_models(d + d.transpose(0, 2, 1) + np.diag(np.arange(9.0) * 5))
    engine.crosstalk(torch.as_tensor(rng.normal(size=(2, 3, 50))).cuda(), [0, 2], [[1, .1], [.2, 1]])
if "signals" in which:
    from c3_b200 import generator as gen_mod
    TP = 2 * np.pi
    env = np.zeros((2, 2, 2, 9)); env[..., 0] = 0.4; env[..., 1] = 7e-9; env[..., 2] = 1.7e-9; env[..., 6] = 1e-9; env[..., 7] = 6e-9; env[..., 8] = 1e-9
    sid = np.array([[gen_mod.SHAPE_IDS["gaussian_nonorm"], gen_mod.SHAPE_IDS["flattop_cut"]],
                    [gen_mod.SHAPE_IDS["fourier_cos"], gen_mod.SHAPE_IDS["slepian_fourier"]]], dtype=np.int32)
    tab = np.zeros((2, 2, 12)); tab[1, 0, :5] = [2, 0.5, 0.2, 1e9, 2e9]; tab[1, 1, :9] = [6e-9, 0.1, 1.5e-9, 2, 1.0, 0.5, 1, 0.3, 0]
    chain = np.tile([100e9, 2e9, 0.3e-9, 1, 0, 1e9, 0, 1, 0, 0, np.nan], (2, 1))
    noise = np.tile([0.01, 0.001, 0.01, 0.01, 0.01, 5, 0.0], (2, 1))
    engine.generate_signals(env, sid, np.zeros((2, 2), np.int32), np.full((2, 2), 5e9 * TP), chain, 0.0, 7e-9, env_table=tab,
                            noise=noise, seed=3, return_noise=True)
    sid2 = np.full((2, 2), gen_mod.SHAPE_IDS["gaussian_nonorm"], dtype=np.int32)
    N = engine.signal_slice_num(0.0, 7e-9, 100e9)
    engine.generate_signals_grad(env, sid2, np.zeros((2, 2), np.int32), np.full((2, 2), 5e9 * TP), chain, 0.0, 7e-9,
                                 torch.ones((2, 2, N), dtype=torch.float64, device="cuda"))
torch.cuda.synchronize()
print("probe done:", which)
