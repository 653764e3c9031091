import sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from c3_b200 import engine as eng
from oracle import c3_oracle as orc
def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
def rand_model(rng, d, K, scale):
    def herm():
        h = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)); return h + h.conj().T
    h0 = herm(); hks = np.stack([herm() for _ in range(K)])
    h0 *= scale / np.abs(h0).sum(axis=0).max()
    for k in range(K): hks[k] *= 0.2 * scale / np.abs(hks[k]).sum(axis=0).max()
    return h0, hks
for nb in (0, 1):
    for d, scale in [(14, 0.8), (14, 3.0), (27, 2.5), (27, 20.0), (33, 6.0)]:
        rng = np.random.default_rng(d + int(scale * 10))
        K, B, N = 2, 3, 7
        h0, hks = rand_model(rng, d, K, scale)
        sig = rng.uniform(-1, 1, size=(B, K, N))
        eng.set_tuning("norm_bound", nb)
        U, dUs = eng.pwc_closed(h0, hks, sig, 1.0, return_dUs=True)
        U2 = eng.pwc_closed(h0, hks, sig, 1.0)
        wU, wd = orc.propagate_batch(h0, hks, sig, 1.0, return_dUs=True)
        print("nb", nb, "d", d, "scale", scale, "U", rel(U.cpu().numpy(), wU), "U(no dUs)", rel(U2.cpu().numpy(), wU), "dUs", rel(dUs.cpu().numpy(), wd), "dUs[0,0] err", rel(dUs[0, 0].cpu().numpy(), wd[0, 0]))
