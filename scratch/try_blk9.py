"""Parity + timing of the own-block d=9 kernel (rows_variant 15) against the shipped one (13)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from c3_b200 import engine, synth
from oracle import c3_oracle as orc

def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
m = synth.two_transmon()
VS = [int(x) for x in sys.argv[1:]] or [13, 15]
for v in VS:
    engine.set_tuning("rows_variant", v)
    sig = synth.controls(m, 3, 203)
    U, dUs = engine.pwc_closed(m.h0, m.hks, sig, 1e-11, return_dUs=True)
    wU, wd = orc.propagate_batch(m.h0, m.hks, sig, 1e-11, return_dUs=True)
    print(v, "d9 U", rel(U.cpu().numpy(), wU), "dUs", rel(dUs.cpu().numpy(), wd))
    # squarings + padded d=7 + H-list
    rng = np.random.default_rng(5)
    for d, scale in ((9, 6.0), (7, 2.0), (9, 30.0)):
        h0 = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d)); h0 = (h0 + h0.conj().T) * scale / d
        hks = rng.normal(size=(2, d, d)) + 1j * rng.normal(size=(2, d, d)); hks = (hks + hks.conj().transpose(0, 2, 1)) / d
        sg = rng.uniform(-1, 1, size=(2, 2, 37))
        U, dUs = engine.pwc_closed(h0, hks, sg, 1.0, return_dUs=True)
        wU, wd = orc.propagate_batch(h0, hks, sg, 1.0, return_dUs=True)
        print(v, "rand d", d, scale, rel(U.cpu().numpy(), wU), rel(dUs.cpu().numpy(), wd))
        Hs = h0[None, None] + np.einsum("bkn,kij->bnij", sg, hks)
        U2 = engine.pwc_closed_hlist(Hs, 1.0)
        print(v, "hlist d", d, rel(U2.cpu().numpy(), wU))
B, N = 4096, 1000
sig = torch.as_tensor(synth.controls(m, B, N)).cuda()
engine.set_tuning("profile", 1)
for v in VS + VS:
    engine.set_tuning("rows_variant", v)
    ts = []
    for _ in range(4):
        U = engine.pwc_closed(m.h0, m.hks, sig, 1e-11)
        torch.cuda.synchronize(); ts.append(engine.last_kernel_ms())
    print("variant", v, "kernel ms", ts, "slices/s %.3e" % (B * N / min(ts) * 1e3))
    if v == VS[0]: U13 = U.clone()
    else: print(v, "vs 13", rel(U.cpu().numpy(), U13.cpu().numpy()))
