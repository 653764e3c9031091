"""A/B of the small-d fused kernels on the headline shape (run on the GPU box)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from c3_b200 import engine, synth
from oracle import c3_oracle as orc

m = synth.two_transmon()
dt = 1e-11
# parity on ragged shapes
for (B, N) in [(3, 37), (5, 1000), (2, 9), (1, 1), (7, 130)]:
    sig = synth.controls(m, B, N)
    want = orc.propagate_batch(m.h0, m.hks, sig, dt)
    for v in (13, 15):
        engine.set_tuning("rows_variant", v)
        U, dUs = engine.pwc_closed(m.h0, m.hks, sig, dt, return_dUs=True)
        U2 = engine.pwc_closed(m.h0, m.hks, sig, dt)
        e = np.linalg.norm(U.cpu().numpy() - want) / np.linalg.norm(want)
        e2 = np.linalg.norm(U2.cpu().numpy() - want) / np.linalg.norm(want)
        # dUs product check
        P = engine.ordered_product(dUs)
        e3 = np.linalg.norm(P.cpu().numpy() - want) / np.linalg.norm(want)
        print(f"B={B} N={N} variant {v}: rel err {e:.2e} {e2:.2e} dUs-product {e3:.2e}")
# large-norm (squarings) check
sigb = synth.controls(m, 4, 64) * 40.0
want = orc.propagate_batch(m.h0, m.hks, sigb, dt)
for v in (13, 15):
    engine.set_tuning("rows_variant", v)
    U = engine.pwc_closed(m.h0, m.hks, sigb, dt)
    print("big norm variant", v, np.linalg.norm(U.cpu().numpy() - want) / np.linalg.norm(want))

B, N = 4096, 1000
sig = torch.as_tensor(synth.controls(m, B, N)).cuda()
engine.set_tuning("profile", 1)
for v, tus in ((13, (32768,)), (15, (32768,))):
    engine.set_tuning("rows_variant", v)
    for tu in tus:
        engine.set_tuning("target_units", tu)
        for _ in range(3):
            U = engine.pwc_closed(m.h0, m.hks, sig, dt)
        torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            U = engine.pwc_closed(m.h0, m.hks, sig, dt)
            torch.cuda.synchronize()
            ms.append(engine.last_kernel_ms())
        t0 = time.perf_counter()
        for _ in range(5):
            U = engine.pwc_closed(m.h0, m.hks, sig, dt)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / 5
        print(f"variant {v} target_units {tu}: kernel {np.median(ms):.3f} ms, call {wall*1e3:.3f} ms -> {B*N/wall:.3e} slices/s")
