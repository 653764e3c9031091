"""Throughput of the gradient path (f-1 [+ f-3, f-2]) on the headline shape."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from c3_b200 import engine, synth
m = synth.two_transmon()
N = 1000
for B in (256, 1024, 4096):
    sig = torch.as_tensor(synth.controls_fast(m, B, N)).cuda()
    Ubar = torch.randn(B, 9, 9, dtype=torch.complex128, device="cuda")
    for _ in range(2):
        U, g = engine.pwc_closed_grad(m.h0, m.hks, sig, 1e-11, Ubar)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 3
    for _ in range(n):
        U, g = engine.pwc_closed_grad(m.h0, m.hks, sig, 1e-11, Ubar)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for _ in range(n):
        U = engine.pwc_closed(m.h0, m.hks, sig, 1e-11)
    torch.cuda.synchronize()
    df = (time.perf_counter() - t0) / n
    print(f"B={B}: forward {df*1e3:.2f} ms ({B*N/df:.3e} slices/s); forward+gradient {dt*1e3:.2f} ms ({B*N/dt:.3e} slices/s), ratio {dt/df:.1f}")
