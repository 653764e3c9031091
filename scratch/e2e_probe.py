import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from c3_b200 import engine, synth, propagation
m = synth.two_transmon()
N = 1000
engine.set_tuning("profile", 1)
for B in (256, 512, 1024, 2048, 4096):
    sig = torch.as_tensor(synth.controls(m, B, N)).cuda()
    ts = []
    for _ in range(4):
        U = engine.pwc_closed(m.h0, m.hks, sig, 1e-11); torch.cuda.synchronize(); ts.append(engine.last_kernel_ms())
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): U = engine.pwc_closed(m.h0, m.hks, sig, 1e-11)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B}: kernel {min(ts):.3f} ms ({B*N/min(ts)*1e3:.3e}/s), whole call {e0.elapsed_time(e1)/5:.3f} ms")
engine.set_tuning("profile", 0)
B = 4096
host = torch.as_tensor(synth.controls(m, B, N)).pin_memory()
Uh = torch.empty((B, 9, 9), dtype=torch.complex128).pin_memory()
for chunk in (512, 1024, 2048):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            U = engine.pwc_closed_from_host(m.h0, m.hks, host, 1e-11, chunk=chunk)
            Uh.copy_(U, non_blocking=True); torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
    engine.set_tuning("profile", 1)
    U = engine.pwc_closed_from_host(m.h0, m.hks, host, 1e-11, chunk=chunk); torch.cuda.synchronize()
    print(f"chunk {chunk}: e2e {dt*1e3:.3f} ms/step  ({B*N/dt:.3e}/s); gated kernel {engine.last_kernel_ms():.3f} ms")
    engine.set_tuning("profile", 0)
for first, chunk in ((256, 4096), (512, 4096), (1024, 4096), (512, 1792), (1024, 1536)):
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            U = engine.pwc_closed_from_host(m.h0, m.hks, host, 1e-11, chunk=chunk, first_chunk=first, gated=False)
            Uh.copy_(U, non_blocking=True); torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 5
    print(f"ungated chunks first {first} then {chunk}: e2e {dt*1e3:.3f} ms/step  ({B*N/dt:.3e}/s)")
# raw copies
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): d = host.to("cuda", non_blocking=True); torch.cuda.synchronize()
print("H2D 65.5MB ms", (time.perf_counter() - t0) / 5 * 1e3)
t0 = time.perf_counter()
for _ in range(5): Uh.copy_(U, non_blocking=True); torch.cuda.synchronize()
print("D2H 5.3MB ms", (time.perf_counter() - t0) / 5 * 1e3)
