import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from c3_b200 import engine, synth
which = sys.argv[1] if len(sys.argv) > 1 else "27"
if which == "27":
    m = synth.tunable_coupler(); B, N = 296, 200
    sig = torch.as_tensor(synth.controls_fast(m, B, N)).cuda()
    Ub = torch.as_tensor(np.random.default_rng(0).normal(size=(B, 27, 27)) + 0j).cuda()
    f = lambda: engine.pwc_closed_grad(m.h0, m.hks, sig, 1e-11, Ub)
else:
    m = synth.two_transmon(); B, N = 148, 40
    sig = torch.as_tensor(synth.controls_fast(m, B, N)).cuda()
    Ub = torch.as_tensor(np.random.default_rng(0).normal(size=(B, 81, 81)) + 0j).cuda()
    f = lambda: engine.pwc_lindblad_grad(m.h0, m.hks, m.col_ops, sig, 1e-11, Ub)
for _ in range(2): f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): f()
e1.record(); torch.cuda.synchronize()
print(f"{which}: forward + gradient: {e0.elapsed_time(e1)/3:.2f} ms for B={B}, N={N}")
