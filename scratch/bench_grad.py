"""Forward + gradient timing at the headline shape (reduced batch) and the kernel breakdown under an ncu launch list."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from c3_b200 import engine, synth
m = synth.two_transmon()
B, N = 1024, 1000
sig = torch.as_tensor(synth.controls_fast(m, B, N)).cuda()
Ub = torch.as_tensor(np.random.default_rng(0).normal(size=(B, 9, 9)) + 0j).cuda()
for _ in range(2):
    U, g = engine.pwc_closed_grad(m.h0, m.hks, sig, 1e-11, Ub)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    U, g = engine.pwc_closed_grad(m.h0, m.hks, sig, 1e-11, Ub)
e1.record(); torch.cuda.synchronize()
print(f"forward + gradient: {e0.elapsed_time(e1)/3:.2f} ms for B={B}, N={N}")
