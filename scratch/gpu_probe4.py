import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from c3_b200 import engine, synth, flops
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
m27 = synth.tunable_coupler(); m9 = synth.two_transmon()
for v in (0, 1):
    engine.set_tuning("cta_variant", v)
    B, N = 512, 200
    sig = torch.as_tensor(synth.controls_fast(m27, B, N)).cuda()
    h0 = torch.as_tensor(m27.h0).cuda(); hks = torch.as_tensor(m27.hks).cuda()
    ms = timeit(lambda: engine.pwc_closed(h0, hks, sig, 1e-11))
    f = flops.flops_closed(27, 3, 13, 0)
    print(f"cta_variant {v} d=27 B={B} N={N}: {ms:.1f} ms {B*N/(ms*1e-3):.3e} slices/s  alg {B*N/(ms*1e-3)*f/1e12:.2f} TF", flush=True)
    B, N = 296, 40
    sig = torch.as_tensor(synth.controls_fast(m9, B, N)).cuda()
    ms = timeit(lambda: engine.pwc_lindblad(m9.h0, m9.hks, m9.col_ops, sig, 1e-11), 2)
    f = flops.flops_lindblad(9, 13, 0)
    print(f"cta_variant {v} Lindblad D=81 B={B} N={N}: {ms:.1f} ms {B*N/(ms*1e-3):.3e} slices/s  alg {B*N/(ms*1e-3)*f/1e12:.2f} TF", flush=True)
