import json, sys
import numpy as np, torch
sys.path.insert(0, ".")
from c3_b200 import engine, synth
from oracle import c3_oracle as orc
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
m = synth.two_transmon()
h0 = torch.as_tensor(m.h0).cuda(); hks = torch.as_tensor(m.hks).cuda()
sig_small = synth.controls(m, 4, 1000)
want = orc.propagate_batch(m.h0, m.hks, sig_small, 1e-11)
B = 4096
sig = torch.as_tensor(synth.controls_fast(m, B, 1000)).cuda()
for v in (0, 1, 2, 3):
    engine.set_tuning("rows_variant", v)
    U = engine.pwc_closed(h0, hks, sig_small, 1e-11).cpu().numpy()
    err = np.linalg.norm(U - want) / np.linalg.norm(want)
    for tu in (16384, 32768, 65536):
        engine.set_tuning("target_units", tu)
        ms = timeit(lambda: engine.pwc_closed(h0, hks, sig, 1e-11))
        print(f"variant {v} target_units {tu}: {ms:.2f} ms  {B*1000/(ms*1e-3)/1e6:.1f} Mslices/s  alg {B*1000/(ms*1e-3)*43.4e3/1e12:.2f} TF  err {err:.2e}", flush=True)
