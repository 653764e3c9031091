import sys
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
want, want_d = orc.propagate_batch(m.h0, m.hks, sig_small, 1e-11, return_dUs=True)
B = 4096
sig = torch.as_tensor(synth.controls_fast(m, B, 1000)).cuda()
for v in (8, 13, 14):
    engine.set_tuning("rows_variant", v)
    for mc in (8,):
        engine.set_tuning("min_chunk", mc)
        U, dUs = engine.pwc_closed(h0, hks, sig_small, 1e-11, return_dUs=True)
        err = np.linalg.norm(U.cpu().numpy() - want) / np.linalg.norm(want)
        errd = np.linalg.norm(dUs.cpu().numpy() - want_d) / np.linalg.norm(want_d)
        ms = timeit(lambda: engine.pwc_closed(h0, hks, sig, 1e-11))
        print(f"variant {v} min_chunk {mc}: {ms:.2f} ms  {B*1000/(ms*1e-3)/1e6:.1f} Mslices/s  alg {B*1000/(ms*1e-3)*43.4e3/1e12:.2f} TF  errU {err:.2e} errdU {errd:.2e}", flush=True)
