"""cfg2-size launch (d = 9, B = 256, N = 1000): segmentation knobs against whole-call time."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from c3_b200 import engine, synth
m = synth.two_transmon()
for B in (64, 256, 512):
    sig = torch.as_tensor(synth.controls_fast(m, B, 1000)).cuda()
    pm = engine.prepare_model(m.h0, m.hks, 1e-11)
    for mc, tu in ((8, 32768), (4, 32768), (2, 32768), (4, 65536), (3, 65536)):
        engine.set_tuning("min_chunk", mc); engine.set_tuning("target_units", tu)
        for _ in range(3): engine.pwc_prepared(pm, sig)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): engine.pwc_prepared(pm, sig)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"B={B} min_chunk={mc} target_units={tu}: {ms:.4f} ms  {B*1000/ms*1e3:.3e} slices/s")
