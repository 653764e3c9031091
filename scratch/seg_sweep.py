"""Sweep of the time-axis segmentation knobs (target_units, min_chunk) of the lane-group kernels over batch sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from c3_b200 import engine, synth
m = synth.two_transmon()
Bs = [int(x) for x in sys.argv[1:]] or [64, 256, 512, 1024, 2048, 4096]
for B in Bs:
    sig = torch.as_tensor(synth.controls_fast(m, B, 1000)).cuda()
    pm = engine.prepare_model(m.h0, m.hks, 1e-11)
    res = []
    for tu in (0, 2048, 4096, 8192, 12288, 16384, 32768):
        for mc in (8,):
            engine.set_tuning("target_units", tu); engine.set_tuning("min_chunk", mc)
            for _ in range(3): engine.pwc_prepared(pm, sig)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): engine.pwc_prepared(pm, sig)
            e1.record(); torch.cuda.synchronize()
            res.append((e0.elapsed_time(e1) / 10, tu, mc))
    print(B, " ".join(f"{tu}:{t:.4f}" for t, tu, mc in sorted(res, key=lambda r: r[1])))
