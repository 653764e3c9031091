"""Headline-shape throughput of the d = 9 kernel variants (device-resident signals, kernel-only timing)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from c3_b200 import engine, synth, flops

m = synth.two_transmon()
B, N = 4096, 1000
sig = torch.as_tensor(synth.controls_fast(m, B, N)).cuda()
peak = engine.measure_fp64_peak("dfma", 0.3)
F = flops.flops_closed(9, 2, 9, 0)
for spec in (sys.argv[1:] or ["1", "2"]):
    variant, _, skew = spec.partition(":")
    variant = int(variant)
    engine.set_tuning("d9_variant", variant)
    if skew:
        engine.set_tuning("d9_skew", int(skew))
    engine.set_tuning("profile", 1)
    for _ in range(3):
        U = engine.pwc_closed(m.h0, m.hks, sig, 1e-11)
    ts = []
    for _ in range(5):
        U = engine.pwc_closed(m.h0, m.hks, sig, 1e-11)
        ts.append(engine.last_kernel_ms())
    ms = float(np.median(ts))
    tf = B * N * F / (ms * 1e-3) / 1e12
    print(f"d9_variant {spec}: kernel {ms:.3f} ms  {B*N/ms*1e3:.3e} slices/s  {tf:.2f} TF = {tf/peak:.3f} of {peak:.2f}")
