"""Throughput of the DMMA CTA kernels on the cfg3 (Lindblad D = 81) and cfg5 (d = 27) shapes (reduced batch, full algorithm)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from c3_b200 import engine, synth, flops

which = sys.argv[1:] or ["81", "27"]
if "t512" in which: engine.set_tuning("cta_threads", 512)
if "big2" in which: engine.set_tuning("gemm_big", 2)
if "big1" in which: engine.set_tuning("gemm_big", 1)
if "big3" in which: engine.set_tuning("gemm_big", 3)
if "big4" in which: engine.set_tuning("gemm_big", 4)
if "big5" in which: engine.set_tuning("gemm_big", 5)
peak = engine.measure_fp64_peak("dfma", 0.3)
engine.set_tuning("profile", 1)
if "81" in which:
    m = synth.two_transmon()
    B, N = 592, 200
    sig = torch.as_tensor(synth.controls_fast(m, B, N)).cuda()
    pm = engine.prepare_model(m.h0, m.hks, 1e-11, col_ops=m.col_ops, lindblad=True)
    F = flops.flops_lindblad(9, 13, 0)
    ts = []
    for _ in range(3):
        U = engine.pwc_prepared(pm, sig); ts.append(engine.last_kernel_ms())
    ms = min(ts[1:])
    print(f"D=81  B={B} N={N}: {ms:.2f} ms  {B*N/ms*1e3:.3e} slices/s  {B*N*F/ms/1e9:.2f} TF = {B*N*F/ms/1e9/peak:.3f}")
if "27" in which:
    m = synth.tunable_coupler()
    B, N = 1184, 400
    sig = torch.as_tensor(synth.controls_fast(m, B, N)).cuda()
    pm = engine.prepare_model(m.h0, m.hks, 1e-11)
    F = flops.flops_closed(27, 3, 13, 0)
    ts = []
    for _ in range(3):
        U = engine.pwc_prepared(pm, sig); ts.append(engine.last_kernel_ms())
    ms = min(ts[1:])
    print(f"d=27  B={B} N={N}: {ms:.2f} ms  {B*N/ms*1e3:.3e} slices/s  {B*N*F/ms/1e9:.2f} TF = {B*N*F/ms/1e9/peak:.3f}")
