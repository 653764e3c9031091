"""BASELINE config 5 shape (tunable coupler d=27, K=3) on a bounded batch -- ncu / timing target."""
import sys
import torch
sys.path.insert(0, ".")
from c3_b200 import engine, synth, flops
m = synth.tunable_coupler()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
sig_np = synth.controls_fast(m, B, N)
sig = torch.as_tensor(sig_np).cuda()
engine.set_tuning("profile", 1)
for _ in range(3):
    U = engine.pwc_closed(m.h0, m.hks, sig, 1e-11)
torch.cuda.synchronize()
ms = engine.last_kernel_ms()
f = flops.flops_per_slice_closed(m.h0, m.hks, sig_np[:2], 1e-11)
print(f"d27 B={B} N={N}: kernel {ms:.3f} ms, {B*N/ms*1e3:.3e} slices/s, {B*N/ms*1e3*f/1e12:.2f} TFLOP/s algorithmic ({f:.0f} flop/slice)")
