import sys, torch
sys.path.insert(0, ".")
from c3_b200 import engine, synth, flops
m = synth.tunable_coupler(); B, N = 592, 100
sig_np = synth.controls_fast(m, B, N); sig = torch.as_tensor(sig_np).cuda()
engine.set_tuning("profile", 1)
f = flops.flops_per_slice_closed(m.h0, m.hks, sig_np[:2], 1e-11)
for t in (512, 256, 512, 256):
    engine.set_tuning("cta_threads", t)
    for _ in range(3): U = engine.pwc_closed(m.h0, m.hks, sig, 1e-11)
    torch.cuda.synchronize(); ms = engine.last_kernel_ms()
    print(t, f"{ms:.3f} ms {B*N/ms*1e3*f/1e12:.2f} TFLOP/s")
