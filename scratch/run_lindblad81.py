"""BASELINE config 3 shape (two 3-level transmons, Lindblad D=81) on a bounded batch -- ncu target / tiling A-B."""
import sys
import torch
sys.path.insert(0, ".")
from c3_b200 import engine, synth
m = synth.two_transmon()
B, N = 296, 40
sig = torch.as_tensor(synth.controls_fast(m, B, N)).cuda()
engine.set_tuning("profile", 1)
ref = None
for v in ([int(x) for x in sys.argv[1:]] or [0]):
    engine.set_tuning("gemm_big", v)
    ts = []
    for _ in range(3):
        U = engine.pwc_lindblad(m.h0, m.hks, m.col_ops, sig, 1e-11); torch.cuda.synchronize(); ts.append(engine.last_kernel_ms())
    if ref is None: ref = U.clone()
    err = float((U - ref).abs().max())
    print("gemm_big", v, "kernel ms", min(ts), "slices/s %.3e" % (B * N / min(ts) * 1e3), "TFLOP/s %.2f" % (B * N * 35.4e6 / min(ts) * 1e3 / 1e12), "max diff vs first", err)
