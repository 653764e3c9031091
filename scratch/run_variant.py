"""Run the headline shape a few times with a chosen small-d kernel variant (for ncu captures)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from c3_b200 import engine, synth
v = int(sys.argv[1]) if len(sys.argv) > 1 else 13
tu = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
N = int(sys.argv[4]) if len(sys.argv) > 4 else 1000
m = synth.two_transmon()
sig = torch.as_tensor(synth.controls(m, B, N)).cuda()
engine.set_tuning("rows_variant", v)
engine.set_tuning("target_units", tu)
engine.set_tuning("profile", 1)
for _ in range(3):
    U = engine.pwc_closed(m.h0, m.hks, sig, 1e-11)
torch.cuda.synchronize()
print("variant", v, "kernel ms", engine.last_kernel_ms())
