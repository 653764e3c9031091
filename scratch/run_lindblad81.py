"""BASELINE config 3 shape (two 3-level transmons, Lindblad D=81) on a bounded batch -- ncu target."""
import sys
import torch
sys.path.insert(0, ".")
from c3_b200 import engine, synth
m = synth.two_transmon()
B, N = 296, 40
sig = torch.as_tensor(synth.controls_fast(m, B, N)).cuda()
for _ in range(3):
    U = engine.pwc_lindblad(m.h0, m.hks, m.col_ops, sig, 1e-11)
torch.cuda.synchronize()
print("ok", U.shape)
