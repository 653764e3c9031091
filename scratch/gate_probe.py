import sys
import numpy as np, torch
sys.path.insert(0, ".")
from c3_b200 import engine, synth, _lib
from c3_b200.engine import _ptr, _workspace, _stream
lib = _lib.load()
m = synth.two_transmon(); B, N, K, d = 4096, 1000, 2, 9
sig = torch.as_tensor(synth.controls(m, B, N)).cuda()
h0 = torch.as_tensor(m.h0).cuda(); hks = torch.as_tensor(m.hks).cuda()
U = torch.empty((B, d, d), dtype=torch.complex128, device="cuda")
ready = torch.full((1,), B, dtype=torch.int32, device="cuda")
ws = _workspace(lib.c3b_pwc_workspace_bytes(B, K, N, d, 0, 0), torch.device("cuda:0"))
engine.set_tuning("profile", 1)
for _ in range(4):
    _lib.check(lib.c3b_pwc_closed_gated(_ptr(h0), _ptr(hks), _ptr(sig), 1e-11, B, K, N, d, _ptr(U), _ptr(ready), _ptr(ws), ws.numel(), _stream()))
    torch.cuda.synchronize(); print("gated kernel, data resident:", engine.last_kernel_ms())
U2 = engine.pwc_closed(m.h0, m.hks, sig, 1e-11); torch.cuda.synchronize(); print("plain:", engine.last_kernel_ms(), float((U - U2).abs().max()))
