"""All five BASELINE.json configs on one GPU (per-GPU shard for the 8-GPU ones): throughput, algorithmic
TFLOP/s and a sampled parity check against the oracle.  Writes gpurun_out/configs.json."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from c3_b200 import engine, synth, flops, propagation as prop
from oracle import c3_oracle as orc

def timeit(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

def rel(a, b): return float(np.linalg.norm(a - b) / np.linalg.norm(b))
out = {}
peak = engine.measure_fp64_peak("dfma", 0.5)
out["fp64_peak_tflops"] = peak

# cfg1: single 3-level qubit, B=1 (latency case)
m = synth.one_qubit()
for N in (50, 800):
    sig = torch.as_tensor(synth.controls(m, 1, N)).cuda()
    h0 = torch.as_tensor(m.h0).cuda(); hks = torch.as_tensor(m.hks).cuda()
    ms = timeit(lambda: engine.pwc_closed(h0, hks, sig, 1e-11), 20)
    U = engine.pwc_closed(h0, hks, sig, 1e-11).cpu().numpy()
    want = orc.propagate_batch(m.h0, m.hks, sig.cpu().numpy(), 1e-11)
    out[f"cfg1_d3_N{N}_B1"] = {"ms_per_call": ms, "slices_per_s": N / (ms * 1e-3), "parity": rel(U, want)}

# cfg2: d=9, N=1000, B=256
m = synth.two_transmon()
sig_np = synth.controls_fast(m, 256, 1000)
sig = torch.as_tensor(sig_np).cuda(); h0 = torch.as_tensor(m.h0).cuda(); hks = torch.as_tensor(m.hks).cuda()
ms = timeit(lambda: engine.pwc_closed(h0, hks, sig, 1e-11), 10)
U = engine.pwc_closed(h0, hks, sig, 1e-11)
rows = [0, 100, 255]
want = orc.propagate_batch(m.h0, m.hks, sig_np[rows], 1e-11)
f = flops.flops_per_slice_closed(m.h0, m.hks, sig_np[:4], 1e-11)
out["cfg2_d9_N1000_B256"] = {"ms": ms, "slices_per_s": 256e3 / (ms * 1e-3), "alg_tflops": 256e3 / (ms * 1e-3) * f / 1e12,
                             "parity": rel(U[rows].cpu().numpy(), want)}

# cfg3: Lindblad D=81, N=1000, B=1024
B, N = 1024, 1000
sig_np = synth.controls_fast(m, B, N)
sig = torch.as_tensor(sig_np).cuda()
t0 = time.time(); UL = engine.pwc_lindblad(m.h0, m.hks, m.col_ops, sig, 1e-11); torch.cuda.synchronize(); first = time.time() - t0
ms = timeit(lambda: engine.pwc_lindblad(m.h0, m.hks, m.col_ops, sig, 1e-11), 1)
b = 517
Nchk = 40    # oracle check on a prefix of the time axis (the full 1000-slice oracle takes minutes)
Uchk = engine.pwc_lindblad(m.h0, m.hks, m.col_ops, sig[b:b + 1, :, :Nchk].contiguous(), 1e-11)[0].cpu().numpy()
dl = orc.tf_batch_propagate(m.h0, m.hks, sig_np[b, :, :Nchk], 1e-11, Nchk, col_ops=m.col_ops, lindbladian=True)
want = orc.tf_matmul_n(dl, orc.compute_folding_stack(Nchk))
fl = flops.flops_lindblad(9, 13, 0)
# trace preservation of the full-size result: sum_i U[(i,i),(j,k)] = delta_jk (vectorised identity is a left eigenvector)
vec_id = torch.eye(9, dtype=torch.complex128, device=UL.device).reshape(-1)
tp = float((vec_id @ UL - vec_id).abs().max())
out["cfg3_lindblad_D81_N1000_B1024"] = {"ms": ms, "slices_per_s": B * N / (ms * 1e-3), "alg_tflops": B * N / (ms * 1e-3) * fl / 1e12,
                                        "parity_prefix40": rel(Uchk, want), "trace_preservation_err": tp}

# cfg4: 4096 random Clifford sequences x 20 (mean 2.25 native gates each), d=9
gates_sig = torch.as_tensor(synth.controls(m, 5, 700)).cuda()
gates = engine.pwc_closed(h0, hks, gates_sig, 1e-11)          # a 5-gate dictionary
idx, lens = synth.rb_sequences(4096, 20, 5, seed=0)
idx_d = torch.as_tensor(idx).cuda(); lens_d = torch.as_tensor(lens).cuda()
ms = timeit(lambda: engine.seq_product(gates, idx_d, lens_d), 10)
Us = engine.seq_product(gates, idx_d, lens_d)
g_np = gates.cpu().numpy()
chk = [0, 7, 4095]
errs = []
for s_ in chk:
    w = np.eye(9, dtype=complex)
    for j in range(lens[s_]): w = g_np[idx[s_, j]] @ w
    errs.append(rel(Us[s_].cpu().numpy(), w))
out["cfg4_orbit_4096seq_d9"] = {"ms": ms, "sequences_per_s": 4096 / (ms * 1e-3), "gate_products_per_s": float(lens.sum()) / (ms * 1e-3),
                                "mean_len": float(lens.mean()), "parity": max(errs)}

# cfg5: d=27, K=3, N=2000, B=8192 over 8 GPUs -> 1024 per GPU
m27 = synth.tunable_coupler()
B, N = 1024, 2000
sig_np = synth.controls_fast(m27, B, N)
sig = torch.as_tensor(sig_np).cuda(); h27 = torch.as_tensor(m27.h0).cuda(); hk27 = torch.as_tensor(m27.hks).cuda()
ms = timeit(lambda: engine.pwc_closed(h27, hk27, sig, 1e-11), 2)
U = engine.pwc_closed(h27, hk27, sig, 1e-11)
want = orc.propagate_batch(m27.h0, m27.hks, sig_np[[3]], 1e-11)
f27 = flops.flops_per_slice_closed(m27.h0, m27.hks, sig_np[:2, :, :200], 1e-11)
eye = torch.eye(27, dtype=torch.complex128, device=U.device)
out["cfg5_d27_N2000_B1024_per_gpu"] = {"ms": ms, "slices_per_s": B * N / (ms * 1e-3), "alg_tflops": B * N / (ms * 1e-3) * f27 / 1e12,
                                       "parity": rel(U[[3]].cpu().numpy(), want),
                                       "unitarity_err": float((U.conj().transpose(-1, -2) @ U - eye).abs().max())}
for k, v in out.items(): print(k, v, flush=True)
json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)
