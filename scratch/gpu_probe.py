"""First GPU probe: fp64 peaks, mm_row primitive ceiling, first timings."""
import json, time, sys
import numpy as np, torch
sys.path.insert(0, ".")
from c3_b200 import engine, synth, _lib
lib = _lib.load()
out = {}
out["gpu"] = torch.cuda.get_device_name(0)
out["dfma_tflops"] = engine.measure_fp64_peak("dfma", 1.0)
out["dmma_tflops"] = engine.measure_fp64_peak("dmma", 1.0)
print(out, flush=True)
mb = {}
for kind, name in [(0, "D9"), (1, "D4"), (2, "D3")]:
    for warps, ctas in [(4, 1), (4, 2), (4, 3), (4, 4), (8, 1), (8, 2), (16, 1), (2, 8)]:
        mb[f"{name}_w{warps}_c{ctas}"] = round(lib.c3b_microbench(kind, warps, ctas), 2)
out["mmrow_pipe_tflops"] = mb
print(mb, flush=True)

def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

m = synth.two_transmon()
for B in (256, 4096):
    sig = torch.as_tensor(synth.controls_fast(m, B, 1000)).cuda()
    h0 = torch.as_tensor(m.h0).cuda(); hks = torch.as_tensor(m.hks).cuda()
    for tu in (4096, 16384, 32768, 131072):
        engine.set_tuning("target_units", tu)
        ms = timeit(lambda: engine.pwc_closed(h0, hks, sig, 1e-11))
        out[f"d9_B{B}_tu{tu}_ms"] = ms
        out[f"d9_B{B}_tu{tu}_slices_per_s"] = B * 1000 / (ms * 1e-3)
        print(B, tu, ms, B * 1000 / (ms * 1e-3), flush=True)
    engine.set_tuning("target_units", 32768)
m1 = synth.one_qubit()
sig = torch.as_tensor(synth.controls_fast(m1, 4096, 800)).cuda()
ms = timeit(lambda: engine.pwc_closed(torch.as_tensor(m1.h0).cuda(), torch.as_tensor(m1.hks).cuda(), sig, 1e-11))
out["d3_B4096_N800_slices_per_s"] = 4096 * 800 / (ms * 1e-3)
m27 = synth.tunable_coupler()
sig = torch.as_tensor(synth.controls_fast(m27, 64, 200)).cuda()
ms = timeit(lambda: engine.pwc_closed(torch.as_tensor(m27.h0).cuda(), torch.as_tensor(m27.hks).cuda(), sig, 1e-11), 2)
out["d27_B64_N200_slices_per_s"] = 64 * 200 / (ms * 1e-3)
sig = torch.as_tensor(synth.controls_fast(m, 16, 50)).cuda()
ms = timeit(lambda: engine.pwc_lindblad(m.h0, m.hks, m.col_ops, sig, 1e-11), 2)
out["lind81_B16_N50_slices_per_s"] = 16 * 50 / (ms * 1e-3)
print(json.dumps(out, indent=1))
json.dump(out, open("gpurun_out/probe1.json", "w"), indent=1)
