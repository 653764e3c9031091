import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from c3_b200 import engine
from oracle import c3_signal_oracle as so
TP = 2 * np.pi
vf = np.array([3.77798058e-01, 9.01503403e-09, 1.89419234e-09, 1.07508920e+00, -2.74129318e+08, 1.48878187e+00, 1.70749557e-09, 7.51167006e-09, 7.03252836e-10])
vg = np.array([0.45, 8.8e-9, 2.1e-9, -0.7, 3.1e8, -1.3, 1e-9, 7e-9, 1e-9])
def run(name, envs, B=1):
    E = len(envs)
    env = np.stack([v for (_, _, v) in envs]).reshape(1, 1, E, 9).repeat(B, 0).copy()
    sid = np.array([[ {"flattop":5,"gaussian_nonorm":2}[s] for (s,_,_) in envs]], dtype=np.int32)
    flags = np.array([[f for (_, f, _) in envs]], dtype=np.int32)
    lo = np.full((B, 1), 5e9 * TP); chain = np.array([[100e9, 1.7e9, 0.37e-9, 0, 0, 1e9, 0, 1, 0, 0, np.nan]])
    N = engine.signal_slice_num(0.0, 9.7e-9, 100e9)
    rng = np.random.default_rng(0); w = rng.normal(size=(B, 1, N))
    genv, glo, gv = engine.generate_signals_grad(env, sid, flags, lo, chain, 0.0, 9.7e-9, w)
    genv = genv.cpu().numpy()
    def loss(env_):
        tot = 0.0
        for b in range(B):
            specs = [so.EnvelopeSpec(shape=s, amp=vv[0], t_final=vv[1], sigma=vv[2], xy_angle=vv[3], freq_offset=vv[4], delta=vv[5], t_up=vv[6], t_down=vv[7], risefall=vv[8], drag=bool(f & 1), use_t_before=bool(f & 2)) for (s, f, _), vv in zip(envs, env_[b, 0])]
            st = {}; so.generate_signal(specs, lo[b, 0], 0.0, 9.7e-9, so.ChainSpec(sim_res=100e9, awg_res=1.7e9, rise_time=0.37e-9, v2hz=1e9), st)
            tot += float(np.sum(w[b, 0] * so.mixer(st["lo_i"], st["lo_q"], st["dac_i"], st["dac_q"]) * 1e9))
        return tot
    worst = 0
    for b in range(B):
        for e in range(E):
            for q in range(9):
                h = abs(env[b, 0, e, q]) * 1e-6
                ep, em = env.copy(), env.copy(); ep[b, 0, e, q] += h; em[b, 0, e, q] -= h
                want = (loss(ep) - loss(em)) / (2 * h)
                if want != 0 or genv[b, 0, e, q] != 0:
                    worst = max(worst, abs(genv[b, 0, e, q] - want) / max(abs(want), 1e-300))
    print(name, "worst rel", worst)
run("gauss drag", [("gaussian_nonorm", 1, vg)])
run("gauss plain", [("gaussian_nonorm", 0, vg)])
run("gauss drag + flattop tb", [("gaussian_nonorm", 1, vg), ("flattop", 2, vf)])
run("gauss + flattop", [("gaussian_nonorm", 0, vg), ("flattop", 0, vf)])
run("B=2 flattop", [("flattop", 0, vf)], B=2)
