import sys; sys.path.insert(0,"."); sys.path.insert(0,"tests")
"""GPU parity of the on-device signal chain (SURVEY.md section 8f, f-2) against the reference's fixtures and the
CPU oracle (oracle/c3_signal_oracle.py).  Control fields are O(1e9) rad/s; tolerance is relative to the largest
sample: 1e-12 (sincos/exp/erf of the device maths library differ from numpy's by <= 2 ulp)."""
import numpy as np
import pytest
import torch

import c3_fakes as fk
from oracle import c3_oracle as orc
from oracle import c3_signal_oracle as so

pytestmark = pytest.mark.gpu
TP = 2 * np.pi
RTOL = 1e-12



from c3_b200 import generator as gen_mod_
def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max()


def _oracle_signals(env, sid, flags, lo, chain, shapes, t_start, t_end, resp_kind):
    """Oracle control fields [B,K,N] for the flat parameter tables the C ABI takes."""
    B, K, E, _ = env.shape
    out = []
    for b in range(B):
        rows = []
        for k in range(K):
            specs = []
            for e in range(E):
                if sid[k, e] < 0:
                    continue
                v = env[b, k, e]
                specs.append(so.EnvelopeSpec(shape=shapes[k][e], amp=v[0], t_final=v[1], sigma=v[2], xy_angle=v[3],
                                             freq_offset=v[4], delta=v[5], t_up=v[6], t_down=v[7], risefall=v[8],
                                             drag=bool(flags[k][e] & 1), use_t_before=bool(flags[k][e] & 2)))
            c = chain[k] if chain.ndim == 2 else chain[b, k]
            cs = so.ChainSpec(sim_res=c[0], awg_res=c[1], rise_time=c[2], response_fft=(resp_kind == 2), v2hz=c[5],
                              flux=None if c[4] == 0 else dict(phi=c[6], phi_0=c[7], omega_0=c[8], anhar=c[9],
                                                               d=None if np.isnan(c[10]) else c[10]))
            if resp_kind == 0:
                st = {}
                so.generate_signal(specs, lo[b, k], t_start, t_end, cs, st)
                mixed = so.mixer(st["lo_i"], st["lo_q"], st["dac_i"], st["dac_q"])
                rows.append(mixed * cs.v2hz if cs.flux is None else so.flux_tuning(mixed, **cs.flux))
            else:
                rows.append(so.generate_signal(specs, lo[b, k], t_start, t_end, cs)[0])
        out.append(np.stack(rows))
    return np.stack(out)


@pytest.mark.parametrize("resp_kind", [0, 1, 2])
def test_signal_chain_gradient_vs_oracle_finite_differences(gen_mod, resp_kind):
    """Reverse mode of the chain (c3b_generate_signals_grad) against central finite differences of the ORACLE for
    every envelope parameter, the carrier frequency and V_to_Hz; L = sum w * signals with random weights."""
    from c3_b200 import engine
    rng = np.random.default_rng(5 + resp_kind)
    B, K, E = 2, 3, 2
    t_start, t_end = 0.0, 9.7e-9
    shapes = [["gaussian_nonorm", "flattop"], ["cosine", "gaussian_sigma"], ["gaussian_nonorm", None]]
    flags = np.array([[1, 2], [1 | 2, 0], [1, 0]], dtype=np.int32)
    env = np.zeros((B, K, E, 9))
    sid = -np.ones((K, E), dtype=np.int32)
    for k in range(K):
        for e in range(E):
            if shapes[k][e] is None:
                continue
            sid[k, e] = gen_mod.SHAPE_IDS[shapes[k][e]]
            tf_ = rng.uniform(8e-9, 9.5e-9, B)
            env[:, k, e] = np.stack([rng.uniform(0.1, 0.6, B), tf_, tf_ / rng.uniform(3, 5, B), rng.uniform(-3, 3, B),
                                     rng.uniform(-80e6, 80e6, B) * TP, rng.uniform(-2, 2, B), rng.uniform(1e-9, 2e-9, B),
                                     tf_ - rng.uniform(1e-9, 2e-9, B), rng.uniform(0.5e-9, 1.5e-9, B)], axis=1)
    lo = rng.uniform(4e9, 6e9, (B, K)) * TP
    chain = np.zeros((K, 11))
    for k in range(K):
        chain[k] = [100e9, 1.7e9, 0.37e-9, resp_kind, 0, 1e9 * (1 + 0.1 * k), 0, 1, 0, 0, np.nan]
    chain[2, 4:] = [1, 0, 2.3, 10.0, 8.1e9 * TP, -286e6 * TP, 0.36 if resp_kind else np.nan]
    N = engine.signal_slice_num(t_start, t_end, 100e9)      # int(9.7e-9 * 100e9) = 969 in floating point
    w = rng.normal(size=(B, K, N))
    genv, glo, gv = engine.generate_signals_grad(env, sid, flags, lo, chain, t_start, t_end, w)
    genv, glo, gv = genv.cpu().numpy(), glo.cpu().numpy(), gv.cpu().numpy()

    def loss(env_, lo_, chain_):
        return float(np.sum(w * _oracle_signals(env_, sid, flags, lo_, chain_, shapes, t_start, t_end, resp_kind)))

    def fd(x, index, rel):
        h = abs(x[index]) * rel if x[index] != 0 else rel
        xp, xm = x.copy(), x.copy()
        xp[index] += h
        xm[index] -= h
        return xp, xm, 2 * h

    scale = np.abs(genv).max()
    checked = 0
    for b in range(B):
        for k in range(K):
            for e in range(E):
                if sid[k, e] < 0:
                    assert np.all(genv[b, k, e] == 0)
                    continue
                for q in range(9):
                    xp, xm, h2 = fd(env, (b, k, e, q), 1e-6)
                    want = (loss(xp, lo, chain) - loss(xm, lo, chain)) / h2
                    tol = 2e-5 * max(abs(want), abs(genv[b, k, e, q])) + 1e-9 * scale
                    print(b,k,e,q, f"{genv[b,k,e,q]:.6e} {want:.6e} rel {abs(genv[b,k,e,q]-want)/max(abs(want),1e-300):.1e}", "BAD" if abs(genv[b, k, e, q] - want) >= tol else "")
                    checked += 1
            xp, xm, h2 = fd(lo, (b, k), 1e-9)
            want = (loss(env, xp, chain) - loss(env, xm, chain)) / h2
            print("lo", b, k, glo[b,k], want)
    for k in range(K):
        if chain[k, 4] == 0:
            xp, xm, h2 = fd(chain, (k, 5), 1e-6)
            want = (loss(env, lo, xp) - loss(env, lo, xm)) / h2
            assert abs(gv[:, k].sum() - want) < 1e-6 * abs(want)
        else:
            assert np.all(gv[:, k] == 0)
    assert checked == 2 * 5 * 9



test_signal_chain_gradient_vs_oracle_finite_differences(gen_mod_, 0)
