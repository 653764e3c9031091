"""GPU parity of the on-device signal chain against the reference's fixtures and the
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


@pytest.fixture(scope="module")
def gen_mod():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from c3_b200 import generator
    return generator


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max()


def test_reference_full_signal_chain(gen_mod, golden_generator):
    """test/test_generator.py:183-190 of the reference, through Generator.generate_signals on the device."""
    devices, chains, instr = fk.reference_generator_setup()
    out = gen_mod.Generator(devices, chains).generate_signals(instr)
    assert set(out) == {"d1"} and set(out["d1"]) == {"values", "ts"}
    assert np.array_equal(out["d1"]["ts"].cpu().numpy(), golden_generator["full_ts"])
    assert _rel(out["d1"]["values"].cpu().numpy(), golden_generator["full_values"]) < RTOL


def test_reference_tunable_coupler_flux_line(gen_mod, golden_tunable_coupler):
    """Flux line of test/test_tunable_coupler.py (flattop, 2.4 GS/s AWG, FluxTuning): the pickled 10 000-sample
    field, then the d = 27 propagators it drives against the pickled partial propagators."""
    g = golden_tunable_coupler
    devices, chains, instr = fk.tunable_coupler_flux_setup()
    out = gen_mod.Generator(devices, chains).generate_signals(instr)
    sig = out["TC"]["values"]
    assert _rel(sig.cpu().numpy(), g["tc_signal"]) < RTOL
    from c3_b200 import engine
    dt = float(out["TC"]["ts"][1] - out["TC"]["ts"][0])
    _, dUs = engine.pwc_closed(g["h0"], g["hk_tc"][None], sig[None, None, :], dt, return_dUs=True)
    got = dUs[0].cpu().numpy()[g["dUs_index"]]
    assert np.linalg.norm(got - g["dUs"]) / np.linalg.norm(g["dUs"]) < 1e-10


@pytest.mark.parametrize("resp_kind", [0, 1, 2])
def test_random_batches_all_shapes(gen_mod, resp_kind):
    """Every envelope shape, DRAG quadrature, use_t_before, two envelopes on one line, per-sample parameters,
    both Response variants and none, VoltsToHertz and FluxTuning outputs; odd resampling ratios."""
    from c3_b200 import engine
    rng = np.random.default_rng(17 + resp_kind)
    B, K, E = 5, 3, 2
    t_start, t_end = 0.0, 12.3e-9
    sim_res, awg_res = 100e9, 1.7e9
    shapes = [["gaussian_nonorm", "flattop"], ["cosine", "gaussian_sigma"], ["rect", None]]
    flags = [[1, 2], [1 | 2, 0], [0, 0]]
    env = np.zeros((B, K, E, 9))
    sid = -np.ones((K, E), dtype=np.int32)
    for k in range(K):
        for e in range(E):
            if shapes[k][e] is None:
                continue
            sid[k, e] = gen_mod.SHAPE_IDS[shapes[k][e]]
            tf_ = rng.uniform(8e-9, 12e-9, B)
            env[:, k, e] = np.stack([rng.uniform(0.1, 0.6, B), tf_, tf_ / rng.uniform(3, 6, B), rng.uniform(-3, 3, B),
                                     rng.uniform(-80e6, 80e6, B) * TP, rng.uniform(-2, 2, B), rng.uniform(1e-9, 2e-9, B),
                                     tf_ - rng.uniform(1e-9, 2e-9, B), rng.uniform(0.5e-9, 1.5e-9, B)], axis=1)
    lo = rng.uniform(4e9, 6e9, (B, K)) * TP
    chain = np.zeros((K, 11))
    for k in range(K):
        chain[k] = [sim_res, awg_res, 0.37e-9, resp_kind, 0, 1e9 * (1 + 0.1 * k), 0, 1, 0, 0, np.nan]
    chain[2, 4:] = [1, 0, 2.3, 10.0, 8.1e9 * TP, -286e6 * TP, 0.36 if resp_kind else np.nan]
    got = engine.generate_signals(env, sid, np.array(flags, dtype=np.int32), lo, chain, t_start, t_end).cpu().numpy()
    assert got.shape == (B, K, 1230)
    for b in range(B):
        for k in range(K):
            specs = []
            for e in range(E):
                if sid[k, e] < 0:
                    continue
                v = env[b, k, e]
                specs.append(so.EnvelopeSpec(shape=shapes[k][e], amp=v[0], t_final=v[1], sigma=v[2], xy_angle=v[3],
                                             freq_offset=v[4], delta=v[5], t_up=v[6], t_down=v[7], risefall=v[8],
                                             drag=bool(flags[k][e] & 1), use_t_before=bool(flags[k][e] & 2)))
            c = chain[k]
            cs = so.ChainSpec(sim_res=c[0], awg_res=c[1], rise_time=c[2], response_fft=(resp_kind == 2), v2hz=c[5],
                              flux=None if c[4] == 0 else dict(phi=c[6], phi_0=c[7], omega_0=c[8], anhar=c[9],
                                                               d=None if np.isnan(c[10]) else c[10]))
            if resp_kind == 0:
                st = {}
                so.generate_signal(specs, lo[b, k], t_start, t_end, cs, st)
                mixed = so.mixer(st["lo_i"], st["lo_q"], st["dac_i"], st["dac_q"])
                want = mixed * cs.v2hz if cs.flux is None else so.flux_tuning(mixed, **cs.flux)
            else:
                want, _ = so.generate_signal(specs, lo[b, k], t_start, t_end, cs)
            assert _rel(got[b, k], want) < RTOL, (b, k)


def test_parameters_to_propagators_on_device(gen_mod):
    """Batched pulse parameters -> control fields -> propagators without leaving the device, against
    oracle chain + oracle propagators (what B serial generate_signals + pwc calls of the reference compute)."""
    from c3_b200 import propagation as prop, synth
    m = synth.two_transmon()
    devices, chains, instr = fk.reference_generator_setup()
    instr2 = fk.Instruction("rx90p", 0.0, 7e-9, ["d1", "d2"])
    for c, f in (("d1", 5.0e9), ("d2", 5.6e9)):
        instr2.add_component(fk.EnvelopeDrag("gauss", "gaussian_nonorm", {
            "amp": fk.Quantity(0.4, "V"), "t_final": fk.Quantity(7e-9, "s"), "sigma": fk.Quantity(7e-9 / 4, "s"),
            "xy_angle": fk.Quantity(0.1, "rad"), "freq_offset": fk.Quantity(-53e6, "Hz 2pi"), "delta": fk.Quantity(-1, "")}), c)
        instr2.add_component(fk.Carrier("carrier", {"freq": fk.Quantity(f + 50e6, "Hz 2pi")}), c)
    chains2 = {"d1": fk.standard_chain(), "d2": fk.standard_chain()}
    gen = gen_mod.Generator(devices, chains2)
    amps = np.linspace(0.2, 0.5, 6)
    sig, ts = gen.generate_signals_batch(instr2, {("d1", "gauss", "amp"): amps, ("d2", "gauss", "amp"): 0 * amps})
    assert sig.is_cuda and tuple(sig.shape) == (6, 2, 700)
    dt = float(ts[1] - ts[0])
    U = prop.pwc_batch(m.h0, m.hks, sig, dt)
    for b in range(6):
        rows = []
        for c, f, a in (("d1", 5.0e9, amps[b]), ("d2", 5.6e9, 0.0)):
            e = so.EnvelopeSpec(shape="gaussian_nonorm", amp=a, t_final=7e-9, sigma=7e-9 / 4, xy_angle=0.1,
                                freq_offset=-53e6 * TP, delta=-1, drag=True)
            rows.append(so.generate_signal([e], (f + 50e6) * TP, 0.0, 7e-9, so.ChainSpec())[0])
        want_sig = np.stack(rows)
        assert _rel(sig[b].cpu().numpy(), want_sig) < RTOL
        want_U = orc.propagate_batch(m.h0, m.hks, want_sig[None], dt)[0]
        assert np.linalg.norm(U[b].cpu().numpy() - want_U) / np.linalg.norm(want_U) < 1e-10


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
            # fixed pulse lengths: a random t_final can put an AWG sample inside the (dt * 1e-6 wide) edge of the
            # reference's sigmoid mask, where a finite difference sees a jump
            tf_ = 8.3e-9 + 0.23e-9 * (np.arange(B) + 2 * e + 0.5 * k)
            env[:, k, e] = np.stack([rng.uniform(0.1, 0.6, B), tf_, tf_ / rng.uniform(3, 5, B), rng.uniform(-3, 3, B),
                                     rng.uniform(-80e6, 80e6, B) * TP, rng.uniform(-2, 2, B), rng.uniform(1e-9, 2e-9, B),
                                     tf_ - rng.uniform(1e-9, 2e-9, B), rng.uniform(0.5e-9, 1.5e-9, B)], axis=1)
    ts_awg = so.create_ts(t_start, t_end, 1.7e9)
    assert np.abs(ts_awg[None, :] - 0.999 * env[..., 1].reshape(-1, 1)).min() > 1e-12
    env[:, 1, 1, 0] *= 2e-9          # gaussian_sigma has unit AREA (values ~ 1/sigma): keep the line O(1) V
    lo = rng.uniform(4e9, 6e9, (B, K)) * TP
    chain = np.zeros((K, 11))
    for k in range(K):
        chain[k] = [100e9, 1.7e9, 0.37e-9, resp_kind, 0, 1e9 * (1 + 0.1 * k), 0, 1, 0, 0, np.nan]
    chain[2, 4:] = [1, 0, 2.3, 10.0, 8.1e9 * TP, -286e6 * TP, 0.36 if resp_kind else np.nan]
    N = engine.signal_slice_num(t_start, t_end, 100e9)      # int(9.7e-9 * 100e9) = 969 in floating point
    w = rng.normal(size=(B, K, N))
    genv, glo, gv = engine.generate_signals_grad(env, sid, flags, lo, chain, t_start, t_end, w)
    genv, glo, gv = genv.cpu().numpy(), glo.cpu().numpy(), gv.cpu().numpy()

    def loss(env_, lo_, chain_, b=None, k=None):
        """sum w * signals restricted to one (sample, line) when given: the finite difference of a parameter of
        that line is then not drowned in the rounding noise of the other lines' sums"""
        if b is None:
            return float(np.sum(w * _oracle_signals(env_, sid, flags, lo_, chain_, shapes, t_start, t_end, resp_kind)))
        sig = _oracle_signals(env_[b:b + 1, k:k + 1], sid[k:k + 1], flags[k:k + 1], lo_[b:b + 1, k:k + 1],
                              chain_[k:k + 1], shapes[k:k + 1], t_start, t_end, resp_kind)
        return float(np.sum(w[b, k] * sig[0, 0]))

    def fd(x, index, rel):
        h = abs(x[index]) * rel if x[index] != 0 else rel
        xp, xm = x.copy(), x.copy()
        xp[index] += h
        xm[index] -= h
        return xp, xm, 2 * h

    checked = 0
    for b in range(B):
        for k in range(K):
            for e in range(E):
                if sid[k, e] < 0:
                    assert np.all(genv[b, k, e] == 0)
                    continue
                for q in range(9):
                    xp, xm, h2 = fd(env, (b, k, e, q), 1e-6)
                    want = (loss(xp, lo, chain, b, k) - loss(xm, lo, chain, b, k)) / h2
                    tol = 2e-5 * max(abs(want), abs(genv[b, k, e, q])) + 1e-7 * np.abs(genv[b, k, :, q]).max()
                    assert abs(genv[b, k, e, q] - want) <= tol, (b, k, e, q, genv[b, k, e, q], want)
                    checked += 1
            xp, xm, h2 = fd(lo, (b, k), 1e-9)
            want = (loss(env, xp, chain, b, k) - loss(env, xm, chain, b, k)) / h2
            assert abs(glo[b, k] - want) < 1e-4 * abs(want) + 1e-6 * np.abs(glo).max(), (b, k, glo[b, k], want)
    for k in range(K):
        if chain[k, 4] == 0:
            xp, xm, h2 = fd(chain, (k, 5), 1e-6)
            want = (loss(env, lo, xp) - loss(env, lo, xm)) / h2
            assert abs(gv[:, k].sum() - want) < 1e-6 * abs(want)
        else:
            assert np.all(gv[:, k] == 0)
    assert checked == 2 * 5 * 9


def test_pulse_parameter_gradient_through_the_whole_pipeline(gen_mod):
    """parameters -> fields (f-2) -> propagators (a1-a11) -> infidelity (f-3) -> backward (f-3, f-1, f-2) on the
    device: dL/d(pulse parameters) against central finite differences of the same device pipeline."""
    from c3_b200 import engine, fidelities as fid, propagation as prop, synth
    from oracle import c3_fid_oracle as fo
    m = synth.two_transmon()
    B, K = 3, 2
    T, N = 7e-9, 700
    rng = np.random.default_rng(9)
    env = np.zeros((B, K, 1, 9))
    env[..., 0, 0] = rng.uniform(0.3, 0.5, (B, K))
    env[..., 0, 1] = T
    env[..., 0, 2] = T / 4
    env[..., 0, 3] = rng.uniform(0, 1, (B, K))
    env[..., 0, 4] = -53e6 * TP
    env[..., 0, 5] = -1.0
    env[..., 0, 8] = 1.0
    lo = np.broadcast_to(np.array([5.05e9, 5.65e9]) * TP, (B, K)).copy()
    sid = np.full((K, 1), 2, dtype=np.int32)
    flags = np.ones((K, 1), dtype=np.int32)
    chain = np.tile([100e9, 2e9, 0.3e-9, 1, 0, 1e9, 0, 1, 0, 0, np.nan], (K, 1))
    G = np.kron(fo.GATES["rx90p"], fo.GATES["id"])

    def pipeline(env_t, lo_t):
        sig = engine.generate_signals_autograd(env_t, lo_t, sid, flags, chain, 0.0, T)
        U = prop.pwc_batch_autograd(m.h0, m.hks, sig, 1e-11)
        return fid.unitary_infid_autograd(G, U, [0, 1], [3, 3]).mean()

    env_t = torch.tensor(env, device="cuda", requires_grad=True)
    lo_t = torch.tensor(lo, device="cuda", requires_grad=True)
    L = pipeline(env_t, lo_t)
    L.backward()
    g = env_t.grad.cpu().numpy()
    for (b, k, q, rel) in [(0, 0, 0, 1e-5), (1, 1, 0, 1e-5), (2, 0, 3, 1e-5), (0, 1, 5, 1e-4), (1, 0, 4, 1e-6), (2, 1, 2, 1e-5)]:
        h = abs(env[b, k, 0, q]) * rel
        ep, em = env.copy(), env.copy()
        ep[b, k, 0, q] += h
        em[b, k, 0, q] -= h
        with torch.no_grad():
            lp = float(pipeline(torch.tensor(ep, device="cuda"), torch.tensor(lo, device="cuda")))
            lm = float(pipeline(torch.tensor(em, device="cuda"), torch.tensor(lo, device="cuda")))
        want = (lp - lm) / (2 * h)
        assert abs(g[b, k, 0, q] - want) < 2e-4 * abs(want) + 1e-9 * np.abs(g[..., q]).max(), (b, k, q, g[b, k, 0, q], want)
    assert lo_t.grad is not None and torch.isfinite(lo_t.grad).all()


@pytest.mark.parametrize("resp_kind", [0, 1])
def test_round2_envelope_shapes(gen_mod, resp_kind):
    """trapezoid, flattop_risefall, gaussian_der(_nonorm), drag_sigma, drag_der: values, DRAG quadrature (analytic time derivative),
    use_t_before, against the oracle (whose shape functions are pinned to the reference's test/envelopes.pickle)."""
    from c3_b200 import engine
    rng = np.random.default_rng(40 + resp_kind)
    shapes = ["trapezoid", "flattop_risefall", "gaussian_der_nonorm", "gaussian_der", "drag_sigma", "drag_der"]
    B, K, E = 3, len(shapes), 1
    t_start, t_end = 0.0, 11.1e-9
    env = np.zeros((B, K, E, 9))
    sid = np.array([[gen_mod.SHAPE_IDS[s_]] for s_ in shapes], dtype=np.int32)
    flags = np.array([[1], [1 | 2], [0], [1], [1], [2]], dtype=np.int32)
    for k in range(K):
        tf_ = 9.13e-9 + 0.21e-9 * np.arange(B) + 0.07e-9 * k
        env[:, k, 0] = np.stack([rng.uniform(0.1, 0.6, B), tf_, tf_ / rng.uniform(3, 5, B), rng.uniform(-3, 3, B),
                                 rng.uniform(-80e6, 80e6, B) * TP, rng.uniform(-2, 2, B), rng.uniform(1e-9, 2e-9, B),
                                 tf_ - rng.uniform(1e-9, 2e-9, B), rng.uniform(0.5e-9, 1.2e-9, B)], axis=1)
    env[:, 2, 0, 0] *= 1e-9          # gaussian_der_nonorm ~ 1 / sigma: keep the line O(1) V
    env[:, 3:, 0, 0] *= 2e-9         # unit-area normalisations (values ~ 1 / sigma)
    env[:, 5, 0, 0] *= 2e-9          # drag_der ~ 1 / sigma^2
    lo = rng.uniform(4e9, 6e9, (B, K)) * TP
    chain = np.tile([100e9, 1.9e9, 0.37e-9, resp_kind, 0, 1e9, 0, 1, 0, 0, np.nan], (K, 1))
    got = engine.generate_signals(env, sid, flags, lo, chain, t_start, t_end).cpu().numpy()
    for b in range(B):
        for k in range(K):
            v = env[b, k, 0]
            spec = so.EnvelopeSpec(shape=shapes[k], amp=v[0], t_final=v[1], sigma=v[2], xy_angle=v[3], freq_offset=v[4], delta=v[5],
                                   t_up=v[6], t_down=v[7], risefall=v[8], drag=bool(flags[k, 0] & 1), use_t_before=bool(flags[k, 0] & 2))
            cs = so.ChainSpec(sim_res=100e9, awg_res=1.9e9, rise_time=0.37e-9)
            if resp_kind == 0:
                st = {}
                so.generate_signal([spec], lo[b, k], t_start, t_end, cs, st)
                want = so.mixer(st["lo_i"], st["lo_q"], st["dac_i"], st["dac_q"]) * cs.v2hz
            else:
                want, _ = so.generate_signal([spec], lo[b, k], t_start, t_end, cs)
            # the oracle differentiates the shape numerically for the DRAG quadrature (1e-9 relative), the kernel analytically
            assert _rel(got[b, k], want) < (1e-7 if flags[k, 0] & 1 else RTOL), (b, shapes[k])


def test_alias_shapes_through_the_generator(gen_mod):
    """gaussian (sigma = t_final / 6), drag (sigma = t_final / 4), flattop_risefall_1ns (risefall = 1 ns): envelopes.py:366-370,
    399-417, 533-542, resolved per sample on the host and run on the kernel of the shape they alias."""
    devices, chains, instr = fk.reference_generator_setup()
    gen = gen_mod.Generator(devices, chains)
    for shape in ("gaussian", "drag", "flattop_risefall_1ns"):
        env = instr.comps["d1"]["gauss"]
        env.shape = fk._Shape(shape)
        env.params["risefall"] = fk.Quantity(0.7e-9, "s")
        amp = 0.5 if shape == "flattop_risefall_1ns" else 0.5 * 2e-9
        env.params["amp"] = fk.Quantity(amp, "V")
        got = gen.generate_signals(instr)["d1"]["values"].cpu().numpy()
        spec = so.EnvelopeSpec(shape=shape, amp=amp, t_final=7e-9, sigma=7e-9 / 4, xy_angle=0.0,
                               freq_offset=(-50e6 - 3e6) * TP, delta=-1, risefall=0.7e-9)
        want, _ = so.generate_signal([spec], (5e9 + 50e6) * TP, 0.0, 7e-9, so.ChainSpec())
        assert _rel(got, want) < RTOL, shape


# shape name -> (scalar envelope parameters as EnvelopeSpec fields, extra parameters); t_final = 7 ns, 14 AWG samples at 2 GS/s
_PWC_X = dict(t_bin_start=0.2e-9, t_bin_end=6.7e-9, inphase=[0.0, 0.1, 0.3, 0.5, 0.1, 1.1, 0.4, 0.1])
EXTENDED_SHAPES = {
    "flattop_cut": (dict(t_up=1.1e-9, t_down=5.9e-9, risefall=0.8e-9), {}),
    "flattop_cut_center": (dict(risefall=0.9e-9), dict(width=4.4e-9)),
    "flattop_variant": (dict(t_up=0.6e-9, t_down=6.6e-9), dict(ramp=1.7e-9)),
    "flattop_variant/wide-ramp": (dict(t_up=0.6e-9, t_down=6.6e-9), dict(ramp=4e-9)),
    "cosine_flattop": ({}, dict(t_rise=2.1e-9)),
    "delta_pulse": ({}, dict(t_sig=[0.5e-9, 4.2e-9])),
    "pwc": ({}, dict(inphase=np.linspace(-0.4, 0.9, 14) ** 2, quadrature=np.cos(np.arange(14.0)))),
    "pwc_shape": ({}, _PWC_X),
    "pwc_symmetric": ({}, _PWC_X),
    "pwc_shape_plateau": ({}, _PWC_X),
    "pwc_shape_plateau/width": ({}, dict(_PWC_X, t_bin_start=0.0, t_bin_end=3e-9, width=5.5e-9)),
    "fourier_sin": ({}, dict(amps=[0.5, 0.2, -0.1], freqs=[1e6, 1e9, 3.3e9], phases=[0.0, 1.0, -0.4])),
    "fourier_cos": ({}, dict(amps=[0.5, 0.2], freqs=[1e6, 2.5e9])),
    "slepian_fourier": ({}, dict(width=6e-9, fourier_coeffs=[1, 0.5, 0.2], offset=0.1)),
    "slepian_fourier/risefall": ({}, dict(width=6e-9, fourier_coeffs=[1, 0.5, 0.2], offset=0.1, risefall=1.5e-9)),
    "slepian_fourier/sin": ({}, dict(width=6e-9, fourier_coeffs=[1, 0.5, 0.2], offset=0.1, risefall=1.5e-9, sin_coeffs=[0.3, -0.1])),
}


@pytest.mark.parametrize("case", sorted(EXTENDED_SHAPES))
def test_extended_envelope_shapes_through_the_generator(gen_mod, case):
    """The array-parametrised and grid-defined shapes of c3/libraries/envelopes.py (pwc*, delta_pulse, fourier_*, slepian_fourier,
    flattop_cut*, flattop_variant, cosine_flattop) through Generator.generate_signals on the device, against the oracle whose
    shape functions are pinned to the reference's test/envelopes.pickle; a second envelope of a closed-form shape shares the
    line so that the per-envelope accumulation and the table rows of a mixed instruction are exercised."""
    shape = case.split("/")[0]
    fields, extra = EXTENDED_SHAPES[case]
    devices, chains, instr = fk.reference_generator_setup()
    gen = gen_mod.Generator(devices, chains)
    params = {"amp": fk.Quantity(0.45, "V"), "t_final": fk.Quantity(7e-9, "s"), "xy_angle": fk.Quantity(0.3, "rad"),
              "freq_offset": fk.Quantity(-53e6, "Hz 2pi"), "delta": fk.Quantity(0.0, "")}
    for key, val in {**fields, **extra}.items():
        params[key] = fk.Quantity(val, "")
    instr.add_component(fk.Envelope("ext", shape, params), "d1")
    got = gen.generate_signals(instr)["d1"]["values"].cpu().numpy()
    first = so.EnvelopeSpec(shape="gaussian_nonorm", amp=0.5, t_final=7e-9, sigma=7e-9 / 4, xy_angle=0.0,
                            freq_offset=(-50e6 - 3e6) * TP, delta=-1)
    spec = so.EnvelopeSpec(shape=shape, amp=0.45, t_final=7e-9, xy_angle=0.3, freq_offset=-53e6 * TP, **fields, extra=dict(extra))
    want, _ = so.generate_signal([first, spec], (5e9 + 50e6) * TP, 0.0, 7e-9, so.ChainSpec())
    assert np.abs(want).max() > 1e6          # the line carries a signal (Hz)
    assert _rel(got, want) < RTOL, case
    # the closed-form envelope alone gives a different line: the extended envelope really contributed
    alone, _ = so.generate_signal([first], (5e9 + 50e6) * TP, 0.0, 7e-9, so.ChainSpec())
    assert _rel(alone, want) > 1e-3


def test_extended_shapes_are_forward_only(gen_mod):
    """DRAG / use_t_before on an extended shape is refused by the host, and the reverse mode marks their rows with NaN."""
    from c3_b200 import engine
    devices, chains, instr = fk.reference_generator_setup()
    gen = gen_mod.Generator(devices, chains)
    instr.add_component(fk.EnvelopeDrag("ext", "fourier_cos", {"amp": fk.Quantity(0.1, "V"), "t_final": fk.Quantity(7e-9, "s"),
                                                               "amps": fk.Quantity([1.0], ""), "freqs": fk.Quantity([1e9], "")}), "d1")
    with pytest.raises(Exception, match="without DRAG"):
        gen.generate_signals(instr)
    env = np.zeros((1, 1, 1, 9)); env[0, 0, 0, :2] = [0.3, 7e-9]; env[0, 0, 0, 2] = 1e-9; env[0, 0, 0, 8] = 1e-9
    sid = np.array([[gen_mod.SHAPE_IDS["flattop_cut"]]], dtype=np.int32)
    chain = np.tile([100e9, 2e9, 0.3e-9, 1, 0, 1e9, 0, 1, 0, 0, np.nan], (1, 1))
    N = engine.signal_slice_num(0.0, 7e-9, 100e9)
    genv, glo, _ = engine.generate_signals_grad(env, sid, np.zeros((1, 1), np.int32), np.full((1, 1), 5e9 * TP), chain, 0.0, 7e-9,
                                                torch.ones((1, 1, N), dtype=torch.float64, device="cuda"))
    assert torch.isnan(genv).all()


def test_crosstalk_device(gen_mod):
    """The Crosstalk post-processing of Generator.generate_signals (c3/generator/generator.py:229-234, devices.py:281-293) on the
    device: the reference's known answers on raw lines, and a three-line instruction whose first and third lines are crossed."""
    from c3_b200 import engine
    raw = torch.tensor(np.stack([np.linspace(0, 100, 101), np.ones(101), np.linspace(100, 200, 101)])[None], device="cuda")
    flip = engine.crosstalk(raw.clone(), [0, 2], [[0, 1], [1, 0]]).cpu().numpy()[0]
    assert (flip[2] == np.linspace(0, 100, 101)).all() and (flip[0] == np.linspace(100, 200, 101)).all() and (flip[1] == 1).all()
    mix = engine.crosstalk(raw.clone(), [0, 2], [[0.5, 0.5], [0.5, 0.5]]).cpu().numpy()[0]
    assert (mix[0] == mix[2]).all()
    with pytest.raises(ValueError):
        engine.crosstalk(raw.clone(), [0, 0], np.eye(2))

    devices, chains, _ = fk.reference_generator_setup()
    lines = ["d1", "d2", "d3"]
    instr = fk.drive_instruction("xt", 7e-9, lines, freq=[5e9, 5.3e9, 4.7e9], amp=0.4)
    chains3 = {c: chains["d1"] for c in lines}
    plain = gen_mod.Generator(dict(devices), chains3).generate_signals(instr)
    M = np.array([[0.9, 0.25], [-0.1, 1.05]])
    devices["crosstalk"] = fk.Crosstalk(channels=["d1", "d3"], crosstalk_matrix=fk.Quantity(M, ""))
    crossed = gen_mod.Generator(devices, chains3).generate_signals(instr)
    base = {c: plain[c]["values"].cpu().numpy() for c in lines}
    want = so.crosstalk(base, ["d1", "d3"], M)
    for c in lines:
        assert _rel(crossed[c]["values"].cpu().numpy(), want[c]) < 1e-15, c
    assert _rel(crossed["d1"]["values"].cpu().numpy(), base["d1"]) > 1e-2
    bad = dict(devices)
    bad["crosstalk"] = fk.Crosstalk(channels=["d1", "q9"], crosstalk_matrix=fk.Quantity(M, ""))
    with pytest.raises(Exception, match="C3:ERROR"):
        gen_mod.Generator(bad, chains3).generate_signals(instr)
