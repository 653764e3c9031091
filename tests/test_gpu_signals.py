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
