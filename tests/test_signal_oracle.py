"""CPU: the signal-chain oracle (oracle/c3_signal_oracle.py) against the reference's own fixtures, stage by
stage (test/generator_data.pickle <- test/test_generator.py:118-190; AWG samples and final flux-line field of
test/tunable_coupler_data.pickle), and the host logic of the Generator mirror (no GPU needed)."""
import numpy as np
import pytest

import c3_fakes as fk
from oracle import c3_signal_oracle as so

TP = 2 * np.pi


def _rx90p_env():
    t_final = 7e-9
    return so.EnvelopeSpec(shape="gaussian_nonorm", amp=0.5, t_final=t_final, sigma=t_final / 4, xy_angle=0.0,
                           freq_offset=(-50e6 - 3e6) * TP, delta=-1)


def test_generator_chain_stage_by_stage(golden_generator):
    g = golden_generator
    st = {}
    vals, ts = so.generate_signal([_rx90p_env()], (5e9 + 50e6) * TP, 0.0, 7e-9, so.ChainSpec(), st)
    assert np.array_equal(ts, g["lo_ts"])                                       # test_LO, :118-127
    assert np.abs(st["lo_i"] - g["lo_I"]).max() < 1e-12 and np.abs(st["lo_q"] - g["lo_Q"]).max() < 1e-12
    assert np.abs(st["awg_i"] - g["awg_I"]).max() < 1e-14                       # test_AWG, :131-138
    assert np.abs(st["awg_q"] - g["awg_Q"]).max() < 1e-14
    assert np.array_equal(so.resize_nearest(g["awg_I"], 700), g["dac_I"])       # test_DAC, :142-151
    assert np.array_equal(so.resize_nearest(g["awg_Q"], 700), g["dac_Q"])
    ri, rq = so.response(g["dac_I"], g["dac_Q"], 0.3e-9, 100e9)                 # test_Response, :155-164
    assert np.abs(ri - g["resp_I"]).max() < 1e-14 and np.abs(rq - g["resp_Q"]).max() < 1e-14
    assert np.array_equal(so.mixer(g["lo_I"], g["lo_Q"], g["resp_I"], g["resp_Q"]), g["mixer"])   # test_mixer
    assert np.abs(g["mixer"] * 1e9 - g["v2hz"]).max() < 1e-6                    # test_v2hz
    scale = np.abs(g["full_values"]).max()
    assert np.abs(vals - g["full_values"]).max() < 1e-12 * scale                # test_full_signal_chain, :183-190


def test_tunable_coupler_flux_line(golden_tunable_coupler):
    """AWG at 2.4 GS/s (non-integer resampling ratio 41.67: pins the half-pixel nearest-neighbour rule), flattop
    envelope, FluxTuning output: the 10 000-sample control field of test/test_tunable_coupler.py."""
    g = golden_tunable_coupler
    env = so.EnvelopeSpec(shape="flattop", amp=1.0, t_final=100e-9, t_up=5e-9, t_down=95e-9, risefall=5e-9,
                          xy_angle=0.3590456701578104)
    chain = so.ChainSpec(awg_res=2.4e9, flux=dict(phi=2.3, phi_0=10.0, omega_0=8.1e9 * TP, anhar=-286e6 * TP, d=0.36))
    st = {}
    vals, ts = so.generate_signal([env], 829e6 * TP, 0.0, 100e-9, chain, st)
    assert np.array_equal(st["ts_awg"], g["tc_awg_ts"]) and np.array_equal(ts, g["tc_ts"])
    assert np.abs(st["awg_i"] - g["tc_awg_I"]).max() < 1e-14 and np.abs(st["awg_q"] - g["tc_awg_Q"]).max() < 1e-14
    assert np.abs(vals - g["tc_signal"]).max() < 1e-12 * np.abs(g["tc_signal"]).max()


def test_convolutions_are_plain_sums():
    rng = np.random.default_rng(0)
    x, r = rng.normal(size=57), rng.normal(size=9)
    full = np.convolve(x, r)
    assert np.allclose(so.tf_convolve(x, r).real, full[:57], atol=1e-13)
    assert np.allclose(so.tf_convolve_legacy(x, r).real, np.concatenate([[0.0], full[:56]]), atol=1e-13)


@pytest.mark.parametrize("shape", ["gaussian_nonorm", "gaussian_sigma", "cosine", "flattop"])
def test_shape_derivatives(shape):
    e = so.EnvelopeSpec(shape=shape, t_final=8e-9, sigma=1.7e-9, t_up=1e-9, t_down=6.5e-9, risefall=0.8e-9)
    t = np.linspace(0.2e-9, 7.8e-9, 41)
    h = 1e-15
    fd = (so.shape_values(shape, t + h, e) - so.shape_values(shape, t - h, e)) / (2 * h)
    an = so.shape_derivative(shape, t, e)
    assert np.abs(fd - an).max() < 1e-4 * np.abs(an).max()


def test_generator_mirror_host_logic():
    from c3_b200.generator import Generator, CHAIN_KEYS
    devices, chains, instr = fk.reference_generator_setup()
    gen = Generator(devices, chains)
    spec = gen._specs["d1"]
    assert spec["sim_res"] == 100e9 and spec["awg_res"] == 2e9 and spec["resp_kind"] == 1.0 and spec["v2hz"] == 1e9
    chans, env, shape, flags, lo, chain = gen._tables(instr, {("d1", "gauss", "amp"): [0.1, 0.2, 0.3]})
    assert chans == ["d1"] and env.shape == (3, 1, 1, 9) and list(env[:, 0, 0, 0]) == [0.1, 0.2, 0.3]
    assert shape[0, 0] == 2 and flags[0, 0] == 0 and abs(lo[0, 0] - 5.05e9 * TP) < 1 and chain.shape == (1, len(CHAIN_KEYS))
    assert abs(env[0, 0, 0, 4] - (-53e6 * TP)) < 1e-3
    d2, c2, i2 = fk.tunable_coupler_flux_setup()
    s2 = Generator(d2, c2)._specs["TC"]
    assert s2["out_kind"] == 1.0 and s2["d"] == 0.36 and abs(s2["omega_0"] - 8.1e9 * TP) < 1
    # unsupported pieces fail loudly (no silent CPU path)
    bad = dict(devices)
    bad["Noise"] = fk.Additive_Noise("noise")
    with pytest.raises(Exception, match="C3:ERROR"):
        Generator(bad, {"d1": dict(chains["d1"], Noise=["Mixer"])})
    with pytest.raises(Exception, match="C3:ERROR"):
        Generator(devices, {"d1": dict(chains["d1"], Mixer=["Response", "LO"])})
    instr.add_component(fk.Envelope("odd", "slepian_fourier", {}), "d1")          # known shape, its parameters missing
    with pytest.raises(Exception, match="C3:ERROR.*lacks the parameter"):
        gen._tables(instr)
    instr.comps["d1"]["odd"].shape = fk._Shape("hann_window")                     # a shape the reference does not have either
    with pytest.raises(Exception, match="C3:ERROR"):
        gen._tables(instr)


ENVELOPE_SHAPES = ["trapezoid", "flattop_risefall", "flattop_risefall_1ns", "flattop", "gaussian_sigma", "gaussian", "gaussian_nonorm",
                   "gaussian_der_nonorm", "gaussian_der", "drag_sigma", "drag_der", "drag", "cosine", "no_drive", "rect"]


@pytest.mark.parametrize("shape", ENVELOPE_SHAPES)
def test_envelope_shapes_against_the_reference_pickle(shape):
    """test/test_envelopes.py of the reference: every shape the on-device chain implements, on ts = linspace(0, 10 ns, 100) with
    t_final 10 ns, sigma 5 ns, risefall 2 ns, t_up 1 ns, t_down 10 ns, against test/envelopes.pickle (tolerance of the
    reference's own assertions: 1e-11 of the maximum)."""
    import os
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, "envelopes.npz")))
    # test_gaussian of the reference evaluates gaussian_sigma with sigma = 5 ns, then "gaussian", which OVERWRITES params["sigma"]
    # with t_final / 6 (envelopes.py:411-416) -- every shape evaluated after it in that test saw sigma = 10 ns / 6
    sigma = 10e-9 / 6 if shape in ("gaussian_nonorm", "gaussian_der_nonorm", "gaussian_der", "drag_sigma", "drag_der") else 5e-9
    e = so.EnvelopeSpec(shape=shape, t_final=10e-9, sigma=sigma, risefall=2e-9, t_up=1e-9, t_down=10e-9)
    got = so.shape_values(shape, g["ts"], e)
    assert np.abs(got - g[shape]).max() <= 1e-11 * max(np.abs(g[shape]).max(), 1e-300)


#: golden key -> (shape, EnvelopeSpec fields, extra parameters): the parameters of test/test_envelopes.py:33-130,158-207,279-289
_PWC = dict(t_bin_start=1e-10, t_bin_end=9.9e-9, inphase=[0, 0.1, 0.3, 0.5, 0.1, 1.1, 0.4, 0.1])
_SLEP = dict(width=9e-9, fourier_coeffs=[1, 0.5, 0.2], offset=0.1)
EXTENDED_CASES = {
    "pwc_shape": ("pwc_shape", {}, _PWC),
    "pwc_symmetric": ("pwc_symmetric", {}, _PWC),
    "pwc_shape_plateau1": ("pwc_shape_plateau", {}, _PWC),
    "pwc_shape_plateau2": ("pwc_shape_plateau", {}, dict(_PWC, width=5e-9)),
    "delta_pulse": ("delta_pulse", {}, dict(t_sig=[0.5e-9])),
    "fourier_sin": ("fourier_sin", {}, dict(amps=[0.5, 0.2], freqs=[1e6, 1e10], phases=[0, 1])),
    "fourier_cos": ("fourier_cos", {}, dict(amps=[0.5, 0.2], freqs=[1e6, 1e10], phases=[0, 1])),
    "slepian_fourier": ("slepian_fourier", dict(amp=0.5), _SLEP),
    "slepian_fourier_risefall": ("slepian_fourier", dict(amp=0.5), dict(_SLEP, risefall=4e-9)),
    "slepian_fourier_sin": ("slepian_fourier", dict(amp=0.5), dict(_SLEP, risefall=4e-9, sin_coeffs=[0.3])),
    "flattop_variant": ("flattop_variant", dict(t_up=1e-9, t_down=10e-9), dict(ramp=2e-9)),
    "flattop_cut": ("flattop_cut", dict(risefall=2e-9, t_up=1e-9, t_down=10e-9), {}),
    "flattop_cut_center": ("flattop_cut_center", dict(risefall=2e-9), dict(width=9e-9)),
    "cosine_flattop": ("cosine_flattop", {}, dict(t_rise=2e-9)),
}


@pytest.mark.parametrize("key", sorted(EXTENDED_CASES))
def test_extended_envelope_shapes_against_the_reference_pickle(key):
    """The array-parametrised and grid-defined shapes (pwc_*, delta_pulse, fourier_*, slepian_fourier, flattop_cut*,
    flattop_variant, cosine_flattop) against test/envelopes.pickle with the parameters of the reference's tests.  This also
    pins the restated tfp.math.interp_regular_1d_grid (tensorflow_probability is not vendored in the reference)."""
    import os
    from conftest import GOLDEN
    g = dict(np.load(os.path.join(GOLDEN, "envelopes.npz")))
    shape, fields, extra = EXTENDED_CASES[key]
    e = so.EnvelopeSpec(shape=shape, t_final=10e-9, **fields, extra=dict(extra))
    got = so.shape_values(shape, g["ts"], e)
    assert np.abs(got - g[key]).max() <= 1e-11 * max(np.abs(g[key]).max(), 1e-300)


@pytest.mark.parametrize("shape", [s_ for s_ in ENVELOPE_SHAPES if s_ not in ("no_drive", "rect")])
def test_envelope_time_derivatives(shape):
    """shape_derivative (the DRAG quadrature's d env / dt) against a central difference of the pinned shape function."""
    e = so.EnvelopeSpec(shape=shape, t_final=10e-9, sigma=2.2e-9, risefall=1.3e-9, t_up=2e-9, t_down=8e-9)
    t = np.linspace(0.3e-9, 9.7e-9, 57)
    t = t[(np.abs(t - 2.5 * e.risefall) > 1e-11) & (np.abs(t - (e.t_final - 2.5 * e.risefall)) > 1e-11)]   # trapezoid kinks
    h = 1e-14
    fd = (so.shape_values(shape, t + 1e-13, e) - so.shape_values(shape, t - 1e-13, e)) / 2e-13
    got = so.shape_derivative(shape, t, e)
    assert np.abs(got - fd).max() <= 2e-5 * max(np.abs(fd).max(), 1e-300)


def test_crosstalk_known_answers():
    """test/test_crosstalk.py:7-27 of the reference: identity, flip and equal mix of two lines."""
    sig = {"TC1": np.linspace(0, 100, 101), "TC2": np.linspace(100, 200, 101), "other": np.ones(101)}
    same = so.crosstalk(sig, ["TC1", "TC2"], [[1, 0], [0, 1]])
    assert all((same[k] == sig[k]).all() for k in sig)
    flip = so.crosstalk(sig, ["TC1", "TC2"], [[0, 1], [1, 0]])
    assert (flip["TC2"] == np.linspace(0, 100, 101)).all() and (flip["TC1"] == np.linspace(100, 200, 101)).all()
    mix = so.crosstalk(sig, ["TC1", "TC2"], [[0.5, 0.5], [0.5, 0.5]])
    assert (mix["TC1"] == mix["TC2"]).all() and (mix["other"] == 1).all()
