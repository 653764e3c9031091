"""GPU tests of the noise devices in the on-device signal chain (c3/generator/devices.py:943-1035).

Two kinds of evidence: (1) exactness of the noise MODEL -- every realised trace and the resulting control field equal the CPU
oracle's restatement fed with the same counter-based random numbers; (2) the STATISTICAL assertions of the reference's own
test (test/test_noise.py:93-138): standard deviations in band, a DC offset that is constant in time and fresh per call, exact
zeros at zero amplitude, fidelities that move when noise is on."""
import numpy as np
import pytest
import torch

import c3_fakes as fk
from oracle import c3_noise_oracle as no
from oracle import c3_signal_oracle as so

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from c3_b200 import engine, generator
    return engine, generator


def _one_line(engine, B, noise_row, seed, N_awg_res=2e9, t_end=7e-9):
    K, E = 1, 1
    env = np.zeros((B, K, E, 9))
    env[..., 0] = 0.5; env[..., 1] = t_end; env[..., 2] = t_end / 4; env[..., 4] = -2 * np.pi * 53e6; env[..., 8] = 1.0
    shape = np.full((K, E), 2, dtype=np.int32)
    flags = np.zeros((K, E), dtype=np.int32)
    lo = np.full((B, K), 2 * np.pi * 5.05e9)
    chain = np.array([[100e9, N_awg_res, 0.3e-9, 1, 0, 1e9, 0, 1, 0, 0, np.nan]])
    sig, tr = engine.generate_signals(env, shape, flags, lo, chain, 0.0, t_end, noise=np.array([noise_row]), seed=seed, return_noise=True)
    spec = so.EnvelopeSpec(shape="gaussian_nonorm", amp=0.5, t_final=t_end, sigma=t_end / 4, freq_offset=-2 * np.pi * 53e6)
    return sig.cpu().numpy(), tr.cpu().numpy(), spec, so.ChainSpec(awg_res=N_awg_res)


def test_noise_traces_and_fields_match_the_oracle(mods):
    """All devices on at once, three realisations (batch rows): every trace and the final field against the oracle."""
    engine, _ = mods
    row = [0.02, 0.01, 0.03, 0.05, 0.04, 9, 0.007]
    B, seed = 3, 0x1234567890ABCDEF
    sig, tr, spec, cs = _one_line(engine, B, row, seed)
    noise = dict(zip(no.NOISE_KEYS, row))
    for b in range(B):
        want, wtr = no.generate_noisy_signal([spec], 2 * np.pi * 5.05e9, 0.0, 7e-9, cs, noise, seed, b)
        n_awg = len(wtr["awg_i"])
        t = {name: tr[b, 0, i] for i, name in enumerate(engine.NOISE_TRACES)}
        np.testing.assert_allclose(t["awg_i"][:n_awg], wtr["awg_i"], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(t["awg_q"][:n_awg], wtr["awg_q"], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(t["lo_cos"], wtr["lo_cos"], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(t["add"], wtr["add"], rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(t["dc"], wtr["dc"], rtol=1e-12, atol=1e-15)
        np.testing.assert_array_equal(np.rint(t["pink"] / row[4]), np.rint(wtr["pink"] / row[4]))      # integer fluctuator sums
        assert np.linalg.norm(sig[b, 0] - want) < 1e-11 * np.linalg.norm(want)
    assert not np.allclose(sig[0], sig[1])                 # rows are independent realisations


def test_noise_statistics_like_the_reference_test(mods):
    """test/test_noise.py:105-138 on one device at a time: params = 0.1 on pink / dc / awg noise, two calls A and B."""
    engine, _ = mods
    for which in range(4):
        amp = [0.0, 0.0, 0.0]
        if which < 3:
            amp[which] = 0.1
        row = [amp[2], 0.0, 0.0, amp[1], amp[0], 15, 0.0]
        sigA, trA, _, _ = _one_line(engine, 1, row, seed=11)
        sigB, trB, _, _ = _one_line(engine, 1, row, seed=12)
        t = {n: i for i, n in enumerate(engine.NOISE_TRACES)}
        pinkA, pinkB = trA[0, 0, t["pink"]], trB[0, 0, t["pink"]]
        dcA, dcB = trA[0, 0, t["dc"]], trB[0, 0, t["dc"]]
        awgA, awgB = trA[0, 0, t["awg_i"]][:14], trB[0, 0, t["awg_i"]][:14]
        assert np.std(pinkA) >= 0.05 * amp[0] and np.std(pinkA) < 10 * amp[0] + 1e-15
        if amp[0] > 1e-15:
            assert np.median(np.abs(pinkA - pinkB) > 1e-10)
        if amp[1] > 1e-15:
            assert 1e-6 < np.abs(np.mean(dcA - dcB)) < 10 * amp[1]
        else:
            assert np.max(dcA - dcB) < 1e-15
        assert np.std(dcA) < 1e-15
        assert np.std(awgA) >= 0.05 * amp[2] and np.std(awgA) < 10 * amp[2] + 1e-15
        if amp[2] > 1e-15:
            assert np.mean(np.abs(awgA - awgB) > 1e-10)
        if max(amp) > 0:
            assert not np.array_equal(sigA, sigB)
        else:
            assert np.array_equal(sigA, sigB)               # all amplitudes zero: bit-identical, and identical to the noise-free chain


def test_generator_with_noise_devices(mods):
    """The chain of test/noise_exp_2.hjson through Generator: zero amplitudes reproduce the noise-free chain bit for bit;
    switched on, every call is a fresh realisation, the devices carry their realised noise like the reference's, a fixed seed
    reproduces, and a batch is a set of independent trajectories that moves the gate infidelity."""
    engine, generator = mods
    dev0, ch0, instr = fk.reference_generator_setup()
    clean = generator.Generator(dev0, ch0).generate_signals(instr)["d1"]["values"].cpu().numpy()
    devices, chains, instr = fk.noisy_generator_setup()
    gen = generator.Generator(devices, chains)
    quiet = gen.generate_signals(instr)["d1"]["values"].cpu().numpy()
    assert np.array_equal(quiet, clean)
    assert float(devices["PinkNoise"].signal["noise"].abs().max()) == 0.0
    devices["PinkNoise"].params["noise_amp"] = fk.Quantity(0.1, "V")
    devices["DCNoise"].params["noise_amp"] = fk.Quantity(0.05, "V")
    devices["AWGNoise"].params["noise_amp"] = fk.Quantity(0.02, "V")
    a = gen.generate_signals(instr)["d1"]["values"].cpu().numpy()
    dcA = devices["DCNoise"].signal["noise"].cpu().numpy()
    b = gen.generate_signals(instr)["d1"]["values"].cpu().numpy()
    dcB = devices["DCNoise"].signal["noise"].cpu().numpy()
    assert not np.array_equal(a, b) and not np.array_equal(a, clean)
    assert np.std(dcA) < 1e-15 and abs(dcA[0] - dcB[0]) > 1e-6
    assert devices["AWGNoise"].signal["noise-inphase"].shape[0] == 14
    gen2 = generator.Generator(devices, chains)
    gen2.noise_draws = gen.noise_draws - 1
    assert np.array_equal(gen2.generate_signals(instr)["d1"]["values"].cpu().numpy(), b)      # same (seed, draw) -> same realisation
    # Monte-Carlo axis: 64 trajectories of the same pulse in one launch
    from c3_b200 import synth
    sig, ts = gen.generate_signals_batch(instr, {("d1", "gauss", "amp"): np.full(64, 0.5)})
    assert sig.shape[0] == 64 and float((sig[0] - sig[1]).abs().max()) > 0
    m = synth.one_qubit()
    U = engine.pwc_closed(m.h0, m.hks, sig, float(ts[1] - ts[0]))
    ideal = np.array([[1, -1j], [-1j, 1]]) / np.sqrt(2)
    infid = engine.gate_infid(U, ideal, [0, 1]).cpu().numpy()
    assert infid.std() > 0


def test_unsupported_chain_still_raises(mods):
    _, generator = mods
    devices, chains, _ = fk.noisy_generator_setup()
    chains["d1"]["DCNoise2"] = ["PinkNoise"]
    devices["DCNoise2"] = fk.DC_Noise("dc2", 100e9, noise_amp=fk.Quantity(0.1, "V"))
    chains["d1"]["DCOffset"] = ["DCNoise2"]
    with pytest.raises(Exception, match="C3:ERROR"):
        generator.Generator(devices, chains)
