"""CPU tests of the noise oracle: the counter-based generator against its published known-answer vectors, and the
reference's noise models (c3/generator/devices.py:943-1035) as functions of given random numbers."""
import numpy as np

from oracle import c3_noise_oracle as no


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32-10."""
    z = np.array([0])
    out = no.philox4x32_10((0, 0), (z, z, z, z))
    assert [int(x[0]) for x in out] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    f = np.array([0xFFFFFFFF])
    out = no.philox4x32_10((0xFFFFFFFF, 0xFFFFFFFF), (f, f, f, f))
    assert [int(x[0]) for x in out] == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    pi = (np.array([0x243F6A88]), np.array([0x85A308D3]), np.array([0x13198A2E]), np.array([0x03707344]))
    out = no.philox4x32_10((0xA4093822, 0x299F31D0), pi)
    assert [int(x[0]) for x in out] == [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_normals_are_standard_and_streams_independent():
    z0, z1 = no.normals(7, 3, no.STREAM_ADD, np.arange(200000))
    for z in (z0, z1):
        assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    assert abs(np.corrcoef(z0, z1)[0, 1]) < 0.01
    y0, _ = no.normals(7, 4, no.STREAM_ADD, np.arange(200000))          # another line
    x0, _ = no.normals(8, 3, no.STREAM_ADD, np.arange(200000))          # another seed
    assert abs(np.corrcoef(z0, y0)[0, 1]) < 0.01 and abs(np.corrcoef(z0, x0)[0, 1]) < 0.01
    a0, _ = no.normals(7, 3, no.STREAM_ADD, np.arange(10))
    assert np.array_equal(a0, z0[:10])                                  # counter-based: a prefix is a prefix


def test_pink_noise_model():
    """Pink_Noise.get_noise: values are amp * (sum of bfl_num signs); fast fluctuators flip often, slow ones almost never."""
    rng = np.random.default_rng(0)
    N, bfl, amp = 700, 15, 0.1
    u = rng.random((bfl, N))
    init = rng.integers(0, 2, bfl)
    noise = no.pink_noise(N, amp, bfl, init, u)
    k = np.rint(noise / amp).astype(int)
    assert np.allclose(noise, k * amp) and np.all(np.abs(k) <= bfl) and np.all((k - bfl) % 2 == 0)
    assert 0.05 * amp <= noise.std() < 10 * amp                        # the band test/test_noise.py:119-120 asserts
    # u = 1 - eps never flips (floor(u * rate) >= 1 for every rate > 1 ... the first rate is 10^(ln N / bfl) > 1)
    frozen = no.pink_noise(N, amp, bfl, init, np.full((bfl, N), 0.999999))
    assert np.all(frozen == frozen[0])


def test_additive_and_dc_models():
    sig = np.linspace(0, 1, 50)
    z = np.random.default_rng(1).normal(size=50)
    out, noise = no.additive_noise(sig, 0.1, z)
    assert np.allclose(out - sig, 0.1 * z) and np.allclose(noise, 0.1 * z)
    out, noise = no.additive_noise(sig, 8.7e-18, z)                    # the "off" value of test/noise_exp_2.hjson
    assert np.array_equal(out, sig) and not noise.any()
    out, noise = no.dc_noise(sig, 0.1, 1.7)
    assert np.allclose(noise, 0.17) and noise.std() < 1e-15
