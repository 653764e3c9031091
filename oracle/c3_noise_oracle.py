"""CPU oracle for the noise devices of the signal chain (TEST INFRASTRUCTURE ONLY, see c3_oracle.py).

Restates c3/generator/devices.py:943-1035 -- LONoise, Additive_Noise, DC_Noise, Pink_Noise (and the deterministic DC_Offset)
-- as pure functions of explicit random numbers, and the counter-based generator the CUDA chain draws them from
(Philox4x32-10, Salmon et al. SC'11; streams and counters as documented in include/c3b200.h), so that a device realisation
can be checked sample by sample.  The reference itself draws from numpy's global generator (np.random.normal /
np.random.randint / np.random.random): its realisations are not reproducible across implementations, so parity with the
reference is STATISTICAL (the assertions of test/test_noise.py:93-138: standard deviations, constancy of the DC offset,
fresh draws per call, exact zero at zero amplitude) plus exactness of the noise MODEL given the same random numbers.
"""
from __future__ import annotations

import numpy as np

from . import c3_signal_oracle as so

NOISE_KEYS = ("awg_amp", "lo_perc", "add_amp", "dc_amp", "pink_amp", "bfl_num", "dc_offset")
STREAM_AWG, STREAM_LO, STREAM_ADD, STREAM_DC, STREAM_PINK_INIT, STREAM_PINK_FLIP = 1, 2, 3, 4, 5, 6
M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(key, ctr):
    """Vectorised Philox4x32-10: key (k0, k1) scalars, ctr = 4 uint64 arrays holding 32-bit values."""
    k0, k1 = int(key[0]) & MASK, int(key[1]) & MASK
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint64) & MASK for c in ctr]
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & np.uint64(MASK)
        hi1, lo1 = p1 >> np.uint64(32), p1 & np.uint64(MASK)
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def noise_words(seed: int, bk: int, stream: int, idx):
    idx = np.asarray(idx, dtype=np.uint64)
    z = np.zeros_like(idx)
    return philox4x32_10((seed & MASK, (seed >> 32) & MASK), (idx, z + np.uint64(bk), z + np.uint64(stream), z))


def u01(hi, lo):
    m = ((hi >> np.uint64(5)) << np.uint64(26)) | (lo >> np.uint64(6))
    return (m.astype(np.float64) + 0.5) / 9007199254740992.0


def normals(seed: int, bk: int, stream: int, idx):
    w = noise_words(seed, bk, stream, idx)
    r = np.sqrt(-2.0 * np.log(u01(w[0], w[1])))
    ang = 2.0 * np.pi * u01(w[2], w[3])
    return r * np.cos(ang), r * np.sin(ang)


# ---- the reference's noise models as functions of given random numbers ------------------------------------------------

def additive_noise(sig: np.ndarray, noise_amp: float, z: np.ndarray):
    """Additive_Noise.process (devices.py:963-996): sig + noise_amp * z, z ~ N(0,1) per sample; exactly sig below 1e-17."""
    noise = np.zeros_like(sig) if noise_amp < 1e-17 else noise_amp * z
    return sig + noise, noise


def dc_noise(sig: np.ndarray, noise_amp: float, z: float):
    """DC_Noise.get_noise (devices.py:999-1007): one offset noise_amp * z for the whole signal."""
    noise = np.zeros_like(sig) if noise_amp < 1e-17 else np.ones_like(sig) * noise_amp * z
    return sig + noise, noise


def pink_noise(num_steps: int, noise_amp: float, bfl_num: int, init_bits: np.ndarray, u: np.ndarray) -> np.ndarray:
    """Pink_Noise.get_noise (devices.py:1010-1035): bfl_num bistable fluctuators; at every step fluctuator i flips if
    floor(u * flip_rates[i + 1]) == 0 with flip_rates = np.logspace(0, np.log(num_steps), bfl_num + 1, base=10); the noise at a
    step is noise_amp * sum(states) after that step's flips.  ``init_bits [bfl]`` in {0,1}, ``u [bfl, num_steps]`` uniforms."""
    bfls = 2 * np.asarray(init_bits, dtype=np.int64) - 1
    flip_rates = np.logspace(0, np.log(num_steps), num=bfl_num + 1, endpoint=True, base=10.0)
    noise = np.empty(num_steps)
    for step in range(num_steps):
        for i in range(bfl_num):
            if np.floor(u[i, step] * flip_rates[i + 1]) == 0:
                bfls[i] = -bfls[i]
        noise[step] = np.sum(bfls) * noise_amp
    return noise


def lo_noise(cos: np.ndarray, sin: np.ndarray, noise_perc: float, z0: np.ndarray, z1: np.ndarray):
    """LONoise.proces (devices.py:943-960)."""
    return cos + noise_perc * z0, sin + noise_perc * z1


def generate_noisy_signal(envelopes, omega_lo: float, t_start: float, t_end: float, chain: "so.ChainSpec", noise: dict, seed: int,
                          bk: int):
    """One drive line of the reference's chain with noise devices where test/noise_exp_2.hjson puts them:
    LO -> [LONoise];  AWG -> [Additive_Noise] -> DigitalToAnalog -> Response -> Mixer -> [Additive_Noise] -> [DC_Noise] ->
    [Pink_Noise] -> [DC_Offset] -> VoltsToHertz | FluxTuning, drawing the random numbers of realisation (seed, bk).
    Returns (values [N], traces dict)."""
    nz = {k: float(noise.get(k, 0.0)) for k in NOISE_KEYS}
    ts = so.create_ts(t_start, t_end, chain.sim_res)
    N = len(ts)
    lo_i, lo_q = so.lo_signal(ts, omega_lo)
    tr = {}
    if nz["lo_perc"] >= 1e-17:
        z0, z1 = normals(seed, bk, STREAM_LO, np.arange(N))
        tr["lo_cos"], tr["lo_sin"] = nz["lo_perc"] * z0, nz["lo_perc"] * z1
        lo_i, lo_q = lo_noise(lo_i, lo_q, nz["lo_perc"], z0, z1)
    ts_awg = so.create_ts(t_start, t_end, chain.awg_res)
    awg_i, awg_q = so.awg_signal(envelopes, ts_awg, t_start)
    z0, z1 = normals(seed, bk, STREAM_AWG, np.arange(len(ts_awg)))
    awg_i, tr["awg_i"] = additive_noise(awg_i, nz["awg_amp"], z0)
    awg_q, tr["awg_q"] = additive_noise(awg_q, nz["awg_amp"], z1)
    dac_i, dac_q = so.resize_nearest(awg_i, N), so.resize_nearest(awg_q, N)
    resp_i, resp_q = so.response(dac_i, dac_q, chain.rise_time, chain.sim_res, chain.response_fft)
    mixed = so.mixer(lo_i, lo_q, resp_i, resp_q)
    z0, _ = normals(seed, bk, STREAM_ADD, np.arange(N))
    mixed, tr["add"] = additive_noise(mixed, nz["add_amp"], z0)
    zdc, _ = normals(seed, bk, STREAM_DC, np.arange(1))
    mixed, tr["dc"] = dc_noise(mixed, nz["dc_amp"], float(zdc[0]))
    bfl = int(nz["bfl_num"]) if nz["pink_amp"] >= 1e-17 else 0
    tr["pink"] = np.zeros(N)
    if bfl > 0:
        init = np.asarray(noise_words(seed, bk, STREAM_PINK_INIT, np.arange(bfl))[0] & np.uint64(1), dtype=np.int64)
        half = (N + 1) // 2
        idx = (np.arange(bfl)[:, None] * half + np.arange(half)[None, :]).ravel()
        w = noise_words(seed, bk, STREAM_PINK_FLIP, idx)
        u_even = u01(w[0], w[1]).reshape(bfl, half)
        u_odd = u01(w[2], w[3]).reshape(bfl, half)
        u = np.empty((bfl, 2 * half))
        u[:, 0::2], u[:, 1::2] = u_even, u_odd
        tr["pink"] = pink_noise(N, nz["pink_amp"], bfl, init, u[:, :N])
        mixed = mixed + tr["pink"]
    mixed = mixed + nz["dc_offset"]
    values = so.flux_tuning(mixed, **chain.flux) if chain.flux is not None else mixed * chain.v2hz
    return values, tr
