"""CPU oracle for SURVEY.md section 8f row f-2: the signal-generation chain that produces the control
fields ``signals[K,N]`` the propagator consumes.

TEST INFRASTRUCTURE ONLY (see oracle/c3_oracle.py).  numpy restatement of the standard chain
LO + AWG -> DigitalToAnalog -> Response -> Mixer -> VoltsToHertz / FluxTuning:

  c3/generator/devices.py:73-122     Device.calc_slice_num / create_ts
  c3/generator/devices.py:1063-1130  LO.process (noise-free branch)
  c3/generator/devices.py:1159-1197  AWG.create_IQ  ->  c3/signal/gates.py:341-370 Instruction.get_awg_signal
  c3/signal/pulse.py:93-142          Envelope.compute_mask / _get_shape_values_just / _before
  c3/signal/pulse.py:171-180         EnvelopeDrag.get_shape_values (derivative quadrature)
  c3/libraries/envelopes.py          no_drive :25, rect :194, flattop :253, gaussian_sigma :373, cosine :420,
                                     gaussian_nonorm :469
  c3/generator/devices.py:296-351    DigitalToAnalog.process (tf.image.resize, method "nearest")
  c3/generator/devices.py:585-701    Response (tf_convolve_legacy) and ResponseFFT (tf_convolve)
  c3/utils/tf_utils.py:441-515       tf_convolve, tf_convolve_legacy
  c3/generator/devices.py:906-939    Mixer.process
  c3/generator/devices.py:187-221    VoltsToHertz.process
  c3/generator/devices.py:480-529    FluxTuning.get_factor / get_freq / process

Pinned stage by stage to the reference's own fixtures (tests/golden/generator.npz from
test/generator_data.pickle, test/test_generator.py:118-190; the AWG samples and the final flux-line field of
test/tunable_coupler_data.pickle) in tests/test_signal_oracle.py.
"""
from __future__ import annotations

from dataclasses import dataclass, replace, field
from typing import Dict, List, Optional, Sequence

import numpy as np
from scipy.special import erf, expit

SHAPES = ("no_drive", "rect", "gaussian_nonorm", "gaussian_sigma", "cosine", "flattop")


@dataclass
class EnvelopeSpec:
    """One Envelope component of an instruction channel (values as ``Quantity.get_value()`` returns them,
    i.e. frequencies already multiplied by 2 pi)."""
    shape: str = "gaussian_nonorm"
    amp: float = 0.0
    t_final: float = 0.0
    sigma: float = 0.0
    xy_angle: float = 0.0
    freq_offset: float = 0.0
    delta: float = 0.0
    t_up: float = 0.0
    t_down: float = 0.0
    risefall: float = 1.0
    drag: bool = False          # EnvelopeDrag: quadrature = -delta * d env/dt * dt
    use_t_before: bool = False
    extra: Dict = field(default_factory=dict)   # parameters outside the nine scalars: width, ramp, t_rise, t_sig, inphase,
                                                # quadrature, t_bin_start, t_bin_end, amps, freqs, phases, fourier_coeffs,
                                                # sin_coeffs, offset (and risefall as "present" for slepian_fourier)


def create_ts(t_start: float, t_end: float, resolution: float, centered: bool = True) -> np.ndarray:
    slice_num = int(np.abs(t_start - t_end) * resolution)
    dt = 1 / resolution
    if centered:
        offset, num = dt / 2, slice_num
    else:
        offset, num = 0.0, slice_num + 1
    return np.linspace(t_start + offset, t_end - offset, num)


def lo_signal(ts: np.ndarray, omega_lo: float):
    return np.cos(omega_lo * ts), np.sin(omega_lo * ts)


def shape_values(shape: str, t: np.ndarray, e: EnvelopeSpec) -> np.ndarray:
    """Real envelope shape functions (the reference returns them complexified with zero imaginary part)."""
    t = np.asarray(t, dtype=np.float64)
    if shape == "no_drive":
        return np.zeros_like(t)
    if shape == "rect":
        return np.ones_like(t)
    if shape == "gaussian_nonorm":
        return np.exp(-((t - e.t_final / 2) ** 2) / (2 * e.sigma ** 2))
    if shape == "gaussian_sigma":
        gauss = np.exp(-((t - e.t_final / 2) ** 2) / (2 * e.sigma ** 2))
        offset = np.exp(-(e.t_final ** 2) / (8 * e.sigma ** 2))
        norm = np.sqrt(2 * np.pi * e.sigma ** 2) * erf(e.t_final / (np.sqrt(8) * e.sigma)) - e.t_final * offset
        return (gauss - offset) / norm
    if shape == "cosine":
        return 0.5 * (1 - np.cos(2 * np.pi * t / e.t_final))
    if shape == "flattop":
        return (1 + erf((t - e.t_up) / e.risefall)) / 2 * (1 + erf((-t + e.t_down) / e.risefall)) / 2
    if shape == "trapezoid":                       # c3/libraries/envelopes.py:200-224
        env = np.ones_like(t)
        env = np.where(t <= e.risefall * 2.5, t / (e.risefall * 2.5), env)
        env = np.where(t >= e.t_final - e.risefall * 2.5, (e.t_final - t) / (e.risefall * 2.5), env)
        return env
    if shape in ("flattop_risefall", "flattop_risefall_1ns"):      # :227-250, :366-370
        rf = 1e-9 if shape.endswith("1ns") else e.risefall
        t_up, t_down = rf, e.t_final - rf
        return (1 + erf((t - t_up) / rf)) / 2 * (1 + erf((-t + t_down) / rf)) / 2
    if shape == "gaussian":                        # :399-417: gaussian_sigma with sigma = t_final / 6
        return shape_values("gaussian_sigma", t, replace(e, sigma=e.t_final / 6))
    if shape in ("gaussian_der_nonorm", "gaussian_der"):           # :490-516
        g = np.exp(-((t - e.t_final / 2) ** 2) / (2 * e.sigma ** 2)) * (t - e.t_final / 2) / e.sigma ** 2
        if shape == "gaussian_der":
            # the reference evaluates sqrt(8) in float32 here and only here (tf.cast(tf.sqrt(8.0), tf.float64), envelopes.py:514):
            # 2.8284270763397217 instead of 2.8284271247461903, visible at 5e-10 relative -- restated as is
            s8 = float(np.float64(np.sqrt(np.float32(8.0))))
            g = g / (np.sqrt(2 * np.pi * e.sigma ** 2) * erf(e.t_final / (s8 * e.sigma)) - e.t_final * np.exp(-(e.t_final ** 2) / (8 * e.sigma ** 2)))
        return g
    if shape in ("drag_sigma", "drag"):            # :519-542 (drag: sigma = t_final / 4)
        if shape == "drag":
            e = replace(e, sigma=e.t_final / 4)
        gauss = np.exp(-((t - e.t_final / 2) ** 2) / (2 * e.sigma ** 2))
        offset = np.exp(-(e.t_final ** 2) / (8 * e.sigma ** 2))
        return (gauss - offset) ** 2 / _gauss_norm(e)
    if shape == "drag_der":                        # :545-562
        gauss = np.exp(-((t - e.t_final / 2) ** 2) / (2 * e.sigma ** 2))
        offset = np.exp(-(e.t_final ** 2) / (8 * e.sigma ** 2))
        return -2 * (gauss - offset) * gauss * (t - e.t_final / 2) / e.sigma ** 2 / _gauss_norm(e)
    x = e.extra
    if shape == "flattop_cut":                     # :281-302
        v = erf((t - e.t_up) / e.risefall) * erf((-t + e.t_down) / e.risefall)
        v = np.clip(v, 0, 1)
        return v / np.max(v)
    if shape == "flattop_cut_center":              # :305-327
        t_up, t_down = e.t_final / 2 - x["width"] / 2, e.t_final / 2 + x["width"] / 2
        return np.clip(erf((t - t_up) / e.risefall) * erf((-t + t_down) / e.risefall), 0, 2)
    if shape == "flattop_variant":                 # :565-587
        ramp = min(x["ramp"], (e.t_down - e.t_up) / 2)
        sigma = np.sqrt(2) * ramp * 0.2
        v = np.zeros_like(t)
        for i, tt in enumerate(t):
            if e.t_up <= tt <= e.t_up + ramp:
                v[i] = np.exp(-((tt - e.t_up - ramp) ** 2) / (2 * sigma ** 2))
            elif e.t_up + ramp < tt < e.t_down - ramp:
                v[i] = 1
            elif e.t_down >= tt >= e.t_down - ramp:
                v[i] = np.exp(-((tt - e.t_down + ramp) ** 2) / (2 * sigma ** 2))
        return v
    if shape == "cosine_flattop":                  # :440-466: indices, and the SAME first n_rise time values for the fall
        t_rise = x["t_rise"]
        n_rise = int(t_rise / (t[1] - t[0]))
        n_flat = len(t) - 2 * n_rise
        return np.concatenate([0.5 * (1 - np.cos(np.pi * t[:n_rise] / t_rise)), np.ones(n_flat),
                               0.5 * (1 + np.cos(np.pi * t[:n_rise] / t_rise))])
    if shape == "delta_pulse":                     # :128-139
        v = np.zeros_like(t)
        for t_s in np.atleast_1d(x["t_sig"]):
            dist = (t - t_s - 1e-9) ** 2
            v = np.where(np.min(dist) == dist, 1.0, v)
        return v
    if shape == "pwc":                             # :31-34 (complex: in-phase + i quadrature, one value per AWG sample)
        return np.asarray(x["inphase"], dtype=np.float64) + 1j * np.asarray(x["quadrature"], dtype=np.float64)
    if shape == "pwc_shape":                       # :37-68
        return interp_regular_1d_grid(t, x["t_bin_start"], x["t_bin_end"], x["inphase"])
    if shape == "pwc_symmetric":                   # :104-125
        return interp_regular_1d_grid(np.where(t > e.t_final / 2, -t + e.t_final, t), x["t_bin_start"], x["t_bin_end"], x["inphase"])
    if shape == "pwc_shape_plateau":               # :71-101
        if "width" not in x:
            return interp_regular_1d_grid(t, x["t_bin_start"], x["t_bin_end"], x["inphase"])
        plateau = x["width"] - (x["t_bin_end"] - x["t_bin_start"])
        t_mid = (x["t_bin_end"] - x["t_bin_start"]) / 2
        xx = t.copy()
        xx = np.where(t > t_mid + plateau, t - plateau, xx)
        xx = np.where(t < t_mid, t, xx)
        xx = np.where((t < t_mid + plateau) & (t > t_mid), t_mid, xx)
        v = interp_regular_1d_grid(xx, x["t_bin_start"], x["t_bin_end"], x["inphase"])
        return np.where(xx == t_mid, 1.0, v)
    if shape == "fourier_sin":                     # :142-168
        a, f, ph = (np.asarray(x[k], dtype=np.float64).reshape(-1, 1) for k in ("amps", "freqs", "phases"))
        return np.sum(a * np.sin(f * t.reshape(1, -1) + ph), axis=0)
    if shape == "fourier_cos":                     # :171-191
        a, f = (np.asarray(x[k], dtype=np.float64).reshape(-1, 1) for k in ("amps", "freqs"))
        return np.sum(a * np.cos(f * t.reshape(1, -1)), axis=0)
    if shape == "slepian_fourier":                 # :330-363
        width = x["width"]
        if "risefall" in x:
            plateau = width - x["risefall"] * 2
            xx = t.copy()
            xx = np.where(t > (e.t_final + plateau) / 2, t - plateau / 2, xx)
            xx = np.where(t < (e.t_final - plateau) / 2, t + plateau / 2, xx)
            xx = np.where(np.abs(t - e.t_final / 2) < plateau / 2, e.t_final / 2, xx)
            length = x["risefall"] * 2
        else:
            xx, length = t, width
        v = np.zeros_like(t)
        for n, coeff in enumerate(np.atleast_1d(x["fourier_coeffs"])):
            v = v + coeff * (1 - np.cos(2 * np.pi * (n + 1) * (xx - (e.t_final - length) / 2) / length))
        for n, coeff in enumerate(np.atleast_1d(x.get("sin_coeffs", []))):
            v = v + coeff * np.sin((np.pi * (2 * n + 1)) * (xx - (e.t_final - length) / 2) / length)
        v = np.where(np.abs(e.t_final / 2 - t) > width / 2, 0.0, v)
        v = v / np.max(v)
        return v * (1 - x["offset"] / e.amp) + x["offset"] / e.amp
    raise ValueError(f"C3:ERROR: envelope shape '{shape}' is not restated")


def interp_regular_1d_grid(x, x_min: float, x_max: float, y_ref, fill: float = 0.0) -> np.ndarray:
    """tfp.math.interp_regular_1d_grid(x, x_ref_min, x_ref_max, y_ref, fill_value_below = fill_value_above = 0): y_ref sits on
    linspace(x_min, x_max, len(y_ref)); linear interpolation inside, the fill value outside (tensorflow_probability
    math/interpolation.py; the dependency is absent from /root/reference -- pinned through the envelope golden vectors that
    use it: pwc_shape, pwc_symmetric, pwc_shape_plateau1/2)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y_ref, dtype=np.float64).reshape(-1)
    ny = len(y)
    u = (x - x_min) / (x_max - x_min) * (ny - 1)
    uc = np.clip(u, 0, ny - 1)
    lo = np.floor(uc).astype(int)
    hi = np.minimum(lo + 1, ny - 1)
    lo = np.maximum(hi - 1, 0)
    w = uc - lo
    v = w * y[hi] + (1 - w) * y[lo]
    return np.where((x < x_min) | (x > x_max), fill, v)


def _gauss_norm(e) -> float:
    """sqrt(2 pi sigma^2) erf(t_final / (sqrt(8) sigma)) - t_final exp(-t_final^2 / (8 sigma^2))  (envelopes.py:391-394)."""
    return np.sqrt(2 * np.pi * e.sigma ** 2) * erf(e.t_final / (np.sqrt(8) * e.sigma)) - e.t_final * np.exp(-(e.t_final ** 2) / (8 * e.sigma ** 2))


def shape_derivative(shape: str, t: np.ndarray, e: EnvelopeSpec) -> np.ndarray:
    """d shape / dt (what tf.GradientTape gives EnvelopeDrag, pulse.py:173-178)."""
    t = np.asarray(t, dtype=np.float64)
    if shape in ("no_drive", "rect"):
        return np.zeros_like(t)
    if shape in ("gaussian_nonorm", "gaussian_sigma"):
        g = np.exp(-((t - e.t_final / 2) ** 2) / (2 * e.sigma ** 2)) * (-(t - e.t_final / 2) / e.sigma ** 2)
        if shape == "gaussian_sigma":
            offset = np.exp(-(e.t_final ** 2) / (8 * e.sigma ** 2))
            norm = np.sqrt(2 * np.pi * e.sigma ** 2) * erf(e.t_final / (np.sqrt(8) * e.sigma)) - e.t_final * offset
            g = g / norm
        return g
    if shape == "cosine":
        return 0.5 * np.sin(2 * np.pi * t / e.t_final) * 2 * np.pi / e.t_final
    if shape == "flattop":
        up, dn = (t - e.t_up) / e.risefall, (-t + e.t_down) / e.risefall
        c = 2 / np.sqrt(np.pi) / e.risefall
        return (c * np.exp(-up ** 2) * (1 + erf(dn)) - (1 + erf(up)) * c * np.exp(-dn ** 2)) / 4
    # the remaining shapes: central difference of the restated shape function (what tf.GradientTape returns, to 1e-9 relative;
    # the CUDA chain evaluates the analytic derivative)
    h = 1e-6 * max(e.t_final, 1e-12)
    if shape == "trapezoid":                       # piecewise linear: exact one-sided slopes (TF differentiates the taken branch)
        w = e.risefall * 2.5
        return np.where(t >= e.t_final - w, -1.0 / w, np.where(t <= w, 1.0 / w, 0.0))
    return (shape_values(shape, t + h, e) - shape_values(shape, t - h, e)) / (2 * h)


def compute_mask(ts: np.ndarray, t_final: float, t_end: float) -> np.ndarray:
    tf_ = min(t_final, t_end)
    dt = ts[1] - ts[0]
    return expit((ts / dt + 0.001) * 1e6) * expit((0.999 * tf_ - ts) / dt * 1e6)


def envelope_values(e: EnvelopeSpec, ts_off: np.ndarray, t_len: float) -> np.ndarray:
    """Complex envelope samples: Envelope.get_shape_values (mask * shape, optionally minus the value one
    sample before the start) and the DRAG quadrature of EnvelopeDrag."""
    mask = compute_mask(ts_off, e.t_final, t_len)
    env = mask * shape_values(e.shape, ts_off, e)        # complex for "pwc"
    if e.use_t_before:
        t_before = 2 * ts_off[0] - ts_off[1]
        env = mask * (shape_values(e.shape, ts_off, e) - shape_values(e.shape, np.array([t_before]), e)[0])
    if not e.drag:
        return env.astype(np.complex128)
    dt = ts_off[1] - ts_off[0]
    denv = mask * shape_derivative(e.shape, ts_off, e) * dt
    return env - 1j * denv * e.delta


def awg_signal(envelopes: Sequence[EnvelopeSpec], ts: np.ndarray, t_start: float = 0.0):
    """Instruction.get_awg_signal: sum over the channel's envelopes of amp * env * exp(i (xy - w_off t))."""
    signal = np.zeros_like(ts, dtype=np.complex128)
    for e in envelopes:
        ts_off = ts - t_start
        phase = e.xy_angle - e.freq_offset * ts_off
        signal = signal + e.amp * envelope_values(e, ts_off, e.t_final) * np.exp(1j * phase)
    return np.real(signal), np.imag(signal)


def crosstalk(signals: Dict[str, np.ndarray], channels: Sequence[str], matrix) -> Dict[str, np.ndarray]:
    """Crosstalk.process (c3/generator/devices.py:281-293): the listed channels mixed by the crosstalk matrix, the others
    untouched.  Pinned to the reference's known answers (test/test_crosstalk.py:7-27) in tests/test_signal_oracle.py."""
    stacked = np.stack([np.asarray(signals[ch], dtype=np.float64) for ch in channels])
    crossed = np.asarray(matrix, dtype=np.float64) @ stacked
    out = dict(signals)
    for i, ch in enumerate(channels):
        out[ch] = crossed[i]
    return out


def resize_nearest(x: np.ndarray, new_dim: int) -> np.ndarray:
    """tf.image.resize(method="nearest") along one axis: TF2 samples at floor((i + 0.5) * old / new)
    (half-pixel centres) -- pinned by the 2.4 GS/s -> 100 GS/s fixture (ratio 41.67)."""
    old = x.shape[0]
    idx = np.minimum(np.floor((np.arange(new_dim) + 0.5) * (old / new_dim)).astype(np.int64), old - 1)
    return x[idx]


def rise_function(rise_time: float, resolution: float, fft_variant: bool = False) -> np.ndarray:
    n_ts = int(np.floor(rise_time * resolution))
    ts = np.linspace(0.0, rise_time, n_ts)
    cen = (rise_time - 1 / resolution) / 2 if fft_variant else (rise_time + 1 / resolution) / 2
    sigma = rise_time / 4
    gauss = np.exp(-((ts - cen) ** 2) / (2 * sigma * sigma))
    offset = np.exp(-((-1 - cen) ** 2) / (2 * sigma * sigma))
    risefun = gauss - offset
    return risefun / np.sum(risefun)


def tf_convolve(sig: np.ndarray, resp: np.ndarray) -> np.ndarray:
    """FFT convolution truncated to the signal length: out[n] = sum_m resp[m] sig[n - m]."""
    n = len(sig) + len(resp)
    out = np.fft.ifft(np.fft.fft(sig, n) * np.fft.fft(resp, n))
    return out[: len(sig)]


def tf_convolve_legacy(sig: np.ndarray, resp: np.ndarray) -> np.ndarray:
    """Legacy variant: out[n] = sum_m resp[m] sig[n - 1 - m]."""
    s, r = len(sig), len(resp)
    n = s + 2 * r
    pad = np.concatenate([np.zeros(r), sig, np.zeros(r)])
    out = np.fft.ifft(np.fft.fft(pad, n) * np.fft.fft(resp, n))
    return out[r - 1: s + r - 1]


def response(i_sig, q_sig, rise_time: float, resolution: float, fft_variant: bool = False):
    rf = rise_function(rise_time, resolution, fft_variant)
    conv = tf_convolve if fft_variant else tf_convolve_legacy
    return np.real(conv(i_sig, rf)), np.real(conv(q_sig, rf))


def mixer(lo_i, lo_q, i_sig, q_sig):
    return lo_i * i_sig + lo_q * q_sig


def flux_tuning(signal, phi: float, phi_0: float, omega_0: float, anhar: float, d: Optional[float]):
    def factor(p):
        x = np.pi * p / phi_0
        if d is not None:
            return np.sqrt(np.sqrt(np.cos(x) ** 2 + d ** 2 * np.sin(x) ** 2))
        return np.sqrt(np.abs(np.cos(x)))

    def freq(p):
        return (omega_0 - anhar) * factor(p) + anhar
    return freq(phi + signal) - freq(phi)


@dataclass
class ChainSpec:
    """The standard device chain of one drive line."""
    sim_res: float = 100e9
    awg_res: float = 2e9
    rise_time: float = 0.3e-9
    response_fft: bool = False          # ResponseFFT instead of the legacy Response
    v2hz: float = 1e9                   # VoltsToHertz factor (with its 2 pi if the unit says so)
    flux: Optional[Dict] = None         # FluxTuning parameters instead of VoltsToHertz: phi, phi_0, omega_0, anhar, d


def generate_signal(envelopes: Sequence[EnvelopeSpec], omega_lo: float, t_start: float, t_end: float,
                    chain: ChainSpec, stages: Optional[dict] = None):
    """One channel of Generator.generate_signals (c3/generator/generator.py:172-229) for the standard chain.
    Returns (values [N], ts [N]); ``stages`` (if given) receives every intermediate signal."""
    ts = create_ts(t_start, t_end, chain.sim_res)
    lo_i, lo_q = lo_signal(ts, omega_lo)
    ts_awg = create_ts(t_start, t_end, chain.awg_res)
    awg_i, awg_q = awg_signal(envelopes, ts_awg, t_start)
    dac_i, dac_q = resize_nearest(awg_i, len(ts)), resize_nearest(awg_q, len(ts))
    resp_i, resp_q = response(dac_i, dac_q, chain.rise_time, chain.sim_res, chain.response_fft)
    mixed = mixer(lo_i, lo_q, resp_i, resp_q)
    if chain.flux is not None:
        values = flux_tuning(mixed, **chain.flux)
    else:
        values = mixed * chain.v2hz
    if stages is not None:
        stages.update(lo_i=lo_i, lo_q=lo_q, ts_awg=ts_awg, awg_i=awg_i, awg_q=awg_q, dac_i=dac_i, dac_q=dac_q,
                      resp_i=resp_i, resp_q=resp_q, mixed=mixed)
    return values, ts
