"""Mirror of ``c3.generator.generator.Generator`` for the standard device chain, on the B200 engine
(SURVEY.md section 8f, row f-2): pulse parameters -> control fields, one kernel for a whole batch of parameter
samples, output already in the ``[B,K,N]`` layout the propagator kernels read (no host round trip).

  Generator(devices, chains).generate_signals(instr) -> {chan: {"values": [N], "ts": [N]}}
        c3/generator/generator.py:172-229 (same call, same dictionary; tensors are torch CUDA float64)
  Generator.generate_signals_batch(instr, samples) -> signals [B,K,N], ts [N]
        the batch axis the reference's optimisers loop over serially

Model / device / instruction objects are duck-typed exactly as the reference uses them:
  devices[name]: class name in {LO, AWG, DigitalToAnalog, Response, ResponseFFT, Mixer, VoltsToHertz, FluxTuning,
                 LONoise, Additive_Noise, DC_Noise, Pink_Noise, DC_Offset},
                 ``.resolution``, ``.params[key].get_value()``
  chains[chan]:  {dev: [sources]}  -- must be the standard topology (LO, AWG -> DAC -> [Response] -> Mixer -> out)
  instr.t_start, instr.t_end, instr.comps[chan][name]: Envelope (``.shape.__name__``, ``.params``) or Carrier
The noise devices LONoise, Additive_Noise (behind the AWG or the mixer), DC_Noise, Pink_Noise and DC_Offset are part of the
kernel (one independent realisation per batch row: the Monte-Carlo trajectory axis).  Anything else (crosstalk, arbitrary
filters, other envelope shapes) raises ``C3:ERROR`` -- those chains stay on the reference's CPU path; there is no silent
fallback.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import engine

SHAPE_IDS = {"no_drive": 0, "rect": 1, "gaussian_nonorm": 2, "gaussian_sigma": 3, "cosine": 4, "flattop": 5, "trapezoid": 6,
             "flattop_risefall": 7, "gaussian_der_nonorm": 8, "gaussian_der": 9, "drag_sigma": 10, "drag_der": 11,
             # extended shapes (csrc/signal_chain.cuh SHAPE_FIRST_EXT...): grid-defined or parametrised by arrays, forward only
             "flattop_cut": 12, "flattop_cut_center": 13, "flattop_variant": 14, "cosine_flattop": 15, "delta_pulse": 16, "pwc": 17,
             "pwc_shape": 18, "pwc_symmetric": 19, "pwc_shape_plateau": 20, "fourier_sin": 21, "fourier_cos": 22,
             "slepian_fourier": 23}
FIRST_EXTENDED_SHAPE = 12
#: extended shapes whose extra SCALAR parameter rides in the sigma column of the envelope row
SIGMA_SLOT = {"flattop_cut_center": "width", "flattop_variant": "ramp", "cosine_flattop": "t_rise"}


def _np_value(q) -> np.ndarray:
    """Quantity-like -> real numpy array of its shape."""
    if hasattr(q, "get_value"):
        q = q.get_value()
    if hasattr(q, "numpy"):
        q = q.numpy()
    return np.real(np.asarray(q))


def _arr(q) -> np.ndarray:
    """Quantity-like -> 1-d float64 array."""
    if hasattr(q, "get_value"):
        q = q.get_value()
    if hasattr(q, "numpy"):
        q = q.numpy()
    return np.real(np.asarray(q)).astype(np.float64).reshape(-1)


def shape_table_row(shape: str, params) -> np.ndarray:
    """The array parameters of an extended shape in the order the kernel reads them (csrc/signal_chain.cuh, next to the shape
    ids); empty for the shapes that have none."""
    P = params
    if shape == "delta_pulse":
        t = _arr(P["t_sig"])
        return np.concatenate([[len(t)], t])
    if shape == "pwc":
        i, q = _arr(P["inphase"]), _arr(P["quadrature"])
        if len(i) != len(q):
            raise Exception("C3:ERROR: pwc envelope: inphase and quadrature differ in length.")
        return np.concatenate([[len(i)], i, q])
    if shape in ("pwc_shape", "pwc_symmetric", "pwc_shape_plateau"):
        y = _arr(P["inphase"])
        width = _val(P["width"]) if (shape == "pwc_shape_plateau" and "width" in P) else -1.0
        return np.concatenate([[len(y), _val(P["t_bin_start"]), _val(P["t_bin_end"]), width], y])
    if shape == "fourier_sin":
        a, f, ph = _arr(P["amps"]), _arr(P["freqs"]), _arr(P["phases"])
        if not (len(a) == len(f) == len(ph)):
            raise Exception("C3:ERROR: fourier_sin envelope: amps, freqs and phases differ in length.")
        return np.concatenate([[len(a)], a, f, ph])
    if shape == "fourier_cos":
        a, f = _arr(P["amps"]), _arr(P["freqs"])
        if len(a) != len(f):
            raise Exception("C3:ERROR: fourier_cos envelope: amps and freqs differ in length.")
        return np.concatenate([[len(a)], a, f])
    if shape == "slepian_fourier":
        c = _arr(P["fourier_coeffs"])
        sc = _arr(P["sin_coeffs"]) if "sin_coeffs" in P else np.zeros(0)
        rf = _val(P["risefall"]) if "risefall" in P else -1.0
        return np.concatenate([[_val(P["width"]), _val(P["offset"]), rf, len(c)], c, [len(sc)], sc])
    return np.zeros(0)
#: shapes of c3/libraries/envelopes.py that are another shape with one parameter fixed: name -> (kernel shape, parameter, rule)
SHAPE_ALIASES = {
    "gaussian": ("gaussian_sigma", "sigma", lambda p: p["t_final"] / 6),                 # envelopes.py:399-417
    "drag": ("drag_sigma", "sigma", lambda p: p["t_final"] / 4),                         # :533-542
    "flattop_risefall_1ns": ("flattop_risefall", "risefall", lambda p: 1e-9 + 0 * p["t_final"]),   # :366-370
}
ENV_KEYS = ("amp", "t_final", "sigma", "xy_angle", "freq_offset", "delta", "t_up", "t_down", "risefall")
ENV_DEFAULTS = {"amp": 0.0, "t_final": 0.0, "sigma": 1.0, "xy_angle": 0.0, "freq_offset": 0.0, "delta": 0.0,
                "t_up": 0.0, "t_down": 0.0, "risefall": 1.0}
CHAIN_KEYS = ("sim_res", "awg_res", "rise_time", "resp_kind", "out_kind", "v2hz", "phi", "phi_0", "omega_0", "anhar", "d")


def _val(q) -> float:
    """Quantity-like -> float (c3/c3objs.py:247-256 get_value)."""
    if hasattr(q, "get_value"):
        q = q.get_value()
    if hasattr(q, "numpy"):
        q = q.numpy()
    return float(np.real(np.asarray(q).reshape(-1)[0]))


def _cls(obj) -> str:
    return type(obj).__name__


class Generator:
    """Generator, creates signal from digital to what arrives to the chip (c3/generator/generator.py:16-60)."""

    def __init__(self, devices: dict = None, chains: dict = None, resolution: float = 0.0, callback=None):
        self.devices = devices or {}
        self.chains = chains or {}
        self.resolution = resolution
        self.callback = callback
        self.seed = 0                 # noise realisations are a function of (seed, number of calls so far, batch row, line)
        self.noise_draws = 0
        self._specs = {chan: self._chain_spec(chan) for chan in self.chains}

    # -- chain topology -> the 11 numbers the kernel needs (+ 7 for the noise devices) -------------------------------------
    SIGNAL_NOISE = ("Additive_Noise", "DC_Noise", "Pink_Noise", "DC_Offset")

    def _chain_spec(self, chan: str) -> Dict[str, float]:
        """Walk the chain of one drive line and check that it is the topology the kernel implements:

            LO -> [LONoise] ----------------------------------------------.
            AWG -> [Additive_Noise] -> DigitalToAnalog -> [Response | ResponseFFT] -> Mixer
                -> {Additive_Noise, DC_Noise, Pink_Noise, DC_Offset}* -> VoltsToHertz | FluxTuning

        (noise devices where test/noise_exp_2.hjson of the reference puts them; at most one of a kind per position)."""
        chain = self.chains[chan]
        cls = {dev_id: _cls(self.devices[dev_id]) for dev_id in chain}

        def fail(msg):
            raise Exception(f"C3:ERROR: chain '{chan}' is not a topology of the on-device signal chain: {msg}.")

        def only(kind):
            ids = [d for d, c in cls.items() if c == kind]
            if len(ids) != 1:
                raise Exception(f"C3:ERROR: chain '{chan}' needs exactly one {kind} device.")
            return ids[0]

        def consumer(dev_id):
            users = [d for d, src in chain.items() if dev_id in src]
            if len(users) != 1:
                fail(f"'{dev_id}' must feed exactly one device")
            return users[0]

        lo, awg, mixer = only("LO"), only("AWG"), only("Mixer")
        if list(chain[lo]) or list(chain[awg]):
            fail("LO and AWG are sources")
        spec = dict(rise_time=0.0, resp_kind=0.0, out_kind=0.0, v2hz=1.0, phi=0.0, phi_0=1.0, omega_0=0.0, anhar=0.0, d=float("nan"))
        noise = {k: 0.0 for k in engine.NOISE_KEYS}
        noise["bfl_num"] = 5.0
        spec["noise_devices"] = {}
        # LO branch
        cur = consumer(lo)
        if cls[cur] == "LONoise":
            noise["lo_perc"] = _val(self.devices[cur].params["noise_perc"])
            spec["noise_devices"]["lo"] = cur
            cur = consumer(cur)
        if cur != mixer:
            fail(f"'{cur}' between the LO and the mixer")
        lo_end = [d for d in chain[mixer] if d == lo or cls.get(d) == "LONoise"]
        # signal branch
        cur = consumer(awg)
        if cls[cur] == "Additive_Noise":
            noise["awg_amp"] = _val(self.devices[cur].params["noise_amp"])
            spec["noise_devices"]["awg"] = cur
            cur = consumer(cur)
        if cls[cur] != "DigitalToAnalog":
            fail(f"'{cur}' ({cls[cur]}) where the DigitalToAnalog converter belongs")
        dac = cur
        cur = consumer(cur)
        if cls[cur] in ("Response", "ResponseFFT"):
            spec["rise_time"] = _val(self.devices[cur].params["rise_time"])
            spec["resp_kind"] = 1.0 if cls[cur] == "Response" else 2.0
            cur = consumer(cur)
        if cur != mixer or len(chain[mixer]) != 2 or len(lo_end) != 1 or list(chain[mixer])[0] != lo_end[0]:
            fail(f"the mixer must take [LO branch, signal branch], got {list(chain[mixer])}")
        cur = consumer(mixer)
        seen = set()
        while cls[cur] in self.SIGNAL_NOISE:
            kind = cls[cur]
            if kind in seen:
                fail(f"more than one {kind} behind the mixer")
            seen.add(kind)
            par = self.devices[cur].params
            if kind == "Additive_Noise":
                noise["add_amp"] = _val(par["noise_amp"])
                spec["noise_devices"]["add"] = cur
            elif kind == "DC_Noise":
                noise["dc_amp"] = _val(par["noise_amp"])
                spec["noise_devices"]["dc"] = cur
            elif kind == "Pink_Noise":
                noise["pink_amp"] = _val(par["noise_amp"])
                if "bfl_num" in par:
                    noise["bfl_num"] = float(int(_val(par["bfl_num"])))
                spec["noise_devices"]["pink"] = cur
            else:
                noise["dc_offset"] = _val(par["offset_amp"])
            cur = consumer(cur)
        if cls[cur] not in ("VoltsToHertz", "FluxTuning") or any(cur in src for src in chain.values()):
            fail(f"'{cur}' ({cls[cur]}) where the VoltsToHertz / FluxTuning output belongs")
        known = {lo, awg, dac, mixer, cur} | set(spec["noise_devices"].values())
        extra = [d for d in chain if d not in known and cls[d] not in ("Response", "ResponseFFT", "DC_Offset")]
        if extra:
            raise Exception(f"C3:ERROR: devices {sorted(extra)} in chain '{chan}' are not part of the on-device signal chain.")
        sim_res = float(self.devices[dac].resolution)
        if float(self.devices[lo].resolution) != sim_res:
            raise Exception("C3:ERROR: LO and DigitalToAnalog must share the simulation resolution.")
        spec.update(sim_res=sim_res, awg_res=float(self.devices[awg].resolution))
        if spec["awg_res"] > sim_res:
            raise Exception("C3:ERROR: the AWG grid must not be finer than the simulation grid.")
        if cls[cur] == "VoltsToHertz":
            spec["v2hz"] = _val(self.devices[cur].params["V_to_Hz"])
        else:
            par = self.devices[cur].params
            spec.update(out_kind=1.0, phi=_val(par["phi"]), phi_0=_val(par["phi_0"]), omega_0=_val(par["omega_0"]),
                        anhar=_val(par["anhar"]))
            if "d" in par:
                spec["d"] = _val(par["d"])
        spec["noise"] = noise
        spec["noisy"] = any(noise[k] != 0.0 for k in engine.NOISE_KEYS if k != "bfl_num") or bool(spec["noise_devices"])
        return spec

    # -- instruction -> envelope table ----------------------------------------------------------------------------
    @staticmethod
    def _channel_components(instr, chan):
        envs, carrier = [], None
        for name, comp in instr.comps[chan].items():
            cls = _cls(comp)
            if cls == "Carrier":
                carrier = comp
            elif cls in ("Envelope", "EnvelopeDrag"):
                shape = getattr(comp.shape, "__name__", str(comp.shape))
                if shape not in SHAPE_IDS and shape not in SHAPE_ALIASES:
                    raise Exception(f"C3:ERROR: envelope shape '{shape}' is not available in the on-device signal chain.")
                opts = getattr(instr, "_options", {}).get(chan, {}).get(name, {})
                if opts:
                    raise Exception("C3:ERROR: component options (delay, trigger_comp, t_final_cut) are not supported on device.")
                fl = (1 if cls == "EnvelopeDrag" else 0) | (2 if getattr(comp, "use_t_before", False) else 0)
                if fl and SHAPE_IDS.get(shape, 0) >= FIRST_EXTENDED_SHAPE:
                    raise Exception(f"C3:ERROR: envelope shape '{shape}' is available on the device without DRAG / use_t_before only.")
                envs.append((name, comp, shape, fl))
            else:
                raise Exception(f"C3:ERROR: component type '{cls}' is not available in the on-device signal chain.")
        if carrier is None:
            raise Exception(f"C3:Error: Probably no carrier proviced for {chan}")
        return envs, carrier

    def _tables(self, instr, samples: Optional[Dict] = None):
        chans = list(instr.comps.keys())
        comps = {c: self._channel_components(instr, c) for c in chans}
        K = len(chans)
        E = max(1, max(len(comps[c][0]) for c in chans))
        B = 1
        if samples:
            B = len(next(iter(samples.values())))
        env = np.zeros((B, K, E, len(ENV_KEYS)))
        shape = -np.ones((K, E), dtype=np.int32)
        flags = np.zeros((K, E), dtype=np.int32)
        lo = np.zeros((B, K))
        chain = np.zeros((K, len(CHAIN_KEYS)))
        rows = {}                                   # (k, e) -> array parameters of an extended shape
        for k, c in enumerate(chans):
            if c not in self._specs:
                raise Exception(f"C3:ERROR: no signal chain for channel '{c}'.")
            chain[k] = [self._specs[c][key] for key in CHAIN_KEYS]
            envs, carrier = comps[c]
            lo[:, k] = _val(carrier.params["freq"])
            if samples and (c, "carrier", "freq") in samples:
                lo[:, k] = np.asarray(samples[(c, "carrier", "freq")], dtype=np.float64)
            for e, (name, comp, sname, fl) in enumerate(envs):
                alias = SHAPE_ALIASES.get(sname)
                shape[k, e], flags[k, e] = SHAPE_IDS[alias[0] if alias else sname], fl
                for i, key in enumerate(ENV_KEYS):
                    env[:, k, e, i] = _val(comp.params[key]) if key in comp.params else ENV_DEFAULTS[key]
                    if samples and (c, name, key) in samples:
                        env[:, k, e, i] = np.asarray(samples[(c, name, key)], dtype=np.float64)
                if alias:       # e.g. "gaussian" = gaussian_sigma with sigma = t_final / 6 (per sample, after the overrides)
                    cols = {key: env[:, k, e, i] for i, key in enumerate(ENV_KEYS)}
                    env[:, k, e, ENV_KEYS.index(alias[1])] = alias[2](cols)
                try:
                    if sname in SIGMA_SLOT:             # width / ramp / t_rise travel in the sigma column
                        key = SIGMA_SLOT[sname]
                        env[:, k, e, ENV_KEYS.index("sigma")] = _val(comp.params[key])
                        if samples and (c, name, key) in samples:
                            env[:, k, e, ENV_KEYS.index("sigma")] = np.asarray(samples[(c, name, key)], dtype=np.float64)
                    if shape[k, e] >= FIRST_EXTENDED_SHAPE:
                        row = shape_table_row(sname, comp.params)
                        if len(row):
                            rows[(k, e)] = row
                except KeyError as err:
                    raise Exception(f"C3:ERROR: envelope '{name}' of shape '{sname}' lacks the parameter {err}.") from None
        # array parameters: one zero-padded row per envelope, shared by the batch (read by engine.generate_signals via self._table)
        self._table = None
        if rows:
            self._table = np.zeros((K, E, max(len(r) for r in rows.values())))
            for (k, e), r in rows.items():
                self._table[k, e, :len(r)] = r
        if len({chain[k, 0] for k in range(K)}) != 1:
            raise Exception("C3:ERROR: all channels of an instruction must share the simulation resolution.")
        return chans, env, shape, flags, lo, chain

    def _noise_table(self, chans):
        """[K,7] noise parameters of the instruction's drive lines, or None for a noise-free chain.  Re-read from the device
        objects on every call: an optimiser may have changed a noise amplitude (test/test_noise.py:93-104)."""
        self._specs = {chan: self._chain_spec(chan) for chan in self.chains}
        if not any(self._specs[c]["noisy"] for c in chans):
            return None
        return np.array([[self._specs[c]["noise"][k] for k in engine.NOISE_KEYS] for c in chans])

    def _publish_noise(self, chans, traces, chain) -> None:
        """Leave the realised noise on the device objects as the reference does (Device.signal["noise"],
        "noise-inphase" / "noise-quadrature" for the AWG noise; c3/generator/devices.py:975-996): first batch row."""
        t = {name: i for i, name in enumerate(engine.NOISE_TRACES)}
        for k, chan in enumerate(chans):
            where = self._specs[chan]["noise_devices"]
            span = traces.shape[-1] / float(chain[k, 0])
            n_awg = int(span * float(chain[k, 1]) + 1e-9)
            for slot, dev_id in where.items():
                dev = self.devices[dev_id]
                if slot == "awg":
                    dev.signal = {"noise-inphase": traces[0, k, t["awg_i"], :n_awg], "noise-quadrature": traces[0, k, t["awg_q"], :n_awg]}
                elif slot == "lo":
                    dev.signal = {"noise": (traces[0, k, t["lo_cos"]], traces[0, k, t["lo_sin"]])}
                else:
                    dev.signal = {"noise": traces[0, k, t[slot]]}

    # -- public API ---------------------------------------------------------------------------------------------------
    def generate_signals(self, instr) -> dict:
        """Perform the signal chain for a specified instruction, including local oscillator, AWG generation
        and IQ mixing (c3/generator/generator.py:172-229)."""
        sig, ts, chans = self._run(instr, None)
        gen_signal = {}
        for k, chan in enumerate(chans):
            gen_signal[chan] = {"values": sig[0, k], "ts": ts}
            if self.callback:
                self.callback(chan, "out", gen_signal[chan])
        return gen_signal

    def generate_signals_batch(self, instr, samples: Dict):
        """``samples[(channel, component, parameter)] = [B] values`` (as ``get_value()`` would return them)
        override the instruction's parameters per batch element.  Returns (signals [B,K,N] CUDA float64 in the
        channel order of ``instr.comps``, ts [N])."""
        sig, ts, _ = self._run(instr, samples)
        return sig, ts

    def _run(self, instr, samples):
        chans, env, shape, flags, lo, chain = self._tables(instr, samples)
        noise = self._noise_table(chans)
        if noise is None:
            sig = engine.generate_signals(env, shape, flags, lo, chain, float(instr.t_start), float(instr.t_end),
                                          env_table=self._table)
        else:
            # a fresh realisation per call (the reference draws from numpy's global generator on every call); every batch
            # row is an independent realisation of the same call
            self.noise_draws += 1
            seed = (int(self.seed) + 0x9E3779B97F4A7C15 * self.noise_draws) & 0xFFFFFFFFFFFFFFFF
            sig, traces = engine.generate_signals(env, shape, flags, lo, chain, float(instr.t_start), float(instr.t_end),
                                                  noise=noise, seed=seed, return_noise=True, env_table=self._table)
            self._publish_noise(chans, traces, chain)
        # the Crosstalk post-processing of the reference (c3/generator/generator.py:229-234): a device NAMED "crosstalk" mixes the
        # finished lines of its channels
        if "crosstalk" in self.devices:
            xt = self.devices["crosstalk"]
            crossed = list(xt.crossed_channels)
            missing = [c for c in crossed if c not in chans]
            if missing:
                raise Exception(f"C3:ERROR: crosstalk channels {missing} are not driven by this instruction.")
            engine.crosstalk(sig, [chans.index(c) for c in crossed], np.asarray(_np_value(xt.params["crosstalk_matrix"]), dtype=np.float64))
        N = sig.shape[-1]
        dt = 1.0 / chain[0, 0]
        ts = torch.as_tensor(np.linspace(float(instr.t_start) + dt / 2, float(instr.t_end) - dt / 2, N)).to(sig.device)
        return sig, ts, chans
