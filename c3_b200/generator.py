"""Mirror of ``c3.generator.generator.Generator`` for the standard device chain, on the B200 engine
(SURVEY.md section 8f, row f-2): pulse parameters -> control fields, one kernel for a whole batch of parameter
samples, output already in the ``[B,K,N]`` layout the propagator kernels read (no host round trip).

  Generator(devices, chains).generate_signals(instr) -> {chan: {"values": [N], "ts": [N]}}
        c3/generator/generator.py:172-229 (same call, same dictionary; tensors are torch CUDA float64)
  Generator.generate_signals_batch(instr, samples) -> signals [B,K,N], ts [N]
        the batch axis the reference's optimisers loop over serially

Model / device / instruction objects are duck-typed exactly as the reference uses them:
  devices[name]: class name in {LO, AWG, DigitalToAnalog, Response, ResponseFFT, Mixer, VoltsToHertz, FluxTuning},
                 ``.resolution``, ``.params[key].get_value()``
  chains[chan]:  {dev: [sources]}  -- must be the standard topology (LO, AWG -> DAC -> [Response] -> Mixer -> out)
  instr.t_start, instr.t_end, instr.comps[chan][name]: Envelope (``.shape.__name__``, ``.params``) or Carrier
Anything else (noise devices, crosstalk, arbitrary filters, other envelope shapes) raises ``C3:ERROR`` -- those
chains stay on the reference's CPU path; there is no silent fallback.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import engine

SHAPE_IDS = {"no_drive": 0, "rect": 1, "gaussian_nonorm": 2, "gaussian_sigma": 3, "cosine": 4, "flattop": 5}
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
        self._specs = {chan: self._chain_spec(chan) for chan in self.chains}

    # -- chain topology -> the 11 numbers the kernel needs ----------------------------------------------------------
    def _chain_spec(self, chan: str) -> Dict[str, float]:
        chain = self.chains[chan]
        by_cls: Dict[str, List[str]] = {}
        for dev_id in chain:
            by_cls.setdefault(_cls(self.devices[dev_id]), []).append(dev_id)
        allowed = {"LO", "AWG", "DigitalToAnalog", "Response", "ResponseFFT", "Mixer", "VoltsToHertz", "FluxTuning"}
        extra = set(by_cls) - allowed
        if extra:
            raise Exception(f"C3:ERROR: devices {sorted(extra)} in chain '{chan}' are not part of the on-device signal chain.")
        for need in ("LO", "AWG", "DigitalToAnalog", "Mixer"):
            if len(by_cls.get(need, [])) != 1:
                raise Exception(f"C3:ERROR: chain '{chan}' needs exactly one {need} device.")
        if ("VoltsToHertz" in by_cls) == ("FluxTuning" in by_cls):
            raise Exception(f"C3:ERROR: chain '{chan}' needs either a VoltsToHertz or a FluxTuning output device.")
        lo, awg, dac, mixer = (by_cls[c][0] for c in ("LO", "AWG", "DigitalToAnalog", "Mixer"))
        resp_cls = "Response" if "Response" in by_cls else ("ResponseFFT" if "ResponseFFT" in by_cls else None)
        out_cls = "VoltsToHertz" if "VoltsToHertz" in by_cls else "FluxTuning"
        out = by_cls[out_cls][0]
        want = {lo: [], awg: [], dac: [awg], mixer: None, out: [mixer]}
        if resp_cls:
            resp = by_cls[resp_cls][0]
            want[resp] = [dac]
            want[mixer] = [lo, resp]
        else:
            want[mixer] = [lo, dac]
        for dev_id, src in want.items():
            if list(chain[dev_id]) != src:
                raise Exception(f"C3:ERROR: chain '{chan}' is not the standard topology at '{dev_id}': {chain[dev_id]} != {src}.")
        sim_res = float(self.devices[dac].resolution)
        if float(self.devices[lo].resolution) != sim_res:
            raise Exception("C3:ERROR: LO and DigitalToAnalog must share the simulation resolution.")
        spec = dict(sim_res=sim_res, awg_res=float(self.devices[awg].resolution), rise_time=0.0, resp_kind=0.0,
                    out_kind=0.0, v2hz=1.0, phi=0.0, phi_0=1.0, omega_0=0.0, anhar=0.0, d=float("nan"))
        if spec["awg_res"] > sim_res:
            raise Exception("C3:ERROR: the AWG grid must not be finer than the simulation grid.")
        if resp_cls:
            spec["rise_time"] = _val(self.devices[resp].params["rise_time"])
            spec["resp_kind"] = 1.0 if resp_cls == "Response" else 2.0
        if out_cls == "VoltsToHertz":
            spec["v2hz"] = _val(self.devices[out].params["V_to_Hz"])
        else:
            par = self.devices[out].params
            spec.update(out_kind=1.0, phi=_val(par["phi"]), phi_0=_val(par["phi_0"]), omega_0=_val(par["omega_0"]),
                        anhar=_val(par["anhar"]))
            if "d" in par:
                spec["d"] = _val(par["d"])
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
                if shape not in SHAPE_IDS:
                    raise Exception(f"C3:ERROR: envelope shape '{shape}' is not available in the on-device signal chain.")
                opts = getattr(instr, "_options", {}).get(chan, {}).get(name, {})
                if opts:
                    raise Exception("C3:ERROR: component options (delay, trigger_comp, t_final_cut) are not supported on device.")
                envs.append((name, comp, SHAPE_IDS[shape], (1 if cls == "EnvelopeDrag" else 0)
                             | (2 if getattr(comp, "use_t_before", False) else 0)))
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
        for k, c in enumerate(chans):
            if c not in self._specs:
                raise Exception(f"C3:ERROR: no signal chain for channel '{c}'.")
            chain[k] = [self._specs[c][key] for key in CHAIN_KEYS]
            envs, carrier = comps[c]
            lo[:, k] = _val(carrier.params["freq"])
            if samples and (c, "carrier", "freq") in samples:
                lo[:, k] = np.asarray(samples[(c, "carrier", "freq")], dtype=np.float64)
            for e, (name, comp, sid, fl) in enumerate(envs):
                shape[k, e], flags[k, e] = sid, fl
                for i, key in enumerate(ENV_KEYS):
                    env[:, k, e, i] = _val(comp.params[key]) if key in comp.params else ENV_DEFAULTS[key]
                    if samples and (c, name, key) in samples:
                        env[:, k, e, i] = np.asarray(samples[(c, name, key)], dtype=np.float64)
        if len({chain[k, 0] for k in range(K)}) != 1:
            raise Exception("C3:ERROR: all channels of an instruction must share the simulation resolution.")
        return chans, env, shape, flags, lo, chain

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
        sig = engine.generate_signals(env, shape, flags, lo, chain, float(instr.t_start), float(instr.t_end))
        N = sig.shape[-1]
        dt = 1.0 / chain[0, 0]
        ts = torch.as_tensor(np.linspace(float(instr.t_start) + dt / 2, float(instr.t_end) - dt / 2, N)).to(sig.device)
        return sig, ts, chans
