"""Duck-typed stand-ins for the reference objects the mirrored APIs consume (TensorFlow, and with it the
reference package, cannot be imported here).  Class NAMES match the reference's because the signal-chain
mirror dispatches on them exactly like a hjson config's ``c3type`` would."""
import numpy as np


class Quantity:
    """c3/c3objs.py Quantity: value * (2 pi if the unit says so)."""

    def __init__(self, value, unit=""):
        self.unit = unit
        self._v = np.asarray(value, dtype=np.float64) * (2 * np.pi if "2pi" in unit else 1.0)

    def get_value(self):
        return self._v


class _Dev:
    def __init__(self, name="", resolution=0.0, **params):
        self.name, self.resolution, self.params = name, resolution, params


class LO(_Dev):
    pass


class AWG(_Dev):
    pass


class DigitalToAnalog(_Dev):
    pass


class Response(_Dev):
    pass


class ResponseFFT(_Dev):
    pass


class Mixer(_Dev):
    pass


class VoltsToHertz(_Dev):
    pass


class FluxTuning(_Dev):
    pass


class Additive_Noise(_Dev):
    pass


class _Shape:
    def __init__(self, name):
        self.__name__ = name


class Envelope:
    def __init__(self, name, shape, params, use_t_before=False):
        self.name, self.shape, self.params, self.use_t_before = name, _Shape(shape), params, use_t_before


class EnvelopeDrag(Envelope):
    pass


class Carrier:
    def __init__(self, name, params):
        self.name, self.params = name, params


class Instruction:
    def __init__(self, name, t_start, t_end, channels):
        self.name, self.t_start, self.t_end = name, t_start, t_end
        self.comps = {c: {} for c in channels}
        self._options = {c: {} for c in channels}

    def add_component(self, comp, chan):
        self.comps[chan][comp.name] = comp
        self._options[chan][comp.name] = {}


def standard_chain(lo="LO", awg="AWG", dac="DigitalToAnalog", resp="Response", mixer="Mixer", out="VoltsToHertz"):
    chain = {lo: [], awg: [], dac: [awg]}
    if resp:
        chain[resp] = [dac]
        chain[mixer] = [lo, resp]
    else:
        chain[mixer] = [lo, dac]
    chain[out] = [mixer]
    return chain


def reference_generator_setup():
    """The objects of test/test_generator.py:21-118 of the reference."""
    sim_res, awg_res = 100e9, 2e9
    devices = {
        "LO": LO("lo", sim_res), "AWG": AWG("awg", awg_res), "DigitalToAnalog": DigitalToAnalog("dac", sim_res),
        "Response": Response("resp", sim_res, rise_time=Quantity(0.3e-9, "s")), "Mixer": Mixer("mixer"),
        "VoltsToHertz": VoltsToHertz("v_to_hz", V_to_Hz=Quantity(1e9, "Hz/V")),
    }
    t_final, sideband = 7e-9, 50e6
    env = Envelope("gauss", "gaussian_nonorm", {
        "amp": Quantity(0.5, "V"), "t_final": Quantity(t_final, "s"), "sigma": Quantity(t_final / 4, "s"),
        "xy_angle": Quantity(0.0, "rad"), "freq_offset": Quantity(-sideband - 3e6, "Hz 2pi"), "delta": Quantity(-1, "")})
    carr = Carrier("carrier", {"freq": Quantity(5e9 + sideband, "Hz 2pi"), "framechange": Quantity(0.0, "rad")})
    instr = Instruction("rx90p", 0.0, t_final, ["d1"])
    instr.add_component(env, "d1")
    instr.add_component(carr, "d1")
    return devices, {"d1": standard_chain()}, instr


def tunable_coupler_flux_setup():
    """The flux line ("TC") of test/test_tunable_coupler.py:158-262 (xy_angle already sign-flipped, :279-283)."""
    sim_res, awg_res, phi_0 = 100e9, 2.4e9, 10.0
    devices = {
        "lo": LO("lo", sim_res), "awg": AWG("awg", awg_res), "dac": DigitalToAnalog("dac", sim_res),
        "resp": Response("resp", sim_res, rise_time=Quantity(0.3e-9, "s")), "mixer": Mixer("mixer"),
        "fluxbias": FluxTuning("fluxbias", phi_0=Quantity(phi_0, "Wb"), phi=Quantity(phi_0 * 0.23, "Wb"),
                               omega_0=Quantity(8.1e9, "Hz 2pi"), d=Quantity(0.36, ""), anhar=Quantity(-286e6, "Hz 2pi")),
    }
    chain = standard_chain("lo", "awg", "dac", "resp", "mixer", "fluxbias")
    t = 100e-9
    env = Envelope("flux", "flattop", {
        "amp": Quantity(0.1 * phi_0, "V"), "t_final": Quantity(t, "s"), "t_up": Quantity(5e-9, "s"),
        "t_down": Quantity(t - 5e-9, "s"), "risefall": Quantity(5e-9, "s"), "freq_offset": Quantity(0.0, "Hz 2pi"),
        "xy_angle": Quantity(0.3590456701578104, "rad")})
    carr = Carrier("carrier", {"freq": Quantity(829e6, "Hz 2pi"), "framechange": Quantity(0.0, "rad")})
    instr = Instruction("crzp", 0.0, t, ["TC"])
    instr.add_component(env, "TC")
    instr.add_component(carr, "TC")
    return devices, {"TC": chain}, instr


# ---------------------------------------------------------------------------------------------------------------------
# Model / Generator / ParameterMap stand-ins for the propagator path (what c3/libraries/propagation.py:282-321 and
# c3/experiment.py:440-534 touch), built from arrays
# ---------------------------------------------------------------------------------------------------------------------

class FakeAWG:
    def __init__(self, amp=0.0):
        self.amp = amp

    def get_average_amp(self):
        return self.amp, self.amp


class TableGenerator:
    """generate_signals(instr) looks the instruction's fields up in a table: {instr.name: {chan: {"values","ts"}}}."""

    def __init__(self, table, avg_amp=0.0):
        self.table = table
        self.devices = {"awg": FakeAWG(avg_amp)}
        self.calls = 0

    def generate_signals(self, instr):
        self.calls += 1
        return self.table[instr.name]


class ArrayModel:
    """Model stand-in: drift / control Hamiltonians (dict by channel), collapse operators, optional excitation cutter,
    optional explicit Hamiltonian list, frame rotation and dephasing channel from the oracle's restatement."""

    def __init__(self, h0, hks, col_ops=(), dims=None, lindbladian=False, max_excitations=0, hlist=None, use_FR=False,
                 dephasing_strength=0.0, line_to_index=None, exact_frames=False):
        from oracle import c3_oracle as orc
        from oracle import c3_model_oracle as mo
        self._orc = orc
        self.h0, self.hks, self.col_ops = np.asarray(h0), {k: np.asarray(v) for k, v in hks.items()}, [np.asarray(c) for c in col_ops]
        self.dims = list(dims) if dims is not None else [self.h0.shape[0]]
        self.tot_dim = int(np.prod(self.dims))
        self.lindbladian, self.max_excitations = lindbladian, max_excitations
        self.controllability = True
        self.use_FR, self.dephasing_strength = use_FR, dephasing_strength
        self.hlist = hlist
        self.ex_cutter = orc.make_ex_cutter(self.dims, max_excitations) if max_excitations else None
        self.ann_opers = mo.annihilators(self.dims)
        self.line_to_index = line_to_index or {}
        # frame rotation / dephasing exponentials: the oracle's restatement of tf.linalg.expm (default: what the reference
        # computes, truncation error of TensorFlow's floor-scaled Pade included) or scipy's expm (accurate to rounding)
        self.exact_frames = exact_frames

    def _cut(self, x):
        return self._orc.cut_excitations(self.ex_cutter, x) if self.max_excitations else x

    def get_Hamiltonians(self):
        return self._cut(self.h0), {k: self._cut(v) for k, v in self.hks.items()}

    def get_Hamiltonian(self, signal=None):
        return np.stack([self._cut(h) for h in self.hlist])

    def get_Lindbladians(self):
        return list(self.col_ops)

    def _expm(self):
        if not self.exact_frames:
            return None
        import scipy.linalg
        return scipy.linalg.expm

    def get_Frame_Rotation(self, t_final, freqs, framechanges):
        return self._orc.frame_rotation(self.ann_opers, self.line_to_index, t_final, freqs, framechanges, expm=self._expm())

    def get_dephasing_channel(self, t_final, amps):
        return self._orc.dephasing_channel(self.ann_opers, self.line_to_index, t_final, amps, self.dephasing_strength, expm=self._expm())


class PMap:
    def __init__(self, model, generator, instructions):
        self.model, self.generator, self.instructions = model, generator, instructions


def drive_instruction(name, t_end, lines, freq=5e9, framechange=0.0, amp=0.5, freq_offset=0.0):
    """An instruction with one envelope + carrier per line (parameters as Quantities, c3/signal/gates.py)."""
    instr = Instruction(name, 0.0, t_end, list(lines))
    for i, line in enumerate(lines):
        f = freq[i] if isinstance(freq, (list, tuple)) else freq
        fc = framechange[i] if isinstance(framechange, (list, tuple)) else framechange
        instr.add_component(Envelope("gauss", "gaussian_nonorm", {
            "amp": Quantity(amp, "V"), "t_final": Quantity(t_end, "s"), "sigma": Quantity(t_end / 4, "s"),
            "xy_angle": Quantity(0.0, "rad"), "freq_offset": Quantity(freq_offset, "Hz 2pi"), "delta": Quantity(0.0, "")}), line)
        instr.add_component(Carrier("carrier", {"freq": Quantity(f, "Hz 2pi"), "framechange": Quantity(fc, "rad")}), line)
    return instr


class LONoise(_Dev):
    pass


class DC_Noise(_Dev):
    pass


class Pink_Noise(_Dev):
    pass


class DC_Offset(_Dev):
    pass


class Crosstalk(_Dev):
    """c3/generator/devices.py:225-263: channels + crosstalk_matrix."""

    def __init__(self, name="crosstalk", channels=None, crosstalk_matrix=None):
        super().__init__(name, crosstalk_matrix=crosstalk_matrix)
        self.crossed_channels = channels


def noisy_generator_setup(awg_amp=0.0, dc_amp=0.0, pink_amp=0.0, add_amp=0.0, lo_perc=0.0, bfl_num=15, offset=0.0):
    """The drive line of test/noise_exp_2.hjson of the reference (without its high-pass filter): AWG -> AWGNoise ->
    DigitalToAnalog -> Response -> Mixer -> DCNoise -> PinkNoise -> DCOffset -> VoltsToHertz, plus an LO noise device."""
    devices, _, instr = reference_generator_setup()
    devices.update({
        "LONoise": LONoise("lo_noise", 100e9, noise_perc=Quantity(lo_perc, "")),
        "AWGNoise": Additive_Noise("awg_noise", 100e9, noise_amp=Quantity(awg_amp, "V")),
        "AddNoise": Additive_Noise("add_noise", 100e9, noise_amp=Quantity(add_amp, "V")),
        "DCNoise": DC_Noise("dc_noise", 100e9, noise_amp=Quantity(dc_amp, "V")),
        "PinkNoise": Pink_Noise("pink_noise", 100e9, noise_amp=Quantity(pink_amp, "V"), bfl_num=Quantity(bfl_num, "")),
        "DCOffset": DC_Offset("dc_offset", 100e9, offset_amp=Quantity(offset, "V")),
    })
    chain = {"LO": [], "LONoise": ["LO"], "AWG": [], "AWGNoise": ["AWG"], "DigitalToAnalog": ["AWGNoise"],
             "Response": ["DigitalToAnalog"], "Mixer": ["LONoise", "Response"], "AddNoise": ["Mixer"], "DCNoise": ["AddNoise"],
             "PinkNoise": ["DCNoise"], "DCOffset": ["PinkNoise"], "VoltsToHertz": ["DCOffset"]}
    return devices, {"d1": chain}, instr
