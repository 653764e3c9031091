"""Re-export the reference's own golden vectors for the PWC propagator path as .npz fixtures.

Run HERE (the build container), where /root/reference exists:

    python tests/golden/make_golden.py

The reference pickles hold ``tf.EagerTensor`` objects; TensorFlow is not installable in this
image, so they are read with a stub unpickler.  Nothing is computed
except the dressed collapse operators of the two-qubit chip, which the pickle does not carry
and which are rebuilt from test/conftest.py:259-320 via oracle/c3_model_oracle.py.

Sources (relative to the reference checkout):
  test/two_qubit_data.pickle      <- test/test_two_qubits.py:46-62,193-213
  test/transmon_expanded.pickle   <- test/test_transmon_expanded.py:252-283
  test/test_tf_utils.pickle       <- test/test_tf_utils.py:81-111
  test/tunable_coupler_data.pickle <- test/test_tunable_coupler.py:31-157,399-409 (d = 27; the model matrices
                                      are rebuilt with oracle/c3_model_oracle.py, checked on the pickled
                                      coupler_01 / coupler_12 eigenfrequencies)
  test/generator_data.pickle      <- test/test_generator.py:21-190 (signal chain, stage by stage)
  test/envelopes.pickle           <- test/test_envelopes.py (envelope shapes on ts = linspace(0, 10e-9, 100))
"""
import os
import pickle
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
REF = os.environ.get("C3_REFERENCE", "/root/reference")

from oracle import c3_model_oracle as mo  # noqa: E402


class _StubUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.startswith("tensorflow"):
            return lambda *a, **k: np.asarray(a[0])
        if module.startswith("numpy.core"):
            module = module.replace("numpy.core", "numpy._core")
        return super().find_class(module, name)


def load(name):
    with open(os.path.join(REF, "test", name), "rb") as f:
        return _StubUnpickler(f).load()


def two_qubit():
    d = load("two_qubit_data.pickle")
    # collapse operators: dims [2,2], T1 = 20 us, T2* = 40 us for both qubits, dressed basis
    dims = [2, 2]
    a1, a2 = mo.annihilators(dims)
    drift = (2 * np.pi * 5e9) * mo.resonator(a1) + (2 * np.pi * 5.6e9) * mo.resonator(a2) \
        + (2 * np.pi * 20e6) * mo.int_XX(a1, a2)
    _, T = mo.dressing_transform(drift)
    col_ops = np.stack([mo.dress(T, mo.qubit_collapse_op(a, t1=20e-6, t2star=40e-6)) for a in (a1, a2)])
    h0_rebuilt = mo.dress(T, drift)
    hks_rebuilt = np.stack([mo.dress(T, mo.x_drive(a)) for a in (a1, a2)])
    np.savez_compressed(
        os.path.join(HERE, "two_qubit.npz"),
        hdrift=d["hdrift"], hks=np.stack([d["hks"]["d1"], d["hks"]["d2"]]),
        signals=np.stack([d["signal"]["d1"]["values"], d["signal"]["d2"]["values"]]),
        ts=d["signal"]["d1"]["ts"], propagator=d["propagator"],
        lindblad_propagator=d["lindblad_propagator"], col_ops=col_ops,
        h0_rebuilt=h0_rebuilt, hks_rebuilt=hks_rebuilt,
    )


def transmon_expanded():
    d = load("transmon_expanded.pickle")
    np.savez_compressed(
        os.path.join(HERE, "transmon_expanded.npz"),
        dims=np.array([6, 4]), max_excitations=np.array(4),
        ts_q1=d["signal_q1"]["ts"], ts_q2=d["signal_q2"]["ts"],
        hamiltonians_q1=d["hamiltonians_q1"], hamiltonians_q2=d["hamiltonians_q2"],
        partial_propagators_q1=d["partial_propagators_q1"],
        partial_propagators_q2=d["partial_propagators_q2"],
        propagators_q1=d["propagators_q1"], propagators_q2=d["propagators_q2"],
    )


def tf_utils():
    d = load("test_tf_utils.pickle")
    out = {}
    for key in ("tf_kron", "tf_spre", "tf_spost", "Id_like"):
        for i, el in enumerate(d[key]):
            if key == "tf_kron":
                out[f"{key}_{i}_inA"], out[f"{key}_{i}_inB"] = el["in"]
            else:
                out[f"{key}_{i}_in"] = el["in"]
            out[f"{key}_{i}_desired"] = el["desired"]
    el = d["tf_super"][0]   # the two tf_super cases are 5x100x100 each; one is kept
    out["tf_super_0_in"], out["tf_super_0_desired"] = el["in"], el["desired"]
    np.savez_compressed(os.path.join(HERE, "tf_utils.npz"), **out)


def tunable_coupler():
    """d = 27 (3-level tunable coupler + two 3-level qubits).  The pickle stores the flux-line control field
    ``tc_signal`` and every 50th slice propagator; the dressed drift and the dressed z-drive Hamiltonian are
    rebuilt (the two qubit drive lines carry ``no_drive``: their fields are identically zero)."""
    d = load("tunable_coupler_data.pickle")
    m = mo.tunable_coupler_model()
    c01 = abs(abs(m["eigenframe"][0]) - abs(m["eigenframe"][9]))     # labels (0,0,0) -> (1,0,0)
    c12 = abs(abs(m["eigenframe"][9]) - abs(m["eigenframe"][18]))
    assert abs(c01 - d["coupler_01"]) / d["coupler_01"] < 1e-12     # test_tunable_coupler.py:295-302
    assert abs(c12 - d["coupler_12"]) / d["coupler_12"] < 1e-12     # :305-312
    keep = np.arange(0, 200, 2)                                      # slices 0, 100, 200, ... of 10 000
    np.savez_compressed(
        os.path.join(HERE, "tunable_coupler.npz"),
        h0=m["h0"], hk_tc=m["hk_tc"], tc_signal=d["tc_signal"], tc_ts=d["tc_ts"],
        dUs_index=keep * 50, dUs=np.asarray(d["dUs"])[keep],
        tc_awg_I=d["tc_awg_I"], tc_awg_Q=d["tc_awg_Q"], tc_awg_ts=d["tc_awg_ts"],
    )


def tunable_coupler_levels():
    """test/test_tunable_coupler.py:315-383: eigenframes over 101 coupler flux values, uncoupled (product_basis),
    dressed and reordered (ordered_basis), dressed ascending (dressed_basis), in GHz."""
    d = load("tunable_coupler_data.pickle")
    np.savez_compressed(os.path.join(HERE, "tunable_coupler_levels.npz"),
                        flux_ratio=np.linspace(-0.10, 0.7, 101, endpoint=True),
                        product_basis=d["product_basis"], ordered_basis=d["ordered_basis"], dressed_basis=d["dressed_basis"])


def generator_chain():
    d = load("generator_data.pickle")
    np.savez_compressed(
        os.path.join(HERE, "generator.npz"),
        lo_I=d["lo_sig"]["values"][0], lo_Q=d["lo_sig"]["values"][1], lo_ts=d["lo_sig"]["ts"],
        awg_I=d["awg_sig"]["inphase"], awg_Q=d["awg_sig"]["quadrature"],
        dac_I=d["dig_to_an_sig"]["inphase"], dac_Q=d["dig_to_an_sig"]["quadrature"],
        resp_I=d["resp_sig"]["inphase"], resp_Q=d["resp_sig"]["quadrature"],
        mixer=d["mixer_sig"], v2hz=d["v2hz_sig"],
        full_values=d["full_signal"][0]["d1"]["values"], full_ts=d["full_signal"][0]["d1"]["ts"],
    )


def envelopes():
    """The shape functions the on-device signal chain implements (real part; the reference complexifies with a zero imaginary
    part), with the parameters of test/test_envelopes.py: t_final 10 ns, sigma 5 ns, risefall 2 ns, t_up 1 ns, t_down 10 ns."""
    d = load("envelopes.pickle")
    keep = ["trapezoid", "flattop_risefall", "flattop_risefall_1ns", "flattop", "gaussian_sigma", "gaussian", "gaussian_nonorm",
            "gaussian_der_nonorm", "gaussian_der", "drag_sigma", "drag_der", "drag", "cosine", "no_drive", "rect"]
    # the array-parametrised / grid-defined shapes (test_pwc_shape, test_delta_pulse, test_fourier, test_flattop,
    # test_flattop_cut, test_cosine of test/test_envelopes.py; the parameters are restated in tests/test_signal_oracle.py)
    keep += ["pwc_shape", "pwc_symmetric", "pwc_shape_plateau1", "pwc_shape_plateau2", "delta_pulse", "fourier_sin", "fourier_cos",
             "slepian_fourier", "slepian_fourier_risefall", "slepian_fourier_sin", "flattop_variant", "flattop_cut",
             "flattop_cut_center", "cosine_flattop"]
    out = {k: np.real(np.asarray(d[k])).reshape(-1).astype(np.float64) for k in keep}
    # pwc_shape_plateau with "width": the reference's tf.where broadcasts its [100, 1] shape against the [100] time vector into
    # [100, 100] (entry [i, j] = 1 where x[j] == t_mid else shape[i]); the element-wise result is the diagonal
    out["pwc_shape_plateau2"] = np.real(np.asarray(d["pwc_shape_plateau2"])).reshape(100, 100).diagonal().astype(np.float64).copy()
    assert all(v.shape == (100,) for v in out.values()), {k: v.shape for k, v in out.items()}
    assert all(np.abs(np.imag(np.asarray(d[k]))).max() == 0 for k in keep)
    np.savez_compressed(os.path.join(HERE, "envelopes.npz"), ts=np.linspace(0, 10e-9, 100), **out)


if __name__ == "__main__":
    two_qubit()
    tunable_coupler()
    tunable_coupler_levels()
    generator_chain()
    transmon_expanded()
    tf_utils()
    envelopes()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
