"""GPU tests of the API the reference's callers use: ``propagation.pwc(model, gen, instr, folding_stack, batch_size)`` and
``Experiment.compute_propagators()`` on CUDA, driven with duck-typed Model / Generator / Instruction objects built from the
reference's golden fixtures.  They mirror test/test_two_qubits.py:46-62 (closed), :193-213 (Lindblad,
propagate_batch_size = 360), test/test_transmon_expanded.py:252-283 (H list + max_excitations) and the frame-rotation /
dephasing post-processing of c3/experiment.py:482-522; the expected values are the pickled reference results where they
exist and the CPU oracle's restatement of the same call otherwise."""
import numpy as np
import pytest
import torch

import c3_fakes as fk
from conftest import rel_fro
from oracle import c3_oracle as orc

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture(scope="module")
def api():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from c3_b200 import engine, propagation, experiment
    return engine, propagation, experiment


def _two_qubit_setup(g, lindbladian=False, **model_kw):
    """The chip of test/test_two_qubits.py as arrays: hdrift, hks{d1,d2}, signal{d1,d2}, collapse operators."""
    ts = g["ts"]
    table = {"gate": {"d1": {"values": g["signals"][0], "ts": ts}, "d2": {"values": g["signals"][1], "ts": ts}}}
    model = fk.ArrayModel(g["hdrift"], {"d1": g["hks"][0], "d2": g["hks"][1]}, col_ops=g["col_ops"], dims=[2, 2],
                          lindbladian=lindbladian, line_to_index={"d1": 0, "d2": 1}, **model_kw)
    t_end = float(len(ts) * (ts[1] - ts[0]))
    instr = fk.drive_instruction("gate", t_end, ["d1", "d2"], freq=[5.05e9, 5.65e9], framechange=[0.3, -1.1], freq_offset=-53e6)
    return model, fk.TableGenerator(table, avg_amp=2.0e5), instr


def test_pwc_closed_golden(api, golden_two_qubit):
    """test/test_two_qubits.py:46-62 through the plugin entry point."""
    _, prop, _ = api
    g = golden_two_qubit
    model, gen, instr = _two_qubit_setup(g)
    res = prop.pwc(model, gen, instr, orc.compute_folding_stack(700), None)
    assert set(res) == {"U", "dUs", "ts"}
    assert res["U"].is_cuda and tuple(res["U"].shape) == (4, 4) and tuple(res["dUs"].shape) == (700, 4, 4)
    assert rel_fro(res["U"].cpu().numpy(), g["propagator"]) < TOL
    want = orc.pwc(model, gen, instr, orc.compute_folding_stack(700))
    assert rel_fro(res["dUs"].cpu().numpy(), want["dUs"]) < TOL
    np.testing.assert_array_equal(np.asarray(res["ts"]), g["ts"])


def test_experiment_closed_and_lindblad_golden(api, golden_two_qubit):
    """Experiment.compute_propagators on CUDA: closed 4x4, then Lindblad 16x16 with propagate_batch_size = 360
    (test/test_two_qubits.py:193-213)."""
    _, _, experiment = api
    g = golden_two_qubit
    model, gen, instr = _two_qubit_setup(g)
    exp = experiment.Experiment(fk.PMap(model, gen, {"gate": instr}), sim_res=100e9)
    props = exp.compute_propagators()
    assert rel_fro(props["gate"].cpu().numpy(), g["propagator"]) < TOL
    assert tuple(exp.partial_propagators["gate"].shape) == (700, 4, 4)
    assert exp.propagators["gate"] is props["gate"]
    model.lindbladian = True
    exp.propagate_batch_size = 360
    props = exp.compute_propagators()
    assert rel_fro(props["gate"].cpu().numpy(), g["lindblad_propagator"]) < TOL
    assert tuple(exp.partial_propagators["gate"].shape) == (700, 16, 16)


@pytest.mark.parametrize("q", ["q1", "q2"])
def test_pwc_hlist_with_max_excitations_golden(api, golden_transmon, q):
    """test/test_transmon_expanded.py:252-283: controllability off (Hamiltonian list from the model), excitation cut
    24 -> 14, blow-up of U and of every partial propagator -- through pwc and through Experiment."""
    _, prop, experiment = api
    g = golden_transmon
    ts = g[f"ts_{q}"]
    dt = ts[2] - ts[1]
    full_ts = np.concatenate([[ts[0] - dt], ts])
    model = fk.ArrayModel(np.zeros((24, 24)), {}, dims=[int(x) for x in g["dims"]], max_excitations=int(g["max_excitations"]),
                          hlist=g[f"hamiltonians_{q}"])
    model.controllability = False
    gen = fk.TableGenerator({"flux": {"Q": {"values": np.zeros(21), "ts": full_ts}}})
    instr = fk.Instruction("flux", 0.0, 20 * dt, ["Q"])
    res = prop.pwc(model, gen, instr, [], None)
    assert tuple(res["dUs"].shape) == (20, 24, 24)
    assert rel_fro(res["dUs"].cpu().numpy(), g[f"partial_propagators_{q}"]) < TOL
    assert rel_fro(res["U"].cpu().numpy(), g[f"propagators_{q}"]) < TOL
    exp = experiment.Experiment(fk.PMap(model, gen, {"flux": instr}), sim_res=1.0 / dt)
    exp.use_control_fields = False
    props = exp.compute_propagators()
    assert rel_fro(props["flux"].cpu().numpy(), g[f"propagators_{q}"]) < TOL
    assert rel_fro(exp.partial_propagators["flux"].cpu().numpy(), g[f"partial_propagators_{q}"]) < TOL


def _transmon_pair(lindbladian, max_excitations, N=60, gates=("a", "b", "c")):
    from c3_b200 import synth
    m = synth.two_transmon()
    dt = 1e-11
    ts = (np.arange(N) + 0.5) * dt
    sig = synth.controls(m, len(gates), N)
    table = {n: {"d1": {"values": sig[i, 0], "ts": ts}, "d2": {"values": sig[i, 1], "ts": ts}} for i, n in enumerate(gates)}
    model = fk.ArrayModel(m.h0, {"d1": m.hks[0], "d2": m.hks[1]}, col_ops=m.col_ops, dims=[3, 3], lindbladian=lindbladian,
                          max_excitations=max_excitations, line_to_index={"d1": 0, "d2": 1})
    # t_end a hair above N dt: the reference's `int((t_end - t_start) * sim_res)` (c3/experiment.py:471) must not round N down
    instrs = {n: fk.drive_instruction(n, N * dt * (1 + 1e-9), ["d1", "d2"], freq=[5.05e9 + 1e7 * i, 5.65e9], framechange=[0.1 * i, 0.7])
              for i, n in enumerate(gates)}
    return model, fk.TableGenerator(table, avg_amp=3.0e6), instrs


def test_lindblad_with_excitation_cutter(api):
    """Lindblad + max_excitations: the collapse operators are cut with the model's projector (propagation.py:317-321), the
    slice superoperators live in the cut space (36x36 for 9 -> 6 states) -- and the final blow-up P^T S P of a superoperator
    with the Hilbert-space projector is not defined: the reference (and its restatement) fail on the matmul shapes at
    propagation.py:338-339.  The engine computes the cut-space propagators correctly and raises a C3 error at the blow-up."""
    engine, prop, _ = api
    model, gen, instrs = _transmon_pair(True, 2, N=24, gates=("a",))
    with pytest.raises(ValueError):
        orc.pwc(model, gen, instrs["a"], orc.compute_folding_stack(24))
    with pytest.raises(Exception, match="C3:ERROR: cannot blow up"):
        prop.pwc(model, gen, instrs["a"], [], None)
    g = prop.gather_gate(model, gen, instrs["a"])
    assert g.col_ops[0].shape == (6, 6)
    U, dUs = engine.pwc_lindblad(g.h0, g.hks, g.col_ops, g.signals, g.dt, return_dUs=True)
    h0c, hkc = model.get_Hamiltonians()
    want_dUs = orc.tf_batch_propagate(h0c, np.stack([hkc["d1"], hkc["d2"]]), g.signals, g.dt, 24, col_ops=np.asarray(g.col_ops),
                                      lindbladian=True)
    assert rel_fro(dUs[0].cpu().numpy(), want_dUs) < TOL
    assert rel_fro(U[0].cpu().numpy(), orc.tf_matmul_left(want_dUs)) < TOL


@pytest.mark.parametrize("lindbladian,dephasing", [(False, 0.0), (True, 0.0), (True, 0.02)])
def test_frame_rotation_and_dephasing(api, lindbladian, dephasing):
    """c3/experiment.py:482-522: U <- FR U (closed), SFR U (Lindblad), dephasing channel from the left, with non-trivial
    frame phases that differ per gate; three gates of equal length = one fused launch + one product launch."""
    _, _, experiment = api
    model, gen, instrs = _transmon_pair(lindbladian, 0, N=40)
    model.use_FR = True
    model.dephasing_strength = dephasing
    exp = experiment.Experiment(fk.PMap(model, gen, instrs), sim_res=100e9)
    props = exp.compute_propagators()
    # The engine applies the EXACT diagonal exponentials.  The reference builds FR with tf.linalg.expm, whose floor-scaled
    # Pade loses accuracy at the norms a frame rotation has (|freq t_final| n ~ 25 here, ~ 400 for a 7 ns gate): its own
    # truncation error, 1e-8 here, is the only difference -- so the expectation uses accurate exponentials, and the
    # TensorFlow-restated ones are checked to agree to that error.
    model.exact_frames = True
    want, want_partial = orc.compute_propagators(model, gen, instrs, 100e9)
    model.exact_frames = False
    want_tf, _ = orc.compute_propagators(model, gen, instrs, 100e9)
    for name in instrs:
        assert rel_fro(props[name].cpu().numpy(), want[name]) < TOL, name
        assert rel_fro(props[name].cpu().numpy(), want_tf[name]) < 1e-6, name
        assert rel_fro(exp.partial_propagators[name].cpu().numpy(), want_partial[name]) < TOL, name
    # FR really does something here (otherwise the left-multiplication order would go untested)
    bare, _ = orc.compute_propagators(fk.ArrayModel(model.h0, model.hks, col_ops=model.col_ops, dims=[3, 3], lindbladian=lindbladian),
                                      gen, instrs, 100e9)
    assert rel_fro(want["b"], bare["b"]) > 1e-2


def test_gate_set_is_one_launch_and_model_is_cached(api):
    """Five gates of 700 slices at d = 9: the first call builds the model (3 setup kernels), every call is one fused launch
    plus one fold; a repeated call with an unchanged model launches no setup kernels; results match per-gate oracle calls."""
    _, _, experiment = api
    gates = ("rx90p", "ry90p", "rx90m", "ry90m", "id")
    model, gen, instrs = _transmon_pair(False, 0, N=700, gates=gates)
    exp = experiment.Experiment(fk.PMap(model, gen, instrs), sim_res=100e9)
    exp.keep_partial_propagators = False
    props = exp.compute_propagators()
    first = exp.launches_last_call
    props2 = exp.compute_propagators()
    assert exp.launches_last_call <= 2 < first <= 5
    assert exp.partial_propagators["id"] is None
    for name in gates:
        want = orc.pwc(model, gen, instrs[name], orc.compute_folding_stack(700))["U"]
        assert rel_fro(props[name].cpu().numpy(), want) < TOL
        assert torch.equal(props[name], props2[name])
    # a model update invalidates the prepared generators
    model.h0 = model.h0 * 1.001
    props3 = exp.compute_propagators()
    assert exp.launches_last_call == first
    want = orc.pwc(model, gen, instrs["id"], orc.compute_folding_stack(700))["U"]
    assert rel_fro(props3["id"].cpu().numpy(), want) < TOL
    # subset of gates, keep the others (overwrite_propagators off)
    exp.overwrite_propagators = False
    exp.set_opt_gates(["rx90p"])
    exp.compute_propagators()
    assert set(exp.propagators) == set(gates)


def test_graph_replay_matches_direct_calls(api):
    """graph_calls: the launch sequence of a (model, gates, slices) shape is captured once and replayed."""
    engine, _, experiment = api
    from c3_b200 import synth
    model, gen, instrs = _transmon_pair(False, 0, N=50, gates=("g",))
    exp = experiment.Experiment(fk.PMap(model, gen, instrs), sim_res=100e9)
    exp.graph_calls = True
    m = synth.two_transmon()
    for rep in range(3):
        sig = synth.controls(m, 1, 50, seed=77 + rep)
        gen.table["g"]["d1"]["values"], gen.table["g"]["d2"]["values"] = sig[0, 0], sig[0, 1]
        got = exp.compute_propagators()["g"]
        want = orc.propagate_batch(m.h0, m.hks, sig, 1e-11)[0]
        assert rel_fro(got.cpu().numpy(), want) < TOL
    assert len(exp._graphs) == 1
    # the captured graph itself: one replay per call, outputs in place
    pm = engine.prepare_model(m.h0, m.hks, 1e-11)
    gp = engine.GraphedPwc(pm, 1, 50, return_dUs=True)
    sig = synth.controls(m, 1, 50, seed=5)
    U, dUs = gp.run(torch.as_tensor(sig).cuda())
    wantU, want_dUs = orc.propagate_batch(m.h0, m.hks, sig, 1e-11, return_dUs=True)
    assert rel_fro(U.cpu().numpy(), wantU) < TOL and rel_fro(dUs.cpu().numpy(), want_dUs) < TOL


@pytest.mark.parametrize("lindblad", [False, True])
def test_prepared_model_matches_unprepared(api, lindblad):
    engine, _, _ = api
    from c3_b200 import synth
    m = synth.two_transmon(levels=2 if lindblad else 3)
    sig = synth.controls(m, 6, 90)
    pm = engine.prepare_model(m.h0, m.hks, 1e-11, col_ops=m.col_ops, lindblad=lindblad)
    U, dUs = engine.pwc_prepared(pm, sig, return_dUs=True)
    if lindblad:
        U0, dUs0 = engine.pwc_lindblad(m.h0, m.hks, m.col_ops, sig, 1e-11, return_dUs=True)
    else:
        U0, dUs0 = engine.pwc_closed(m.h0, m.hks, sig, 1e-11, return_dUs=True)
    assert torch.equal(U, U0) and torch.equal(dUs, dUs0)
    # per-sample models
    h0b = np.stack([m.h0 * (1 + 0.01 * b) for b in range(6)])
    hkb = np.stack([m.hks] * 6)
    pmb = engine.prepare_model(h0b, hkb, 1e-11, col_ops=m.col_ops, lindblad=lindblad)
    Ub = engine.pwc_prepared(pmb, sig)
    for b in (0, 5):
        want = orc.propagate_batch(h0b[b], m.hks, sig[b:b + 1], 1e-11, col_ops=m.col_ops if lindblad else None, lindbladian=lindblad)[0]
        assert rel_fro(Ub[b].cpu().numpy(), want) < TOL


def test_batched_lindblad_model_shapes(api):
    """Per-sample Lindblad models: shared collapse operators are expanded to [B,C,d,d], per-sample ones are accepted as a
    list of [B,d,d] or as [B,C,d,d]; wrong shapes raise instead of reading past the buffer."""
    engine, _, _ = api
    from c3_b200 import synth
    m = synth.two_transmon(levels=2)
    B, N = 4, 30
    sig = synth.controls(m, B, N)
    h0b = np.stack([m.h0 * (1 + 0.02 * b) for b in range(B)])
    hkb = np.stack([m.hks * (1 - 0.01 * b) for b in range(B)])
    colb = np.stack([m.col_ops * (1 + 0.1 * b) for b in range(B)])          # [B,C,d,d]
    U_shared = engine.pwc_lindblad(h0b, hkb, m.col_ops, sig, 1e-11)
    U_batched = engine.pwc_lindblad(h0b, hkb, colb, sig, 1e-11)
    U_list = engine.pwc_lindblad(h0b, hkb, [colb[:, c] for c in range(colb.shape[1])], sig, 1e-11)
    assert torch.equal(U_batched, U_list)
    for b in range(B):
        want = orc.propagate_batch(h0b[b], hkb[b], sig[b:b + 1], 1e-11, col_ops=m.col_ops, lindbladian=True)[0]
        assert rel_fro(U_shared[b].cpu().numpy(), want) < TOL
        want = orc.propagate_batch(h0b[b], hkb[b], sig[b:b + 1], 1e-11, col_ops=colb[b], lindbladian=True)[0]
        assert rel_fro(U_batched[b].cpu().numpy(), want) < TOL
    with pytest.raises(ValueError):
        engine.pwc_lindblad(h0b, m.hks, m.col_ops, sig, 1e-11)              # hks must be [B,K,d,d] next to a batched h0
    with pytest.raises(ValueError):
        engine.pwc_lindblad(m.h0, m.hks, colb, sig, 1e-11)                  # batched collapse operators need a batched model
    with pytest.raises(ValueError):
        engine.pwc_lindblad(h0b, hkb, colb[:2], sig, 1e-11)


def test_per_slice_drift_with_control_terms(api):
    """tf_propagation_vectorized with a 3-dim h0 [N,d,d] AND hks: a per-slice drift added to the control terms
    (propagation.py:430-436), not a per-sample model."""
    _, prop, _ = api
    rng = np.random.default_rng(3)
    N, d, K = 11, 5, 2
    h0 = rng.normal(size=(N, d, d)) + 1j * rng.normal(size=(N, d, d))
    h0 = h0 + np.conj(np.swapaxes(h0, 1, 2))
    hks = rng.normal(size=(K, d, d)) + 0j
    hks = hks + np.swapaxes(hks, 1, 2)
    c = rng.uniform(-1, 1, size=(K, N))
    want = orc.tf_propagation_vectorized(h0 + np.einsum("kn,kij->nij", c, hks), None, None, 0.1)
    got = prop.tf_propagation_vectorized(h0, hks, c, 0.1)
    got2 = prop.tf_batch_propagate(h0, hks, c, 0.1, batch_size=4)
    assert rel_fro(got.cpu().numpy(), want) < TOL and rel_fro(got2.cpu().numpy(), want) < TOL


def test_compute_propagators_batch_from_parameter_samples(api):
    """[B] pulse-parameter samples -> control fields on the device -> all samples of a gate in one launch -> infidelities:
    the CMA-ES population loop (c3/libraries/algorithms.py:553-559) as one call."""
    engine, prop, experiment = api
    from c3_b200 import synth
    from c3_b200.generator import Generator
    devices, chains, instr = fk.reference_generator_setup()
    gen = Generator(devices, chains)
    m = synth.one_qubit()
    model = fk.ArrayModel(m.h0, {"d1": m.hks[0]}, dims=[3])
    exp = experiment.Experiment(fk.PMap(model, gen, {"rx90p": instr}), sim_res=100e9)
    B = 33
    amps = np.linspace(0.2, 0.6, B)
    samples = {("d1", "gauss", "amp"): amps}
    U = exp.compute_propagators_batch(samples)["rx90p"]
    assert tuple(U.shape) == (B, 3, 3)
    sig, ts = gen.generate_signals_batch(instr, samples)
    dt = float(ts[1] - ts[0])
    want = orc.propagate_batch(m.h0, m.hks, sig.cpu().numpy(), dt)
    assert rel_fro(U.cpu().numpy(), want) < TOL
    # one sample equals the single-gate path with that parameter value
    instr.comps["d1"]["gauss"].params["amp"] = fk.Quantity(amps[7], "V")
    single = exp.compute_propagators()["rx90p"]
    assert rel_fro(U[7].cpu().numpy(), single.cpu().numpy()) < 1e-12
    ideal = np.array([[1, -1j], [-1j, 1]]) / np.sqrt(2)
    infid = exp.compute_propagators_batch(samples, goal=lambda gate, Ub: engine.gate_infid(Ub, ideal, [0, 1]))["rx90p"]
    assert tuple(infid.shape) == (B,)
    np.testing.assert_allclose(infid.cpu().numpy(), engine.gate_infid(U, ideal, [0, 1]).cpu().numpy(), rtol=0, atol=1e-14)


def test_unknown_gate_message(api):
    _, _, experiment = api
    model, gen, instrs = _transmon_pair(False, 0, N=16, gates=("a",))
    exp = experiment.Experiment(fk.PMap(model, gen, instrs))
    exp.set_opt_gates(["nope"])
    with pytest.raises(Exception, match="C3:Error: Gate 'nope' is not defined"):
        exp.compute_propagators()


@pytest.mark.parametrize("lindbladian", [False, True])
def test_frame_rotation_host_fallback_matches_device_path(api, lindbladian):
    """A model that only offers get_Frame_Rotation / get_dephasing_channel (no number operators to read): the factors come
    from the model's own methods and are multiplied on by one product launch -- same result as the row-scaling kernel."""
    _, _, experiment = api

    class OpaqueModel(fk.ArrayModel):
        """hides the pieces the device path reads"""
        def __getattribute__(self, name):
            if name in ("line_to_index", "couplings", "subsystems", "names") and object.__getattribute__(self, "_hide"):
                raise AttributeError(name)
            return object.__getattribute__(self, name)

    model, gen, instrs = _transmon_pair(lindbladian, 0, N=30)
    opaque = OpaqueModel(model.h0, model.hks, col_ops=model.col_ops, dims=[3, 3], lindbladian=lindbladian,
                         line_to_index={"d1": 0, "d2": 1})
    for m in (model, opaque):
        m.use_FR = True
        m.dephasing_strength = 0.03 if lindbladian else 0.0
    object.__setattr__(opaque, "_hide", False)
    object.__setattr__(model, "_hide", False)
    exp_dev = experiment.Experiment(fk.PMap(model, gen, instrs), sim_res=100e9)
    got_dev = exp_dev.compute_propagators()
    object.__setattr__(opaque, "_hide", True)
    exp_host = experiment.Experiment(fk.PMap(opaque, gen, instrs), sim_res=100e9)
    orig = fk.ArrayModel.get_Frame_Rotation

    def fr_with_map(self, t_final, freqs, framechanges):
        object.__setattr__(self, "_hide", False)
        try:
            return orig(self, t_final, freqs, framechanges)
        finally:
            object.__setattr__(self, "_hide", True)
    OpaqueModel.get_Frame_Rotation = fr_with_map
    origd = fk.ArrayModel.get_dephasing_channel

    def deph_with_map(self, t_final, amps):
        object.__setattr__(self, "_hide", False)
        try:
            return origd(self, t_final, amps)
        finally:
            object.__setattr__(self, "_hide", True)
    OpaqueModel.get_dephasing_channel = deph_with_map
    got_host = exp_host.compute_propagators()
    # the host path reproduces whatever matrices the model hands out (here: the TensorFlow-restated exponentials) ...
    want_tf, _ = orc.compute_propagators(model, gen, instrs, 100e9)
    # ... the device path the exact ones
    model.exact_frames = True
    want, _ = orc.compute_propagators(model, gen, instrs, 100e9)
    for name in instrs:
        assert rel_fro(got_host[name].cpu().numpy(), want_tf[name]) < TOL
        assert rel_fro(got_dev[name].cpu().numpy(), want[name]) < TOL
        assert rel_fro(got_dev[name].cpu().numpy(), got_host[name].cpu().numpy()) < 1e-6
    assert rel_fro(exp_dev.FR.cpu().numpy(), exp_host.FR.cpu().numpy()) < 1e-6


def test_frame_dephase_kernel_directly(api):
    """engine.frame_dephase against explicit matrices: FR U (closed), dephasing . SFR . S (Lindblad), per-row phases."""
    engine, _, _ = api
    from oracle import c3_model_oracle as mo
    rng = np.random.default_rng(0)
    dims = [3, 2]
    d = 6
    ann = mo.annihilators(dims)
    occ = np.stack([np.rint(np.real(np.diag(a.T.conj() @ a))).astype(np.int32) for a in ann])
    B = 5
    phases = rng.uniform(-3, 3, size=(B, 2))
    probs = rng.uniform(0, 0.3, size=(B, 2))
    U = rng.normal(size=(B, d, d)) + 1j * rng.normal(size=(B, d, d))
    got = engine.frame_dephase(torch.as_tensor(U).cuda(), occ, phases).cpu().numpy()
    S = rng.normal(size=(B, d * d, d * d)) + 1j * rng.normal(size=(B, d * d, d * d))
    gotS = engine.frame_dephase(torch.as_tensor(S).cuda(), occ, phases, probs, lindblad=True).cpu().numpy()
    for b in range(B):
        # exact diagonal exponential (the oracle's expm_tf restates TensorFlow's Pade evaluation, which is only good to ~1e-11
        # at these norms -- the kernel's sincos of the exponent is the more accurate of the two)
        FR = np.diag(np.exp(1j * (occ.T @ phases[b])))
        assert rel_fro(got[b], FR @ U[b]) < 1e-13
        assert rel_fro(FR, orc.frame_rotation(ann, {"a": 0, "b": 1}, 1.0, {"a": phases[b, 0], "b": phases[b, 1]}, {"a": 0.0, "b": 0.0})) < 1e-9
        deph = orc.dephasing_channel(ann, {"a": 0, "b": 1}, 1.0, {"a": probs[b, 0], "b": probs[b, 1]}, 1.0)
        assert rel_fro(gotS[b], deph @ orc.tf_super(FR) @ S[b]) < 1e-12
    with pytest.raises(Exception, match="Dephasing can only be added when lindblad is on"):
        engine.frame_dephase(torch.as_tensor(U).cuda(), occ, phases, probs)


def test_batch_with_per_sample_frame_rotation(api):
    """compute_propagators_batch with use_FR and a sampled carrier frequency: every sample gets ITS frame rotation on the
    device (the phases differ per row), equal to the single-gate path evaluated at that sample's parameters."""
    engine, prop, experiment = api
    from c3_b200 import synth
    from c3_b200.generator import Generator
    devices, chains, instr = fk.reference_generator_setup()
    gen = Generator(devices, chains)
    m = synth.one_qubit()
    model = fk.ArrayModel(m.h0, {"d1": m.hks[0]}, dims=[3], line_to_index={"d1": 0})
    model.use_FR = True
    exp = experiment.Experiment(fk.PMap(model, gen, {"rx90p": instr}), sim_res=100e9)
    B = 9
    base = float(instr.comps["d1"]["carrier"].params["freq"].get_value())
    freqs = base + 2 * np.pi * np.linspace(-5e6, 5e6, B)
    U = exp.compute_propagators_batch({("d1", "carrier", "freq"): freqs})["rx90p"]
    for b in (0, 4, 8):
        instr.comps["d1"]["carrier"].params["freq"] = fk.Quantity(freqs[b] / (2 * np.pi), "Hz 2pi")
        single = exp.compute_propagators()["rx90p"]
        assert rel_fro(U[b].cpu().numpy(), single.cpu().numpy()) < 1e-11
    assert rel_fro(U[0].cpu().numpy(), U[8].cpu().numpy()) > 1e-3
