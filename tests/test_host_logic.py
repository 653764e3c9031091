"""CPU tests of the host-side mirror of the reference interface: registries, Experiment
plumbing with a stub propagation method, flop accounting, synthetic workloads, sharding."""
import numpy as np
import pytest
import torch

from oracle import c3_oracle as orc


def test_registries_mirror_reference():
    from c3_b200 import propagation as prop
    assert "pwc" in prop.unitary_provider and prop.unitary_provider["pwc"] is prop.pwc

    @prop.unitary_deco
    def my_method(model, gen, instr, folding_stack, batch_size=None):
        return {}
    assert prop.unitary_provider["my_method"] is my_method
    del prop.unitary_provider["my_method"]


def test_folding_stack_matches_reference_rule():
    from c3_b200 import tf_utils as tu
    for n in [1, 2, 3, 7, 50, 700, 1000]:
        mine = [f.__name__ for f in tu.compute_folding_stack(n)]
        want = [f.__name__ for f in orc.compute_folding_stack(n)]
        assert mine == want


class _Q:
    def __init__(self, v):
        self.v = v

    def get_value(self):
        return self.v


class _Ctrl:
    def __init__(self, **p):
        self.params = {k: _Q(v) for k, v in p.items()}


class _Instr:
    def __init__(self, t_end):
        self.t_start, self.t_end = 0.0, t_end
        self.comps = {"d1": {"gauss": _Ctrl(amp=0.5, freq_offset=1e6), "carrier": _Ctrl(freq=5e9, framechange=0.1)}}


class _Model:
    controllability = True
    lindbladian = False
    max_excitations = 0
    use_FR = False
    dephasing_strength = 0.0


class _PMap:
    def __init__(self):
        self.model = _Model()
        self.generator = object()
        self.instructions = {"rx90p[0]": _Instr(7e-9), "ry90p[0]": _Instr(7e-9)}


def test_experiment_calls_plugin_positionally_and_stores_results():
    """c3/experiment.py:472-478: (model, generator, instr, folding_stack[steps], batch_size)."""
    from c3_b200.experiment import Experiment
    calls = []

    def fake_prop(model, gen, instr, folding_stack, batch_size):
        calls.append((instr, len(folding_stack), batch_size))
        return {"U": torch.eye(2, dtype=torch.complex128), "dUs": torch.zeros(3, 2, 2), "ts": [0.0, 1.0]}

    exp = Experiment(_PMap(), prop_method=fake_prop, sim_res=100e9)
    exp.propagate_batch_size = 360
    props = exp.compute_propagators()
    assert set(props) == {"rx90p[0]", "ry90p[0]"}
    assert calls[0][2] == 360
    assert set(exp.partial_propagators) == set(props)
    exp.set_opt_gates("rx90p[0]")
    exp.overwrite_propagators = False
    exp.compute_propagators()
    assert set(exp.propagators) == {"rx90p[0]", "ry90p[0]"}
    exp.set_opt_gates(["nope"])
    with pytest.raises(Exception, match="C3:Error: Gate 'nope' is not defined"):
        exp.compute_propagators()


def test_experiment_default_method_builds_folding_stack():
    from c3_b200.experiment import Experiment
    from c3_b200 import propagation as prop
    exp = Experiment(_PMap(), sim_res=100e9)
    assert exp.propagation is prop.pwc
    assert 700 in exp.folding_stack and len(exp.folding_stack[700]) == 10
    exp.set_prop_method("pwc")
    assert exp.propagation is prop.pwc


def test_dephasing_requires_lindblad():
    from c3_b200.experiment import Experiment
    pm = _PMap()
    pm.model.dephasing_strength = 0.1
    exp = Experiment(pm, prop_method=lambda *a: {"U": torch.eye(2), "dUs": None, "ts": []})
    with pytest.raises(ValueError, match="Dephasing can only be added when lindblad is on"):
        exp.compute_propagators()


def test_flop_accounting_matches_survey_table():
    from c3_b200 import flops
    assert flops.higham_order(1.32) == (9, 0)
    assert flops.higham_order(0.5) == (7, 0)
    assert flops.higham_order(2.4) == (13, 0)
    assert flops.higham_order(6.0) == (13, 1)
    assert abs(flops.flops_closed(9, 2, 9, 0) - 43.4e3) < 0.1e3      # 43.4 kflop
    assert abs(flops.flops_closed(27, 3, 13, 0) - 1.32e6) < 0.01e6   # 1.32 Mflop
    assert abs(flops.flops_lindblad(9, 13, 0) - 35.4e6) < 0.1e6      # 35.4 Mflop
    # same rule as the oracle's bookkeeping
    for x in [1e-3, 0.1, 0.9, 1.4, 3.0, 12.0]:
        assert flops.higham_order(x) == orc.pade_order_and_squarings(x)


def test_synthetic_models_are_hermitian_and_in_the_expected_norm_range():
    from c3_b200 import synth
    m = synth.two_transmon()
    assert m.d == 9 and m.K == 2 and m.col_ops.shape == (2, 9, 9)
    assert np.abs(m.h0 - m.h0.conj().T).max() < 1e-12 * np.abs(m.h0).max()
    assert np.abs(m.hks - np.conj(np.swapaxes(m.hks, -1, -2))).max() < 1e-12
    sig = synth.controls(m, 3, 1000)
    assert sig.shape == (3, 2, 1000) and sig.dtype == np.float64
    H = m.h0[None, None] + np.einsum("bkn,kij->bnij", sig, m.hks)
    n1 = np.abs(H * 1e-11).sum(axis=-2).max(axis=-1)
    assert 1.25 < n1.min() and n1.max() < 1.45          # Pade-9, no squarings
    assert synth.one_qubit().d == 3 and synth.tunable_coupler().d == 27
    idx, lens = synth.rb_sequences(64, 20, 5)
    assert idx.shape[0] == 64 and lens.min() >= 20 and idx.max() < 5
    # reproducible, rank-disjoint draws
    a = synth.controls_fast(m, 4, 50, b_offset=0)
    b = synth.controls_fast(m, 4, 50, b_offset=1000003)
    assert np.array_equal(a, synth.controls_fast(m, 4, 50, b_offset=0)) and not np.array_equal(a, b)


def test_shard_bounds_cover_the_batch():
    from c3_b200.distributed import shard_bounds
    for B in [1, 7, 8, 4096, 4099]:
        for world in [1, 2, 3, 8]:
            spans = [shard_bounds(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_blowup_and_cut_are_inverse_scatter_gather():
    from c3_b200 import propagation as prop
    cutter = orc.make_ex_cutter([3, 2], 2)
    rng = np.random.default_rng(0)
    small = torch.as_tensor(rng.normal(size=(4, cutter.shape[0], cutter.shape[0])) + 0j)
    big = prop.blowup_excitations(cutter, small)
    for i in range(4):
        np.testing.assert_allclose(big[i].numpy(), orc.blowup_excitations(cutter, small[i].numpy()))
    back = prop.cut_excitations(cutter, big)
    assert torch.equal(back, small)
