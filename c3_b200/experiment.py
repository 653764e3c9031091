"""``Experiment``-shaped host for the propagator path (mirror of c3/experiment.py:59-107,
440-558): picks the propagation method, loops over gates, applies frame rotation and
dephasing, stores ``propagators`` / ``partial_propagators``.

Only what ``compute_propagators`` touches is mirrored; Model / Generator / Instruction /
ParameterMap are duck-typed exactly as the reference uses them:

  pmap.model, pmap.generator, pmap.instructions{name: instr}
  instr.t_start, instr.t_end, instr.comps{line: {name: ctrl.params{...}}}
  model.controllability, .lindbladian, .max_excitations, .ex_cutter, .use_FR, .dephasing_strength,
  model.get_Hamiltonians(), .get_Hamiltonian(signal), .get_Lindbladians(),
  model.get_Frame_Rotation(t_final, freqs, framechanges), .get_dephasing_channel(t_final, amps)
  generator.generate_signals(instr), generator.devices["awg"].get_average_amp()
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional

import numpy as np
import torch

from . import engine
from .propagation import state_provider, unitary_provider, _host
from .tf_utils import compute_folding_stack, tf_super


def _value(q):
    """Quantity-like -> python number (c3/c3objs.py:247-256 get_value)."""
    if hasattr(q, "get_value"):
        q = q.get_value()
    if hasattr(q, "numpy"):
        q = q.numpy()
    return complex(np.asarray(q).reshape(-1)[0]) if np.iscomplexobj(q) else float(np.asarray(q).reshape(-1)[0])


class Experiment:
    """It models all of the behaviour of the physical experiment, serving as a host for the
    individual parts making up the experiment (c3/experiment.py:29-57)."""

    def __init__(self, pmap=None, prop_method=None, sim_res=100e9):
        self.pmap = pmap
        self.opt_gates: Optional[List[str]] = None
        self.propagators: Dict[str, torch.Tensor] = {}
        self.partial_propagators: Dict = {}
        self.created_by = None
        self.logdir: str = ""
        self.propagate_batch_size = None
        self.use_control_fields = True
        self.overwrite_propagators = True  # Keep only currently computed propagators
        self.compute_propagators_timestamp = 0
        self.stop_partial_propagator_gradient = True
        self.sim_res = sim_res
        self.prop_method = prop_method
        self.folding_stack: Dict[int, list] = {}
        self.set_prop_method(prop_method)

    def set_prop_method(self, prop_method=None) -> None:
        """Configure the selected propagation method by either linking the function handle or
        looking it up in the library (c3/experiment.py:76-91)."""
        if prop_method is None:
            self.propagation = unitary_provider["pwc"]
            if self.pmap is not None:
                self._compute_folding_stack()
        elif isinstance(prop_method, str):
            try:
                self.propagation = unitary_provider[prop_method]
            except KeyError:
                self.propagation = state_provider[prop_method]
        elif callable(prop_method):
            self.propagation = prop_method

    def _compute_folding_stack(self):
        """c3/experiment.py:93-107 (kept for call compatibility; the kernel ignores it)."""
        self.folding_stack = {}
        for instr in self.pmap.instructions.values():
            n_steps = int((instr.t_end - instr.t_start) * self.sim_res)
            if n_steps not in self.folding_stack:
                self.folding_stack[n_steps] = compute_folding_stack(n_steps)

    def set_opt_gates(self, gates):
        """c3/experiment.py:536-547."""
        if type(gates) is str:
            gates = [gates]
        self.opt_gates = gates

    def compute_propagators(self):
        """Compute the unitary representation of operations. If no operations are specified in
        self.opt_gates the complete gateset is computed (c3/experiment.py:440-534)."""
        model = self.pmap.model
        generator = self.pmap.generator
        instructions = self.pmap.instructions
        propagators = {}
        partial_propagators = {}
        gate_ids = self.opt_gates
        if gate_ids is None:
            gate_ids = instructions.keys()

        self.set_prop_method(self.prop_method)

        for gate in gate_ids:
            try:
                instr = instructions[gate]
            except KeyError:
                raise Exception(
                    f"C3:Error: Gate '{gate}' is not defined."
                    f" Available gates are:\n {list(instructions.keys())}."
                )

            model.controllability = self.use_control_fields
            steps = int((instr.t_end - instr.t_start) * self.sim_res)
            result = self.propagation(
                model,
                generator,
                instr,
                self.folding_stack.get(steps, []),
                self.propagate_batch_size,
            )
            U = result["U"]
            dUs = result["dUs"]
            self.ts = result["ts"]
            if getattr(model, "use_FR", False):
                freqs = {}
                framechanges = {}
                for line, ctrls in instr.comps.items():
                    offset = 0.0
                    for ctrl in ctrls.values():
                        if "freq_offset" in ctrl.params.keys():
                            if _value(ctrl.params["amp"]) != 0.0:
                                offset = _value(ctrl.params["freq_offset"])
                    freqs[line] = complex(_value(ctrls["carrier"].params["freq"]) + offset)
                    framechanges[line] = complex(_value(ctrls["carrier"].params["framechange"]))
                t_final = complex(instr.t_end - instr.t_start)
                FR = torch.as_tensor(_host(model.get_Frame_Rotation(t_final, freqs, framechanges)),
                                     dtype=torch.complex128, device=U.device)
                if model.lindbladian:
                    SFR = tf_super(FR)
                    U = engine.ordered_product(torch.stack([U, SFR]))
                    self.FR = SFR
                else:
                    U = engine.ordered_product(torch.stack([U, FR]))
                    self.FR = FR
            if getattr(model, "dephasing_strength", 0.0) != 0.0:
                if not model.lindbladian:
                    raise ValueError("Dephasing can only be added when lindblad is on.")
                else:
                    amps = {}
                    for line, ctrls in instr.comps.items():
                        amp, _sum = generator.devices["awg"].get_average_amp()
                        amps[line] = complex(_value(amp))
                    t_final = complex(instr.t_end - instr.t_start)
                    dephasing_channel = torch.as_tensor(_host(model.get_dephasing_channel(t_final, amps)),
                                                        dtype=torch.complex128, device=U.device)
                    U = engine.ordered_product(torch.stack([U, dephasing_channel]))
            propagators[gate] = U
            partial_propagators[gate] = dUs

        if self.overwrite_propagators:
            self.propagators = propagators
            self.partial_propagators = partial_propagators
        else:
            self.propagators.update(propagators)
            self.partial_propagators.update(partial_propagators)
        self.compute_propagators_timestamp = time.time()
        return propagators
