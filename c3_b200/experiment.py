"""Host of the propagator path for the B200 engine.

Takes the place of ``c3.experiment.Experiment`` for ONE job: turning a gate set into propagators
(reference: ``Experiment.compute_propagators``, c3/experiment.py:440-534).  It accepts the attributes the reference's
optimisers set and read (``pmap``, ``opt_gates``, ``propagators``, ``partial_propagators``, ``propagate_batch_size``,
``use_control_fields``, ``overwrite_propagators``, ``set_prop_method``, ``set_opt_gates``), but it is organised around
what the engine is good at rather than around a per-gate Python loop:

* **a gate set is one launch.**  All gates of a call are gathered first (signals, time grid, collapse operators);
  gates that share a slice count and control lines become the batch axis of ONE fused kernel (B = number of gates).
  Frame rotations and dephasing channels of the whole set are applied by one batched product launch.
* **the model is prepared once.**  Generators (-i dt H_k, or the Lindblad superoperators), trace shifts and row sums live
  in a device-resident :class:`engine.PreparedModel`, rebuilt only when the Hamiltonians, the collapse operators or dt
  change (fingerprint of the host arrays).  A call with an unchanged model launches no setup kernels.
* **small fixed shapes replay a CUDA graph.**  With ``graph_calls`` the launch sequence of a (model, batch, slices) shape is
  captured once and replayed: one graph launch per ``compute_propagators`` for per-gate optimiser loops.
* **the batch axis is reachable.**  :meth:`compute_propagators_batch` takes per-sample pulse parameters ``[B]``, generates
  the control fields on the device and propagates all samples of a gate in one launch (optionally straight to
  infidelities) -- the CMA-ES population loop (c3/libraries/algorithms.py:553-559) or the noise grid of
  c3/optimizers/optimalcontrol_robust.py:54-62 as one call.

A user-supplied propagation method (``prop_method`` callable or registry name other than ``"pwc"``) is still honoured with
the reference's positional call ``(model, generator, instr, folding_stack, batch_size)`` (c3/experiment.py:472-478), one
gate at a time.  Model / Generator / Instruction / ParameterMap are duck-typed as the reference uses them.
"""
from __future__ import annotations

import hashlib
import time
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import engine
from . import propagation as prop
from .tf_utils import compute_folding_stack, tf_super


def _number(q):
    """Quantity-like -> python number (``get_value()`` of c3/c3objs.py:247-256, tf / numpy / torch scalars)."""
    if hasattr(q, "get_value"):
        q = q.get_value()
    if hasattr(q, "numpy"):
        q = q.numpy()
    a = np.asarray(q).reshape(-1)[0]
    return complex(a) if np.iscomplexobj(a) else float(a)


def _fingerprint(*arrays) -> str:
    h = hashlib.blake2b(digest_size=16)
    for a in arrays:
        if a is None:
            h.update(b"-")
            continue
        a = np.ascontiguousarray(prop._host(a))
        h.update(str(a.shape).encode())
        h.update(a.tobytes())
    return h.hexdigest()


class _Job:
    """One gate of a ``compute_propagators`` call on its way through the engine."""

    __slots__ = ("name", "instr", "inputs", "U", "dUs", "ts")

    def __init__(self, name, instr):
        self.name, self.instr = name, instr
        self.inputs = self.U = self.dUs = self.ts = None


class Experiment:
    """Gate set -> propagators on the GPU.  See the module docstring for how this differs from the reference's host."""

    def __init__(self, pmap=None, prop_method=None, sim_res=100e9):
        self.pmap = pmap
        self.sim_res = sim_res
        self.opt_gates: Optional[List[str]] = None
        self.propagators: Dict[str, torch.Tensor] = {}
        self.partial_propagators: Dict[str, torch.Tensor] = {}
        self.ts = None
        self.FR = None
        self.created_by = None
        self.logdir = ""
        self.propagate_batch_size = None          # the reference's host-memory guard; the fused kernel has no use for it
        self.use_control_fields = True
        self.overwrite_propagators = True
        self.stop_partial_propagator_gradient = True
        self.compute_propagators_timestamp = 0
        self.keep_partial_propagators = True      # store dUs [N,D,D] per gate as the reference does; off = U only, faster
        self.graph_calls = False                  # replay a captured CUDA graph for repeated shapes
        self.folding_stack: Dict[int, list] = {}
        self.launches_last_call = 0
        self._models: Dict[Tuple, Tuple[str, engine.PreparedModel]] = {}
        self._graphs: Dict[Tuple, engine.GraphedPwc] = {}
        self.prop_method = prop_method
        self.set_prop_method(prop_method)

    # ------------------------------------------------------------------------------------------------------------------
    # configuration (names as in c3/experiment.py:76-107, 536-547)
    # ------------------------------------------------------------------------------------------------------------------
    def set_prop_method(self, prop_method=None) -> None:
        """``None`` -> the engine's ``pwc``; a string -> registry lookup (unitary, then state providers); a callable ->
        used as is with the reference's positional call."""
        if callable(prop_method):
            self.propagation = prop_method
        else:
            name = "pwc" if prop_method is None else prop_method
            table = prop.unitary_provider if name in prop.unitary_provider else prop.state_provider
            self.propagation = table[name]
        if prop_method is None and self.pmap is not None:
            self._compute_folding_stack()

    def _compute_folding_stack(self) -> None:
        """Pairwise-product plans per slice count (c3/experiment.py:93-107).  The engine folds on chip and never reads
        them; they are handed to user plugins that follow the reference's convention."""
        counts = {self._slice_count(instr) for instr in self.pmap.instructions.values()}
        self.folding_stack = {n: compute_folding_stack(n) for n in sorted(counts)}

    def set_opt_gates(self, gates) -> None:
        self.opt_gates = [gates] if isinstance(gates, str) else gates

    def _slice_count(self, instr) -> int:
        return int((instr.t_end - instr.t_start) * self.sim_res)

    # ------------------------------------------------------------------------------------------------------------------
    # gate set -> propagators
    # ------------------------------------------------------------------------------------------------------------------
    def compute_propagators(self) -> Dict[str, torch.Tensor]:
        """Propagators of ``opt_gates`` (default: every instruction), with frame rotation and dephasing applied as in
        c3/experiment.py:482-522.  Returns ``{gate: U}``; also fills ``propagators`` / ``partial_propagators``."""
        jobs = self._jobs()
        model = self.pmap.model
        model.controllability = self.use_control_fields
        before = engine.launch_count() if torch.cuda.is_available() else 0
        self.set_prop_method(self.prop_method)
        if self.propagation is prop.pwc:
            self._run_gate_set(model, jobs)
        else:
            self._run_plugin(model, jobs)
        self._apply_frame_and_dephasing(model, jobs)
        done = {j.name: j.U for j in jobs}
        partial = {j.name: j.dUs for j in jobs}
        if self.overwrite_propagators:
            self.propagators, self.partial_propagators = done, partial
        else:
            self.propagators.update(done)
            self.partial_propagators.update(partial)
        if jobs:
            self.ts = jobs[-1].ts
        self.compute_propagators_timestamp = time.time()
        self.launches_last_call = (engine.launch_count() - before) if torch.cuda.is_available() else 0
        return done

    def _jobs(self) -> List[_Job]:
        instructions = self.pmap.instructions
        names = list(instructions.keys()) if self.opt_gates is None else list(self.opt_gates)
        missing = [n for n in names if n not in instructions]
        if missing:
            raise Exception(f"C3:Error: Gate '{missing[0]}' is not defined. Available gates are:\n {list(instructions.keys())}.")
        return [_Job(n, instructions[n]) for n in names]

    def _run_plugin(self, model, jobs: Sequence[_Job]) -> None:
        """A user propagation method: the reference's per-gate positional call."""
        for job in jobs:
            steps = self._slice_count(job.instr)
            if steps not in self.folding_stack:
                self.folding_stack[steps] = compute_folding_stack(steps)
            res = self.propagation(model, self.pmap.generator, job.instr, self.folding_stack[steps], self.propagate_batch_size)
            job.U, job.dUs, job.ts = res["U"], res["dUs"], res["ts"]

    def _run_gate_set(self, model, jobs: Sequence[_Job]) -> None:
        """Gather every gate, then one launch per group of gates with the same (mode, slice count, dt, control lines)."""
        groups: Dict[Tuple, List[_Job]] = {}
        for job in jobs:
            g = prop.gather_gate(model, self.pmap.generator, job.instr)
            job.inputs, job.ts = g, g.ts
            mode = "hlist" if g.hlist is not None else ("lindblad" if g.col_ops is not None else "closed")
            shape = tuple(g.hlist.shape[-2:]) if g.hlist is not None else ()
            groups.setdefault((mode, g.n_slices, round(g.dt / 1e-18), g.channels, shape), []).append(job)
        for (mode, _n, _dt, _ch, _shape), members in groups.items():
            if mode == "hlist":
                self._launch_hlist(members)
            else:
                self._launch_fields(model, members, lindblad=(mode == "lindblad"))
        for job in jobs:
            cutter = job.inputs.cutter
            if cutter is not None:
                job.U = prop.blowup_excitations(cutter, job.U)
                if job.dUs is not None:
                    job.dUs = prop.blowup_excitations(cutter, job.dUs)

    def _launch_hlist(self, members: Sequence[_Job]) -> None:
        first = members[0].inputs
        dev = engine.default_device()
        hs = torch.stack([torch.as_tensor(prop._host(m.inputs.hlist) if not isinstance(m.inputs.hlist, torch.Tensor)
                                          else m.inputs.hlist).to(torch.complex128).to(dev) for m in members])
        out = engine.pwc_closed_hlist(hs, first.dt, return_dUs=self.keep_partial_propagators)
        self._scatter(members, out)

    def _launch_fields(self, model, members: Sequence[_Job], lindblad: bool) -> None:
        first = members[0].inputs
        pm = self._prepared(model, first, lindblad)
        dev = pm.device
        rows = [m.inputs.signals if isinstance(m.inputs.signals, torch.Tensor) else torch.as_tensor(m.inputs.signals)
                for m in members]
        sig = torch.stack([r.to(dev, non_blocking=True) for r in rows])            # [G,K,N]
        if self.graph_calls:
            key = (id(pm), sig.shape[0], sig.shape[2], self.keep_partial_propagators)
            graph = self._graphs.get(key)
            if graph is None:
                if len(self._graphs) > 32:
                    self._graphs.clear()
                graph = self._graphs[key] = engine.GraphedPwc(pm, sig.shape[0], sig.shape[2], self.keep_partial_propagators)
            out = graph.run(sig)
            out = tuple(t.clone() for t in out) if isinstance(out, tuple) else out.clone()
        else:
            out = engine.pwc_prepared(pm, sig, return_dUs=self.keep_partial_propagators)
        self._scatter(members, out)

    def _scatter(self, members: Sequence[_Job], out) -> None:
        U, dUs = out if isinstance(out, tuple) else (out, None)
        for i, m in enumerate(members):
            m.U = U[i]
            m.dUs = dUs[i] if dUs is not None else None

    def _prepared(self, model, g: "prop.GateInputs", lindblad: bool) -> engine.PreparedModel:
        """The device-resident generators of ``model`` for slice length ``g.dt``; rebuilt only when an input changed."""
        key = (id(model), lindblad, g.channels)
        stamp = _fingerprint(g.h0, g.hks, *(g.col_ops or []), np.float64(g.dt))
        hit = self._models.get(key)
        if hit is not None and hit[0] == stamp:
            return hit[1]
        pm = engine.prepare_model(g.h0, g.hks, g.dt, col_ops=g.col_ops, lindblad=lindblad)
        self._models[key] = (stamp, pm)
        self._graphs = {k: v for k, v in self._graphs.items() if v.model is not (hit[1] if hit else None)}
        return pm

    # ------------------------------------------------------------------------------------------------------------------
    # frame rotation and dephasing: the whole gate set in one product launch
    # ------------------------------------------------------------------------------------------------------------------
    @staticmethod
    def _frame_arguments(instr) -> Tuple[dict, dict]:
        """Per drive line: carrier frequency (+ the frequency offset of a driven envelope) and frame change
        (c3/experiment.py:485-503)."""
        freqs, framechanges = {}, {}
        for line, comps in instr.comps.items():
            offset = 0.0
            for comp in comps.values():
                if "freq_offset" in comp.params and _number(comp.params["amp"]) != 0.0:
                    offset = _number(comp.params["freq_offset"])
            carrier = comps["carrier"].params
            freqs[line] = complex(_number(carrier["freq"]) + offset)
            framechanges[line] = complex(_number(carrier["framechange"]))
        return freqs, framechanges

    @staticmethod
    def _line_occupations(model, lines) -> Optional[np.ndarray]:
        """occ[l, s] = occupation number of the qubit driven by line l in product state s: the diagonal of the bare number
        operator Model.get_Frame_Rotation / get_dephasing_channel exponentiate (c3/model.py:553-570, 620-628).  ``None`` when
        the model does not expose the pieces (then the host matrices from the model's own methods are used)."""
        try:
            occ = []
            for line in lines:
                if hasattr(model, "line_to_index"):
                    idx = model.line_to_index[line]
                elif line in getattr(model, "couplings", {}):
                    idx = model.names.index(model.couplings[line].connected[0])
                elif line in getattr(model, "subsystems", {}):
                    idx = model.names.index(line)
                else:
                    return None
                a = prop._host(model.ann_opers[idx])
                num = a.T.conj() @ a
                diag = np.real(np.diag(num))
                if np.abs(num - np.diag(np.diag(num))).max() > 1e-12 or np.abs(diag - np.rint(diag)).max() > 1e-12:
                    return None
                occ.append(np.rint(diag).astype(np.int32))
            return np.stack(occ) if occ else None
        except (AttributeError, KeyError, ValueError, IndexError):
            return None

    def _apply_frame_and_dephasing(self, model, jobs: Sequence[_Job]) -> None:
        """U <- dephasing . FR . U for every gate (c3/experiment.py:482-522).  Both factors are diagonal in the product basis,
        so when the model exposes its number operators the whole gate set is ONE row-scaling launch on the device
        (engine.frame_dephase); otherwise the model's own matrices are multiplied on by one batched product launch."""
        use_fr = bool(getattr(model, "use_FR", False))
        deph = getattr(model, "dephasing_strength", 0.0) != 0.0
        if deph and not model.lindbladian:
            raise ValueError("Dephasing can only be added when lindblad is on.")
        if not jobs or not (use_fr or deph):
            return
        lines = list(jobs[0].instr.comps.keys())
        same_lines = all(list(j.instr.comps.keys()) == lines for j in jobs)
        same_shape = all(j.U.shape == jobs[0].U.shape for j in jobs)
        occ = self._line_occupations(model, lines) if (same_lines and same_shape) else None
        if occ is None:
            return self._apply_host_factors(model, jobs, use_fr, deph)
        phases = np.zeros((len(jobs), len(lines)))
        probs = np.zeros((len(jobs), len(lines))) if deph else None
        for g, job in enumerate(jobs):
            t_final = float(job.instr.t_end - job.instr.t_start)
            if use_fr:
                freqs, framechanges = self._frame_arguments(job.instr)
                phases[g] = [np.real(freqs[l] * t_final + framechanges[l]) for l in lines]
            if deph:
                amp, _ = self.pmap.generator.devices["awg"].get_average_amp()
                probs[g] = t_final * float(np.real(_number(amp))) * float(model.dephasing_strength)
        if deph and (probs.min() < 0 or probs.max() > 1):
            raise ValueError(f"Dephasing channel strength {probs.max()} is outside [0,1] range")
        U = torch.stack([j.U for j in jobs]).contiguous()
        engine.frame_dephase(U, occ, phases if use_fr else None, probs, lindblad=bool(model.lindbladian))
        for g, job in enumerate(jobs):
            job.U = U[g]
        if use_fr:       # the attribute the reference leaves behind: the last gate's (super-)frame rotation, as a matrix
            f = np.exp(1j * (occ.T @ phases[-1]))
            fr = np.diag(f)
            self.FR = torch.as_tensor(np.kron(fr, fr.conj()) if model.lindbladian else fr, device=U.device)

    def _apply_host_factors(self, model, jobs: Sequence[_Job], use_fr: bool, deph: bool) -> None:
        dev = jobs[0].U.device
        chains = []                                  # per gate: [U, FR?, dephasing?] -- later factors act from the left
        for job in jobs:
            t_final = complex(job.instr.t_end - job.instr.t_start)
            factors = [job.U]
            if use_fr:
                freqs, framechanges = self._frame_arguments(job.instr)
                FR = torch.as_tensor(prop._host(model.get_Frame_Rotation(t_final, freqs, framechanges)),
                                     dtype=torch.complex128, device=dev)
                if model.lindbladian:
                    FR = tf_super(FR)
                self.FR = FR
                factors.append(FR)
            if deph:
                amps = {}
                for line in job.instr.comps:
                    amp, _ = self.pmap.generator.devices["awg"].get_average_amp()
                    amps[line] = complex(_number(amp))
                factors.append(torch.as_tensor(prop._host(model.get_dephasing_channel(t_final, amps)),
                                               dtype=torch.complex128, device=dev))
            chains.append(torch.stack(factors))
        by_shape: Dict[Tuple, List[int]] = {}
        for i, c in enumerate(chains):
            by_shape.setdefault(tuple(c.shape), []).append(i)
        for idxs in by_shape.values():
            out = engine.ordered_product(torch.stack([chains[i] for i in idxs]))      # [G,M,D,D] -> [G,D,D]
            for k, i in enumerate(idxs):
                jobs[i].U = out[k]

    # ------------------------------------------------------------------------------------------------------------------
    # the batch axis: parameter samples -> propagators (-> infidelities)
    # ------------------------------------------------------------------------------------------------------------------
    def compute_propagators_batch(self, samples: Dict, gates: Optional[Iterable[str]] = None, goal: Optional[Callable] = None):
        """Propagators of every gate for B parameter samples at once.

        ``samples[(channel, component, parameter)] = [B] values`` override the instructions' pulse parameters per sample
        (what ``get_value()`` would return); the generator must offer ``generate_signals_batch`` (c3_b200.generator).
        Per gate: control fields ``[B,K,N]`` are generated on the device, all B samples propagate in one launch with the
        prepared model.  Returns ``{gate: U [B,D,D]}`` -- or, with ``goal(gate, U) -> [B]``, ``{gate: goal values}`` (e.g.
        ``lambda gate, U: engine.gate_infid(U, ideal[gate], sel)``), so that only 8 bytes per sample leave the device.
        Frame rotation / dephasing are per-gate constants and are applied to all samples by one product launch."""
        gen = self.pmap.generator
        if not hasattr(gen, "generate_signals_batch"):
            raise Exception("C3:ERROR: compute_propagators_batch needs a generator with generate_signals_batch (c3_b200.generator.Generator).")
        model = self.pmap.model
        model.controllability = True
        names = list(self.pmap.instructions.keys()) if gates is None else list(gates)
        out = {}
        for name in names:
            if name not in self.pmap.instructions:
                raise Exception(f"C3:Error: Gate '{name}' is not defined. Available gates are:\n {list(self.pmap.instructions.keys())}.")
            instr = self.pmap.instructions[name]
            sig, ts = gen.generate_signals_batch(instr, samples)                     # [B,K,N] on the device
            g = prop.GateInputs()
            h0, hctrls = model.get_Hamiltonians()
            g.channels = tuple(instr.comps.keys())
            g.h0, g.hks = prop._np(h0), np.stack([prop._host(hctrls[c]) for c in g.channels])
            ts_np = prop._host(ts)
            g.dt = float(ts_np[1] - ts_np[0])
            if model.max_excitations:
                g.cutter = prop._host(model.ex_cutter)
            if model.lindbladian:
                cols = [prop._host(c) for c in model.get_Lindbladians()]
                g.col_ops = [g.cutter @ c @ g.cutter.T for c in cols] if g.cutter is not None else cols
            pm = self._prepared(model, g, lindblad=bool(model.lindbladian))
            U = engine.pwc_prepared(pm, sig)
            if g.cutter is not None:
                U = prop.blowup_excitations(g.cutter, U)
            job = _Job(name, instr)
            job.U = U
            if getattr(model, "use_FR", False) or getattr(model, "dephasing_strength", 0.0) != 0.0:
                U = self._apply_constant_factors(model, job, samples)
            out[name] = goal(name, U) if goal is not None else U
        return out

    def _apply_constant_factors(self, model, job: _Job, samples: Optional[Dict] = None) -> torch.Tensor:
        """FR / dephasing of one gate onto all samples U [B,D,D].  With the model's number operators at hand the frame phases
        are PER SAMPLE (sampled carrier frequency / frame change) and the whole batch is one row-scaling launch; otherwise the
        gate's constant factor from the model's own methods is multiplied onto every sample by one product launch."""
        U = job.U
        B = U.shape[0]
        lines = list(job.instr.comps.keys())
        occ = self._line_occupations(model, lines)
        use_fr = bool(getattr(model, "use_FR", False))
        deph = getattr(model, "dephasing_strength", 0.0) != 0.0
        if deph and not model.lindbladian:
            raise ValueError("Dephasing can only be added when lindblad is on.")
        if occ is not None:
            t_final = float(job.instr.t_end - job.instr.t_start)
            phases = None
            if use_fr:
                freqs, framechanges = self._frame_arguments(job.instr)
                phases = np.zeros((B, len(lines)))
                for l, line in enumerate(lines):
                    f = np.full(B, np.real(freqs[line]))
                    fc = np.full(B, np.real(framechanges[line]))
                    if samples and (line, "carrier", "freq") in samples:       # the envelope's frequency offset stays on top
                        base = _number(job.instr.comps[line]["carrier"].params["freq"])
                        f = f - base + np.asarray(samples[(line, "carrier", "freq")], dtype=np.float64)
                    if samples and (line, "carrier", "framechange") in samples:
                        fc = np.asarray(samples[(line, "carrier", "framechange")], dtype=np.float64)
                    phases[:, l] = f * t_final + fc
            probs = None
            if deph:
                amp, _ = self.pmap.generator.devices["awg"].get_average_amp()
                probs = np.full((B, len(lines)), t_final * float(np.real(_number(amp))) * float(model.dephasing_strength))
            return engine.frame_dephase(U.contiguous(), occ, phases, probs, lindblad=bool(model.lindbladian))
        probe = _Job(job.name, job.instr)
        probe.U = torch.eye(U.shape[-1], dtype=torch.complex128, device=U.device)
        self._apply_host_factors(model, [probe], use_fr, deph)                       # F = (dephasing) (FR)
        chain = torch.stack([U, probe.U.unsqueeze(0).expand(B, -1, -1)], dim=1)      # [B,2,D,D]: F U
        return engine.ordered_product(chain)
