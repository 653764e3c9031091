"""Mirror of the propagator-consuming goal functions of ``c3.libraries.fidelities`` on the B200 engine
(SURVEY.md section 8f, row f-3).  Same names, arguments and registries as the reference
(c3/libraries/fidelities.py @ 48b7917e):

  unitary_infid / unitary_infid_set                    :152-218
  lindbladian_unitary_infid / _set                     :221-285
  average_infid / average_infid_set / average_infid_seq  :288-374
  lindbladian_average_infid / _set                     :377-432
  orbit_infid                                          :753-790

plus the batch axis: ``actual`` may be ``[B,d,d]`` (then the result is ``[B]``), and
:func:`unitary_infid_autograd` is differentiable w.r.t. the propagators so that
``pwc_batch_autograd -> unitary_infid_autograd -> backward`` runs a whole GRAPE step on the device.
Every reduction is ONE kernel launch (c3b_gate_infid / c3b_seq_populations); there is no CPU path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import engine
from .propagation import _np

state_providers: Dict = dict()
unitary_providers: Dict = dict()
set_providers: Dict = dict()
super_providers: Dict = dict()
fidelities: Dict = dict()


def fid_reg_deco(func):
    fidelities[str(func.__name__)] = func
    return func


def state_deco(func):
    state_providers[str(func.__name__)] = func
    return func


def unitary_deco(func):
    unitary_providers[str(func.__name__)] = func
    return func


def set_deco(func):
    set_providers[str(func.__name__)] = func
    return func


def open_system_deco(func):
    super_providers[str(func.__name__)] = func
    return func


def comp_indices(dims: Sequence[int], index: Optional[Sequence[int]] = None) -> np.ndarray:
    """Rows of the full space kept by ``qt_utils.projector(dims, index)`` (c3/utils/qt_utils.py:178-193):
    the lowest two levels of every subsystem in ``index``, the lowest level of the others; ordered like
    the projector's columns (Kronecker order of the kept levels)."""
    dims = [int(d) for d in dims]
    if not index:
        index = list(range(len(dims)))
    sel = np.zeros(1, dtype=np.int64)
    for i, dim in enumerate(dims):
        keep = np.arange(min(2, dim)) if i in index else np.arange(1)
        sel = (sel[:, None] * dim + keep[None, :]).reshape(-1)
    return sel.astype(np.int32)


def _super_sel(sel: np.ndarray, d: int) -> np.ndarray:
    return (sel[:, None].astype(np.int64) * d + sel[None, :]).reshape(-1).astype(np.int32)


def _host(x) -> np.ndarray:
    x = _np(x)
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def _infid(ideal, actual, index, dims, mode: str):
    dims = [int(d) for d in dims]
    d = int(np.prod(dims))
    sel = comp_indices(dims, index)
    G = _host(ideal).astype(np.complex128)
    if mode.startswith("lindbladian"):
        G = np.kron(G, G.conj())                      # tf_super(ideal) = G (x) G^*  (tf_utils.py:283-289)
        sel = _super_sel(sel, d)
    return engine.gate_infid(_np(actual), G, sel, mode)


@fid_reg_deco
@unitary_deco
def unitary_infid(ideal, actual, index: List[int] = None, dims=None):
    """Unitary overlap between ideal and actually performed gate (fidelities.py:152-183)."""
    if index is None:
        index = list(range(len(dims)))
    return _infid(ideal, actual, index, dims, "unitary")


@fid_reg_deco
@open_system_deco
def average_infid(ideal, actual, index: List[int] = [0], dims=[2]):
    """Average fidelity in the Pauli basis (fidelities.py:288-311)."""
    return _infid(ideal, actual, index, dims, "average")


@fid_reg_deco
@open_system_deco
def lindbladian_unitary_infid(ideal, actual, index=[0], dims=[2]):
    """Variant of the unitary fidelity for the Lindbladian propagator (fidelities.py:221-249)."""
    return _infid(ideal, actual, index, dims, "lindbladian_unitary")


@fid_reg_deco
@open_system_deco
def lindbladian_average_infid(ideal, actual, index=[0], dims=[2]):
    """Average fidelity of a Lindbladian propagator (fidelities.py:377-399).  As in the reference the
    double projection only type-checks for two-level subsystems (``dims`` all 2)."""
    if any(int(x) != 2 for x in dims):
        raise ValueError("C3:ERROR: lindbladian_average_infid needs two-level subsystems (dims all 2), as in the reference.")
    return _infid(ideal, actual, index, dims, "lindbladian_average")


def _set(fn, propagators: dict, instructions: dict, index, dims, ideal_args):
    infids = []
    for gate, propagator in propagators.items():
        perfect_gate = instructions[gate].get_ideal_gate(*ideal_args)
        infids.append(fn(perfect_gate, propagator, index, dims))
    return torch.stack([torch.as_tensor(x) for x in infids]).mean(dim=0)


@fid_reg_deco
@unitary_deco
@set_deco
def unitary_infid_set(propagators: dict, instructions: dict, index, dims, n_eval=-1):
    """Mean unitary infidelity over the gates in ``propagators`` (fidelities.py:186-218)."""
    return _set(unitary_infid, propagators, instructions, index, dims, (dims, index))


@fid_reg_deco
@open_system_deco
@set_deco
def lindbladian_unitary_infid_set(propagators: dict, instructions: dict, index, dims, n_eval=-1):
    """fidelities.py:252-285."""
    return _set(lindbladian_unitary_infid, propagators, instructions, index, dims, (dims,))


@fid_reg_deco
@open_system_deco
@set_deco
def average_infid_set(propagators: dict, instructions: dict, index: List[int], dims, n_eval=-1):
    """Mean average infidelity over all gates in ``propagators`` (fidelities.py:314-347)."""
    return _set(average_infid, propagators, instructions, index, dims, (dims, index))


@fid_reg_deco
@open_system_deco
@set_deco
def average_infid_seq(propagators: dict, instructions: dict, index, dims, n_eval=-1):
    """Average sequence fidelity over all gates in ``propagators`` (fidelities.py:350-374)."""
    fid = 1
    for gate, propagator in propagators.items():
        perfect_gate = instructions[gate].get_ideal_gate(dims)
        fid = fid * (1 - average_infid(perfect_gate, propagator, index, dims))
    return 1 - fid


@fid_reg_deco
@open_system_deco
@set_deco
def lindbladian_average_infid_set(propagators: dict, instructions: dict, index, dims, n_eval=-1):
    """fidelities.py:402-432."""
    return _set(lindbladian_average_infid, propagators, instructions, index, dims, (dims,))


def _encode_sequences(propagators: Dict, sequences: list):
    names = list(propagators.keys())
    lookup = {n: i for i, n in enumerate(names)}
    first = propagators[names[0]]
    dev = first.device if isinstance(first, torch.Tensor) and first.is_cuda else engine.default_device()
    gates = torch.stack([torch.as_tensor(_np(propagators[n])).to(torch.complex128).to(dev) for n in names])
    lens = np.array([len(s) for s in sequences], dtype=np.int32)
    Lmax = int(lens.max()) if len(sequences) else 0
    idx = np.zeros((len(sequences), max(Lmax, 1)), dtype=np.int32)
    for i, seq in enumerate(sequences):
        for j, g in enumerate(seq):
            idx[i, j] = lookup[g]
    return gates, idx, lens, dev


def sequence_populations(propagators: Dict, sequences: list, psi_init=None, lindbladian: bool = False) -> torch.Tensor:
    """Populations after every gate sequence, ``[S, d]`` -- what ``Experiment.evaluate_legacy`` computes
    with a Python loop over sequences and gates (c3/experiment.py:273-302, 603-624)."""
    gates, idx, lens, dev = _encode_sequences(propagators, sequences)
    D = gates.shape[-1]
    ld = int(round(np.sqrt(D))) if lindbladian else 0
    return engine.seq_populations(gates, idx, lens, psi0=None if psi_init is None else _np(psi_init), lindblad_d=ld,
                                  device=dev)


@fid_reg_deco
def orbit_infid(propagators, RB_number: int = 30, RB_length: int = 20, lindbladian=False, shots: int = None,
                seqs=None, noise=None):
    """ORBIT goal function (fidelities.py:753-790): mean over random Clifford sequences of
    ``1 - |<0|U_seq|0>|^2``; with ``shots`` a binomial sample of it, with ``noise`` additive Gaussian noise."""
    if not seqs:
        from .synth import single_length_RB
        seqs = single_length_RB(RB_number=RB_number, RB_length=RB_length)
    gates, idx, lens, dev = _encode_sequences(propagators, seqs)
    pops = engine.seq_populations(gates, idx, lens, device=dev)        # psi_init = basis(dim, 0), |psi|^2
    p1 = 1.0 - pops[:, 0]
    if shots:
        infids = torch.binomial(torch.full_like(p1, float(shots)), p1.clamp(0.0, 1.0)) / float(shots)
    else:
        infids = p1
    if noise:
        infids = infids + torch.randn_like(infids) * float(noise)
    return infids.mean()


class _GateInfidFn(torch.autograd.Function):
    """infid[b] = f(U[b]) with the analytic cotangent (c3b_gate_infid_grad)."""

    @staticmethod
    def forward(ctx, U, ideal, sel, mode):
        infid, ov = engine.gate_infid(U.detach(), ideal, sel, mode, return_overlap=True)
        ctx.save_for_backward(ov)
        ctx.ideal, ctx.sel, ctx.mode, ctx.D = ideal, sel, mode, U.shape[-1]
        return infid

    @staticmethod
    def backward(ctx, gbar):
        (ov,) = ctx.saved_tensors
        return engine.gate_infid_grad(ov, ctx.ideal, ctx.sel, gbar.contiguous(), ctx.D, ctx.mode), None, None, None


def unitary_infid_autograd(ideal, actual: torch.Tensor, index=None, dims=None, average: bool = False) -> torch.Tensor:
    """Differentiable ``unitary_infid`` (or ``average_infid``) of a batch of propagators ``[B,d,d]``."""
    dims = [int(d) for d in dims]
    if index is None:
        index = list(range(len(dims)))
    sel = torch.as_tensor(comp_indices(dims, index), device=actual.device)
    G = torch.as_tensor(_host(ideal).astype(np.complex128), device=actual.device)
    U = actual if actual.dim() == 3 else actual.unsqueeze(0)
    return _GateInfidFn.apply(U, G, sel, "average" if average else "unitary")
