"""Drop-in mirror of ``c3.libraries.propagation`` for the PWC propagator path, backed by the
B200 engine (c3_b200.engine -> libc3b200.so).

Same names, argument meaning, return dictionaries and error strings as the reference
(c3/libraries/propagation.py @ 48b7917e), so the parity tests read like the reference's own:

  unitary_provider / state_provider / solver_dict / step_dict   registries   (:18-68)
  pwc(model, gen, instr, folding_stack, batch_size=None) -> {"U","dUs","ts"}   (:258-341)
  tf_batch_propagate(hamiltonian, hks, signals, dt, batch_size, col_ops=None, lindbladian=False)  (:460-515)
  tf_propagation_vectorized(h0, hks, cflds_t, dt)                               (:426-440)
  tf_propagation_lind(h0, hks, col_ops, cflds_t, dt)                            (:551-585)
  evaluate_sequences(propagators, sequences)                                    (:588-627)

plus the batch axis the reference lacks (its callers loop serially over ORBIT sequences,
noise trajectories and optimiser samples): ``signals`` may be ``[B,K,N]`` and
:func:`pwc_batch` propagates a whole batch in one kernel launch.

Tensors are torch complex128 CUDA tensors (``numpy()``-convertible after ``.cpu()``).  There is no
CPU path: every function raises if the CUDA library or a GPU is missing.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import engine

unitary_provider: Dict[str, Callable] = dict()
state_provider: Dict[str, Callable] = dict()
solver_dict: Dict[str, Callable] = dict()
step_dict: Dict[str, Callable] = dict()


def unitary_deco(func):
    """Decorator for making registry of functions (propagation.py:39-44)."""
    unitary_provider[str(func.__name__)] = func
    return func


def state_deco(func):
    state_provider[str(func.__name__)] = func
    return func


def solver_deco(func):
    solver_dict[str(func.__name__)] = func
    return func


def step_deco(func):
    step_dict[str(func.__name__)] = func
    return func


def _np(x):
    """Reference-style input (numpy, list, tf tensor, torch tensor) -> numpy array or torch tensor."""
    if isinstance(x, torch.Tensor):
        return x
    if hasattr(x, "numpy") and not isinstance(x, np.ndarray):
        return np.asarray(x.numpy())
    return np.asarray(x)


def _host(x) -> np.ndarray:
    """Same, but always a host numpy array (small model matrices, time stamps)."""
    x = _np(x)
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else x


def _real_signals(signals):
    """The reference casts real control fields to complex128 before use (:287-293, :429);
    the engine keeps them float64.  Complex inputs must have zero imaginary part."""
    if isinstance(signals, torch.Tensor):
        if signals.is_complex():
            if bool((signals.imag != 0).any()):
                raise ValueError("C3:ERROR: control signals must be real")
            signals = signals.real
        return signals.to(torch.float64)
    s = np.asarray(signals)
    if np.iscomplexobj(s):
        if np.any(s.imag != 0):
            raise ValueError("C3:ERROR: control signals must be real")
        s = s.real
    return np.ascontiguousarray(s, dtype=np.float64)


# --------------------------------------------------------------------------------------------
# array-level API
# --------------------------------------------------------------------------------------------

def _slice_hamiltonians(h0, hks, cflds_t) -> torch.Tensor:
    """H_n = h0[n] + sum_k c_k[n] hks[k] for a per-slice drift ``h0 [N,d,d]`` (host-side staging of an H list)."""
    h = torch.as_tensor(_host(h0), dtype=torch.complex128)
    hk = torch.as_tensor(_host(hks), dtype=torch.complex128)
    c = torch.as_tensor(_host(_real_signals(cflds_t)), dtype=torch.complex128)
    if h.shape[0] != c.shape[-1]:
        raise ValueError(f"C3:ERROR: {h.shape[0]} drift slices for {c.shape[-1]} signal samples")
    return h + torch.einsum("kn,kij->nij", c, hk)

def tf_propagation_vectorized(h0, hks, cflds_t, dt):
    """dU_n = expm(-i (h0 + sum_k c_k[n] hks[k]) dt) for every slice -> [n,d,d]
    (propagation.py:426-440).  With ``hks is None`` ``h0`` is the list of Hamiltonians [n,d,d]."""
    if hks is not None and cflds_t is not None:
        if _np(h0).ndim == 3:
            _, dUs = engine.pwc_closed_hlist(_slice_hamiltonians(h0, hks, cflds_t), float(np.real(dt)), return_dUs=True)
            return dUs[0]
        _, dUs = engine.pwc_closed(_np(h0), _np(hks), _real_signals(cflds_t), float(np.real(dt)), return_dUs=True)
    else:
        _, dUs = engine.pwc_closed_hlist(_np(h0), float(np.real(dt)), return_dUs=True)
    return dUs[0]


def tf_propagation_lind(h0, hks, col_ops, cflds_t, dt, history=False):
    """dU_n = expm(L_n dt) with the Lindblad superoperator L_n -> [n,d^2,d^2] (propagation.py:551-585)."""
    if hks is None or cflds_t is None:
        # dead in the reference as well: tf.cast(None) at propagation.py:552 via :511
        raise Exception("C3:ERROR: Lindblad propagation needs control Hamiltonians and signals.")
    _, dUs = engine.pwc_lindblad(_np(h0), _np(hks), [_np(c) for c in col_ops], _real_signals(cflds_t),
                                 float(np.real(dt)), return_dUs=True)
    return dUs[0]


def tf_batch_propagate(hamiltonian, hks, signals, dt, batch_size, col_ops=None, lindbladian=False):
    """All slice propagators dUs [N,d,d] (or [B,N,d,d] for batched ``signals [B,K,N]``).

    The reference chunks the TIME axis into ceil(N / batch_size) pieces only to bound host
    memory (propagation.py:482-515); the fused kernel has no such temporaries, so
    ``batch_size`` is accepted and ignored -- the result is identical.
    """
    del batch_size
    if signals is not None:
        sig = _real_signals(signals)
        batched = sig.ndim == 3
        if lindbladian:
            _, dUs = engine.pwc_lindblad(_np(hamiltonian), _np(hks), [_np(c) for c in col_ops], sig,
                                         float(np.real(dt)), return_dUs=True)
        elif _np(hamiltonian).ndim == 3 and not batched:
            # a per-slice drift [N,d,d] next to the control terms (the reference's `len(h0.shape) < 3` test,
            # propagation.py:430-436): assemble the slice Hamiltonians and take the H-list entry
            _, dUs = engine.pwc_closed_hlist(_slice_hamiltonians(hamiltonian, hks, sig), float(np.real(dt)), return_dUs=True)
        else:
            _, dUs = engine.pwc_closed(_np(hamiltonian), _np(hks), sig, float(np.real(dt)), return_dUs=True)
        return dUs if batched else dUs[0]
    if lindbladian:
        raise Exception("C3:ERROR: Lindblad propagation needs control Hamiltonians and signals.")
    h = _np(hamiltonian)
    _, dUs = engine.pwc_closed_hlist(h, float(np.real(dt)), return_dUs=True)
    return dUs if h.ndim == 4 else dUs[0]


def pwc_batch(h0, hks, signals, dt, col_ops=None, lindbladian=False, return_dUs=False):
    """The batched fast path: U [B,D,D] for ``signals [B,K,N]`` in one launch, no ``dUs`` store
    unless asked.  This is what B serial reference calls of ``pwc`` compute."""
    sig = _real_signals(signals)
    if lindbladian:
        return engine.pwc_lindblad(_np(h0), _np(hks), [_np(c) for c in col_ops], sig, float(np.real(dt)),
                                   return_dUs=return_dUs)
    host = not (isinstance(sig, torch.Tensor) and sig.is_cuda)
    if host and not return_dUs and sig.ndim == 3:
        # host-resident signals: chunked copy/compute pipeline (PCIe hidden behind the kernel)
        return engine.pwc_closed_from_host(_np(h0), _np(hks), sig, float(np.real(dt)))
    return engine.pwc_closed(_np(h0), _np(hks), sig, float(np.real(dt)), return_dUs=return_dUs)


class _PwcClosedFn(torch.autograd.Function):
    """U = pwc_batch(h0, hks, signals, dt), differentiable w.r.t. ``signals``."""

    @staticmethod
    def forward(ctx, signals, h0, hks, dt, hermitian):
        # where a fused gradient kernel serves the shape, the forward pass keeps its chunk products for the backward pass
        # (forward + backward = 1 + 3 forward passes' worth of work instead of 1 + 4)
        U, ctx.saved = engine.pwc_closed_saving(h0, hks, signals.detach(), dt, hermitian=hermitian)
        ctx.save_for_backward(signals.detach())
        ctx.h0, ctx.hks, ctx.dt = h0, hks, dt
        return U

    @staticmethod
    def backward(ctx, grad_U):
        (signals,) = ctx.saved_tensors
        if ctx.saved is not None:
            g = engine.pwc_closed_grad_saved(signals, grad_U.contiguous(), ctx.saved)
        else:
            _, g = engine.pwc_closed_grad(ctx.h0, ctx.hks, signals, ctx.dt, grad_U.contiguous())
        return g.to(signals.dtype).reshape(signals.shape), None, None, None, None


class _PwcLindbladFn(torch.autograd.Function):
    """U = pwc_batch(..., lindbladian=True), differentiable w.r.t. ``signals`` (d^2 <= 128)."""

    @staticmethod
    def forward(ctx, signals, h0, hks, col_ops, dt):
        U = engine.pwc_lindblad(h0, hks, col_ops, signals.detach(), dt)
        ctx.save_for_backward(signals.detach())
        ctx.h0, ctx.hks, ctx.col_ops, ctx.dt = h0, hks, col_ops, dt
        return U

    @staticmethod
    def backward(ctx, grad_U):
        (signals,) = ctx.saved_tensors
        _, g = engine.pwc_lindblad_grad(ctx.h0, ctx.hks, ctx.col_ops, signals, ctx.dt, grad_U.contiguous())
        return g.to(signals.dtype).reshape(signals.shape), None, None, None, None


def pwc_batch_autograd(h0, hks, signals: torch.Tensor, dt, col_ops=None, lindbladian: bool = False) -> torch.Tensor:
    """Differentiable batched propagators: ``signals`` is a CUDA float64 tensor [B,K,N] with
    ``requires_grad``; gradients of any real loss of U flow back to it (what the reference gets from
    tf.GradientTape, c3/optimizers/optimizer.py:210-215).  Shared model; closed system (d <= 128) or, with
    ``lindbladian=True`` and ``col_ops``, the Lindblad superoperator propagators (d^2 <= 128, e.g. D = 81)."""
    if not (isinstance(signals, torch.Tensor) and signals.is_cuda):
        raise ValueError("C3:ERROR: pwc_batch_autograd needs a CUDA tensor for `signals`")
    dev = signals.device
    hermitian = engine._hermitian_host(h0, hks) if not lindbladian else None       # host arrays: inspected before they move
    h0_t = torch.as_tensor(_host(h0), dtype=torch.complex128).to(dev) if not isinstance(h0, torch.Tensor) else h0.to(dev)
    hks_t = torch.as_tensor(_host(hks), dtype=torch.complex128).to(dev) if not isinstance(hks, torch.Tensor) else hks.to(dev)
    sig = signals if signals.dim() == 3 else signals.unsqueeze(0)
    if lindbladian:
        cols = torch.stack([torch.as_tensor(_host(c), dtype=torch.complex128) for c in col_ops]).to(dev)
        return _PwcLindbladFn.apply(sig, h0_t, hks_t, cols, float(np.real(dt)))
    return _PwcClosedFn.apply(sig, h0_t, hks_t, float(np.real(dt)), hermitian)


# --------------------------------------------------------------------------------------------
# gate-level API (duck-typed Model / Generator / Instruction exactly as the reference uses them)
# --------------------------------------------------------------------------------------------

class GateInputs:
    """What one gate hands to the engine, gathered from the duck-typed Model / Generator / Instruction exactly as the
    reference's ``pwc`` gathers it (propagation.py:282-321): either control fields + control Hamiltonians
    (``signals [K,N]``, ``hks [K,d,d]``) or, with ``model.controllability`` off, the list of slice Hamiltonians
    (``hlist [N,d,d]``); collapse operators when the model is Lindbladian; the excitation cutter if one is set."""

    __slots__ = ("h0", "hks", "signals", "hlist", "col_ops", "ts", "dt", "cutter", "channels")

    def __init__(self):
        self.h0 = self.hks = self.signals = self.hlist = self.col_ops = self.ts = self.cutter = None
        self.dt = 0.0
        self.channels = ()

    @property
    def n_slices(self) -> int:
        return int(self.signals.shape[-1] if self.signals is not None else self.hlist.shape[0])


def gather_gate(model, gen, instr) -> GateInputs:
    """Signals, Hamiltonians, time grid and collapse operators of one instruction (propagation.py:282-321)."""
    g = GateInputs()
    signal = gen.generate_signals(instr)
    g.channels = tuple(signal.keys())
    if model.controllability:
        h0, hctrls = model.get_Hamiltonians()
        g.h0 = _np(h0)
        # channel order = iteration order of the signal dictionary (propagation.py:289-292)
        g.signals = _stack_fields([signal[key]["values"] for key in g.channels])
        g.hks = np.stack([_host(hctrls[key]) for key in g.channels])
        g.ts = signal[g.channels[-1]]["ts"]
        ts_np = _host(g.ts)
    else:
        g.hlist = _np(model.get_Hamiltonian(signal))
        ts_all = np.asarray([_host(sig["ts"])[1:] for sig in signal.values()])
        ts_np = ts_all.mean(axis=0)
        g.ts = ts_np
        step = ts_np[1] - ts_np[0]
        # all lines on one grid, and that grid uniform (propagation.py:301-308)
        if not np.all(ts_all.var(axis=0) < 1e-5 * step) or not np.all(np.var(np.diff(ts_np)) < 1e-5 * step):
            raise Exception("C3Error:Something with the times happend.")
    g.dt = float(ts_np[1] - ts_np[0])
    if model.max_excitations:
        g.cutter = _host(model.ex_cutter)
    if model.lindbladian:
        if g.signals is None:
            raise Exception("C3:ERROR: Lindblad propagation needs control Hamiltonians and signals.")
        cols = [_host(c) for c in model.get_Lindbladians()]
        if g.cutter is not None:
            cols = [g.cutter @ c @ g.cutter.T for c in cols]
        g.col_ops = cols
    return g


def _stack_fields(values):
    """[K,N] float64 control fields from per-line arrays; stays on the device when the generator produced CUDA tensors."""
    if all(isinstance(v, torch.Tensor) for v in values):
        return _real_signals(torch.stack([v.reshape(-1) for v in values]))
    return _real_signals(np.stack([_host(v) for v in values]))


@unitary_deco
def pwc(model, gen, instr, folding_stack: list, batch_size=None) -> Dict:
    """Solve the equation of motion (Lindblad or Schroedinger) for one gate
    (propagation.py:258-341).  ``folding_stack`` and ``batch_size`` are accepted for call
    compatibility (c3/experiment.py:472-478); the ordered product is folded on chip.

    Returns ``{"U": [D,D], "dUs": [N,D,D], "ts": [N]}`` (torch CUDA tensors, ``ts`` as given).
    """
    del folding_stack, batch_size
    g = gather_gate(model, gen, instr)
    if g.col_ops is not None:
        U, dUs = engine.pwc_lindblad(g.h0, g.hks, g.col_ops, g.signals, g.dt, return_dUs=True)
    elif g.signals is not None:
        U, dUs = engine.pwc_closed(g.h0, g.hks, g.signals, g.dt, return_dUs=True)
    else:
        U, dUs = engine.pwc_closed_hlist(g.hlist, g.dt, return_dUs=True)
    U, dUs = U[0], dUs[0]
    if g.cutter is not None:
        # blow-up P^T A P is a scatter of the cut matrix into the full space (c3/model.py:222-224)
        U = blowup_excitations(g.cutter, U)
        dUs = blowup_excitations(g.cutter, dUs)
    return {"U": U, "dUs": dUs, "ts": g.ts}


def blowup_excitations(cutter, op: torch.Tensor) -> torch.Tensor:
    """P^T op P for a 0/1 row-selection matrix P [d_cut, d] as an index scatter."""
    c = _host(cutter)
    if op.shape[-1] != c.shape[0] or op.shape[-2] != c.shape[0]:
        # e.g. a Lindblad superoperator of the cut space: P^T S P is not defined for it, and the reference's
        # model.blowup_excitations (c3/model.py:222-224) fails on the same shapes
        raise Exception(f"C3:ERROR: cannot blow up a {tuple(op.shape[-2:])} operator with a {tuple(c.shape)} excitation cutter.")
    keep = torch.as_tensor(np.argmax(np.real(c), axis=1), device=op.device)
    d_full = c.shape[1]
    out = torch.zeros(op.shape[:-2] + (d_full, d_full), dtype=op.dtype, device=op.device)
    out[..., keep[:, None], keep[None, :]] = op
    return out


def cut_excitations(cutter, op: torch.Tensor) -> torch.Tensor:
    """P op P^T as an index gather (c3/model.py:218-220)."""
    c = _host(cutter)
    keep = torch.as_tensor(np.argmax(np.real(c), axis=1), device=op.device)
    return op[..., keep[:, None], keep[None, :]]


def evaluate_sequences(propagators: Dict, sequences: list) -> list:
    """Total propagator of each gate sequence, multiplied from the left
    (``sequence = [U0, U1, U2]`` is applied as ``U2 U1 U0``); an empty sequence gives the
    identity (propagation.py:588-627).  One kernel launch for all sequences."""
    names = list(propagators.keys())
    lookup = {n: i for i, n in enumerate(names)}
    first = propagators[names[0]]
    dev = first.device if isinstance(first, torch.Tensor) and first.is_cuda else engine.default_device()
    gates = torch.stack([torch.as_tensor(_np(propagators[n])).to(torch.complex128).to(dev) for n in names])
    S = len(sequences)
    if S == 0:
        return []
    lens = np.array([len(s) for s in sequences], dtype=np.int32)
    Lmax = int(lens.max())
    idx = np.zeros((S, max(Lmax, 1)), dtype=np.int32)
    for i, seq in enumerate(sequences):
        for j, g in enumerate(seq):
            idx[i, j] = lookup[g]
    out = engine.seq_product(gates, idx if Lmax > 0 else np.zeros((S, 0), np.int32), lens, device=dev)
    return [out[i] for i in range(S)]
