"""Tensor-level entry points of the B200 propagator engine.

Thin host code over the C ABI (include/c3b200.h): checks shapes/dtypes, owns the output and
workspace tensors (the C library never allocates) and passes torch's current CUDA stream.
PyTorch is used for device memory and streams only.  Everything here requires the CUDA
library and a CUDA device; there is no CPU path.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

_workspaces: Dict[Tuple[int, int], torch.Tensor] = {}


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _as(t, dtype, device) -> torch.Tensor:
    """Contiguous tensor of the wanted dtype on the wanted device (accepts numpy/lists)."""
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(t)
    if t.dtype != dtype:
        if dtype == torch.float64 and t.is_complex():
            t = t.real
        t = t.to(dtype)
    return t.to(device, non_blocking=True).contiguous()


def default_device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("C3:ERROR: c3_b200 needs a CUDA device (there is no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    """Per-(device, stream) scratch buffer, grown on demand and reused across calls."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def release_workspaces() -> None:
    _workspaces.clear()


_TUNING_SET = {}      # knobs this thread has set through set_tuning (the library keeps the values; this mirrors the non-default ones)


def _tuning_value(key: str, default: int = -1) -> int:
    import threading
    return _TUNING_SET.get((threading.get_ident(), key), default)


def set_tuning(key: str, value: int) -> None:
    import threading
    _lib.check(_lib.load().c3b_set_tuning(key.encode(), int(value)))
    _TUNING_SET[(threading.get_ident(), key)] = int(value)


def pwc_path(K: int, D: int, batched_model: bool = False) -> int:
    return _lib.load().c3b_pwc_path(int(K), int(D), int(batched_model))


def pwc_closed(h0, hks, signals, dt: float, return_dUs: bool = False, device=None, out=None):
    """U[b] = prod_n expm(-i (h0 + sum_k signals[b,k,n] hks[k]) dt), later slices on the left.

    h0 [d,d] (or [B,d,d]), hks [K,d,d] (or [B,K,d,d]), signals [B,K,N] (or [K,N] -> B=1).
    Returns U [B,d,d] and, if ``return_dUs``, dUs [B,N,d,d].
    (c3/libraries/propagation.py:426-440,460-515; c3/utils/tf_utils.py:144-193)
    """
    lib = _lib.load()
    device = torch.device(device) if device is not None else default_device()
    with torch.cuda.device(device):
        signals = _as(signals, torch.float64, device)
        if signals.dim() == 2:
            signals = signals.unsqueeze(0)
        B, K, N = signals.shape
        h0 = _as(h0, torch.complex128, device)
        batched = h0.dim() == 3
        d = h0.shape[-1]
        if K > 0:
            hks = _as(hks, torch.complex128, device)
            want = (B, K, d, d) if batched else (K, d, d)
            if tuple(hks.shape) != want:
                raise ValueError(f"C3:ERROR: hks has shape {tuple(hks.shape)}, expected {want}")
        else:
            hks = None
        if batched and h0.shape[0] != B:
            raise ValueError("C3:ERROR: batched h0 must have the batch size of signals")
        if out is not None:
            if tuple(out.shape) != (B, d, d) or out.dtype != torch.complex128 or not out.is_contiguous() or out.device != device:
                raise ValueError("C3:ERROR: `out` must be a contiguous complex128 [B,d,d] tensor on the compute device")
            U = out
        else:
            U = torch.empty((B, d, d), dtype=torch.complex128, device=device)
        dUs = torch.empty((B, N, d, d), dtype=torch.complex128, device=device) if return_dUs else None
        nbytes = lib.c3b_pwc_workspace_bytes(B, K, N, d, 0, int(batched))
        ws = _workspace(nbytes, device)
        _lib.check(lib.c3b_pwc_closed(_ptr(h0), _ptr(hks), _ptr(signals), float(dt), B, K, N, d, int(batched),
                                      _ptr(U), _ptr(dUs), _ptr(ws), ws.numel(), _stream()))
    return (U, dUs) if return_dUs else U


def pwc_closed_from_host(h0, hks, signals_host, dt: float, chunk: int = 1024, first_chunk: int = 256, device=None,
                         gated: bool = True):
    """Closed-system propagators for HOST-resident signals [B,K,N] (numpy or CPU tensor, ideally pinned).

    d = 9 (the headline kernel): ONE persistent launch over the whole batch, gated by a device counter.  The host
    enqueues the chunked host->device copies on a copy stream, each followed by a 4-byte copy that raises the counter
    to the number of batch rows that have landed; the kernel starts once the first (small) chunk is in and its warps,
    which take batch rows in order, wait only if they overtake the copy engine (``c3b_pwc_closed_gated``).
    Other d: the batch is cut into chunks and the copy of chunk i+1 overlaps the kernel of chunk i (two streams).
    Returns U [B,d,d] on the device, ordered on the caller's current stream."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else default_device()
    if not isinstance(signals_host, torch.Tensor):
        signals_host = torch.as_tensor(signals_host)
    if signals_host.is_cuda:
        return pwc_closed(h0, hks, signals_host, dt, device=device)
    if signals_host.dim() == 2:
        signals_host = signals_host.unsqueeze(0)
    signals_host = signals_host.to(torch.float64).contiguous()
    B, K, N = signals_host.shape
    with torch.cuda.device(device):
        h0 = _as(h0, torch.complex128, device)
        hks = _as(hks, torch.complex128, device) if K > 0 else None
        d = h0.shape[-1]
        U = torch.empty((B, d, d), dtype=torch.complex128, device=device)
        chunk = max(1, int(chunk))
        if B <= chunk or h0.dim() == 3 or K == 0:
            return pwc_closed(h0, hks, signals_host.to(device, non_blocking=True), dt, device=device, out=U)
        main = torch.cuda.current_stream()
        streams = _side_streams(device)
        start = torch.cuda.Event()
        start.record(main)
        if gated and lib.c3b_pwc_gated_supported(int(d)):
            if tuple(hks.shape) != (K, d, d):
                raise ValueError(f"C3:ERROR: hks has shape {tuple(hks.shape)}, expected {(K, d, d)}")
            bounds = [0, min(B, max(1, int(first_chunk)))]
            while bounds[-1] < B:
                bounds.append(min(B, bounds[-1] + chunk))
            U.fill_(float("nan"))            # rows the kernel never gets to see (copy failure) stay NaN
            sig = torch.empty((B, K, N), dtype=torch.float64, device=device)
            gate = torch.empty((2,), dtype=torch.int32, device=device)     # [rows landed, timeout flag]
            marks = _gate_marks(tuple(bounds[1:]))
            cs = streams[0]
            cs.wait_event(start)
            first = torch.cuda.Event()
            with torch.cuda.stream(cs):
                gate.zero_()
                for i in range(len(bounds) - 1):
                    sig[bounds[i]:bounds[i + 1]].copy_(signals_host[bounds[i]:bounds[i + 1]], non_blocking=True)
                    gate[0:1].copy_(marks[i:i + 1], non_blocking=True)
                    if i == 0:
                        first.record(cs)
            sig.record_stream(cs)
            gate.record_stream(cs)
            main.wait_event(first)          # every copy is already enqueued: the kernel can only wait on the copy engine
            nbytes = lib.c3b_pwc_workspace_bytes(B, K, N, d, 0, 0)
            ws = _workspace(nbytes, device)
            _lib.check(lib.c3b_pwc_closed_gated(_ptr(h0), _ptr(hks), _ptr(sig), float(dt), B, K, N, d, _ptr(U), _ptr(gate),
                                                _ptr(ws), ws.numel(), _stream()))
            _watch_gate(gate, main)
            return U
        done = []
        starts = [0] + list(range(min(B, max(1, int(first_chunk))), B, chunk)) if not gated else list(range(0, B, chunk))
        for i, b0 in enumerate(starts):
            b1 = starts[i + 1] if i + 1 < len(starts) else B
            st = streams[i % 2]
            st.wait_event(start)
            with torch.cuda.stream(st):
                sig = signals_host[b0:b1].to(device, non_blocking=True)
                pwc_closed(h0, hks, sig, dt, device=device, out=U[b0:b1])
                sig.record_stream(st)
                ev = torch.cuda.Event()
                ev.record(st)
                done.append(ev)
        for ev in done:
            main.wait_event(ev)
        U.record_stream(main)
    return U


_marks_cache = {}


def _gate_marks(bounds) -> torch.Tensor:
    """Pinned int32 row counts of the gated launch (cached: pinning costs a cudaHostAlloc, and the contents never change)."""
    t = _marks_cache.get(bounds)
    if t is None:
        if len(_marks_cache) > 64:
            _marks_cache.clear()
        t = torch.tensor(bounds, dtype=torch.int32).pin_memory()
        _marks_cache[bounds] = t
    return t


_gates_in_flight = []
_flag_pool = []


def _watch_gate(gate: torch.Tensor, stream) -> None:
    """Queue an asynchronous read-back of a gated launch's status word (gate[1]) into pinned host memory behind the
    kernel; check_gated_launches looks at it once the copy has completed.  Nothing here blocks the host."""
    flag = _flag_pool.pop() if _flag_pool else torch.zeros((1,), dtype=torch.int32).pin_memory()
    with torch.cuda.stream(stream):
        flag.copy_(gate[1:2], non_blocking=True)
    ev = torch.cuda.Event()
    ev.record(stream)
    _gates_in_flight.append((ev, flag, gate))
    check_gated_launches(wait=False)


def check_gated_launches(wait: bool = True) -> None:
    """Raise if a gated launch (pwc_closed_from_host) gave up on a batch row that never arrived from the host: the
    kernel sets gate[1] and leaves the affected rows of U as NaN instead of hanging the GPU.  ``wait=False`` (what every
    later gated call does on entry) only looks at launches that have already finished; ``wait=True`` synchronises on the
    outstanding ones first -- call it wherever the result is consumed on the host."""
    keep = []
    failed = False
    for ev, flag, gate in _gates_in_flight:
        if wait:
            ev.synchronize()
        if ev.query():
            failed = failed or bool(int(flag[0]) != 0)
            _flag_pool.append(flag)
        else:
            keep.append((ev, flag, gate))
    _gates_in_flight[:] = keep
    if failed:
        raise _lib.C3BError("C3:ERROR: gated launch timed out waiting for host control fields; the affected rows of U are NaN")


_side = {}


def _side_streams(device):
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _side:
        _side[key] = (torch.cuda.Stream(device=device), torch.cuda.Stream(device=device))
    return _side[key]


def pwc_closed_grad(h0, hks, signals, dt: float, Ubar, max_workspace_bytes: int = 6 << 30, device=None):
    """Forward U and the gradient of a real scalar loss w.r.t. the control fields.

    ``Ubar`` [B,d,d] is the cotangent of U in torch's convention (dL = Re tr(Ubar^dag dU), i.e. what
    ``autograd`` hands to ``backward``).  Returns (U [B,d,d], grad [B,K,N] float64).
    Replaces tf.GradientTape through the propagator (c3/optimizers/optimizer.py:210-215)."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else default_device()
    with torch.cuda.device(device):
        signals = _as(signals, torch.float64, device)
        if signals.dim() == 2:
            signals = signals.unsqueeze(0)
        B, K, N = signals.shape
        # Hamiltonians handed over as host arrays: look at them here, so that the library need not read a flag back from the
        # device (a stream synchronisation per call) to choose the fused unitary-recurrence kernels
        hermitian = _hermitian_host(h0, hks)
        h0 = _as(h0, torch.complex128, device)
        hks = _as(hks, torch.complex128, device)
        if h0.dim() != 2:
            raise ValueError("C3:ERROR: the gradient path needs a shared model h0 [d,d]")
        d = h0.shape[-1]
        Ubar = _as(Ubar, torch.complex128, device)
        if tuple(Ubar.shape) != (B, d, d):
            raise ValueError(f"C3:ERROR: Ubar has shape {tuple(Ubar.shape)}, expected {(B, d, d)}")
        U = torch.empty((B, d, d), dtype=torch.complex128, device=device)
        grad = torch.empty((B, K, N), dtype=torch.float64, device=device)
        assert_flag = hermitian is not None and _tuning_value("grad_unitary") == -1
        if assert_flag:
            set_tuning("grad_unitary", 1 if hermitian else 0)
        try:
            chunk = B
            while chunk > 1 and lib.c3b_pwc_grad_workspace_bytes(B, K, N, d, chunk) > max_workspace_bytes:
                chunk = (chunk + 1) // 2
            nbytes = lib.c3b_pwc_grad_workspace_bytes(B, K, N, d, chunk)
            ws = _workspace(nbytes, device)
            _lib.check(lib.c3b_pwc_closed_grad(_ptr(h0), _ptr(hks), _ptr(signals), float(dt), B, K, N, d, _ptr(Ubar),
                                               _ptr(grad), _ptr(U), chunk, _ptr(ws), ws.numel(), _stream()))
        finally:
            if assert_flag:
                set_tuning("grad_unitary", -1)
    return U, grad


class SavedForward:
    """What pwc_closed_saving leaves for pwc_closed_grad_saved: the state buffer (model, chunk products, prefix products,
    U) and the chunking of the forward call."""
    __slots__ = ("state", "B", "K", "N", "d", "Q", "CL")


def _hermitian_host(h0, hks) -> Optional[bool]:
    """True / False for Hamiltonians handed over as host arrays, None when they live on the device."""
    if isinstance(h0, torch.Tensor) or isinstance(hks, torch.Tensor):
        return None
    a0, ak = np.asarray(h0), np.asarray(hks)
    if a0.ndim != 2 or ak.ndim != 3:
        return None
    scale = max(float(np.abs(a0).max(initial=0.0)), float(np.abs(ak).max(initial=0.0)))
    return bool(np.abs(a0 - a0.conj().T).max(initial=0.0) <= 1e-13 * scale
                and np.abs(ak - ak.conj().transpose(0, 2, 1)).max(initial=0.0) <= 1e-13 * scale)


def pwc_closed_saving(h0, hks, signals, dt: float, hermitian: Optional[bool] = None, device=None):
    """Forward pass that keeps what the backward pass needs: returns (U [B,d,d], SavedForward), or (U, None) where no fused
    gradient kernel serves the shape (the caller then differentiates with pwc_closed_grad).  ``hermitian``: the caller's
    knowledge about h0 / hks (None: host arrays are inspected here, device tensors by the library)."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else default_device()
    with torch.cuda.device(device):
        if hermitian is None:
            hermitian = _hermitian_host(h0, hks)
        signals = _as(signals, torch.float64, device)
        if signals.dim() == 2:
            signals = signals.unsqueeze(0)
        B, K, N = signals.shape
        h0 = _as(h0, torch.complex128, device)
        hks = _as(hks, torch.complex128, device)
        if h0.dim() != 2 or K == 0:
            return pwc_closed(h0, hks, signals, dt), None
        d = h0.shape[-1]
        assert_flag = hermitian is not None and _tuning_value("grad_unitary") == -1
        if assert_flag:
            set_tuning("grad_unitary", 1 if hermitian else 0)
        try:
            nbytes = int(lib.c3b_pwc_closed_saved_bytes(B, K, N, d))
            if nbytes == 0:
                return pwc_closed(h0, hks, signals, dt), None
            state = torch.empty(nbytes, dtype=torch.uint8, device=device)
            U = torch.empty((B, d, d), dtype=torch.complex128, device=device)
            rc = lib.c3b_pwc_closed_fwd_saved(_ptr(h0), _ptr(hks), _ptr(signals), float(dt), B, K, N, d, _ptr(U), _ptr(state), nbytes,
                                              _stream())
            if rc == _lib.C3B_EUNSUPPORTED:                  # non-Hermitian Hamiltonians found on the device
                return pwc_closed(h0, hks, signals, dt), None
            _lib.check(rc)
            import ctypes
            q, cl = ctypes.c_int(0), ctypes.c_int(0)
            _lib.check(lib.c3b_pwc_closed_saved_chunks(B, K, N, d, ctypes.byref(q), ctypes.byref(cl)))
        finally:
            if assert_flag:
                set_tuning("grad_unitary", -1)
        sv = SavedForward()
        sv.state, sv.B, sv.K, sv.N, sv.d, sv.Q, sv.CL = state, B, K, N, d, int(q.value), int(cl.value)
    return U, sv


def pwc_closed_grad_saved(signals, Ubar, saved: SavedForward) -> torch.Tensor:
    """grad [B,K,N] of a real loss w.r.t. the control fields from the cotangent ``Ubar`` of U and the state of
    pwc_closed_saving (same signals): no forward pass is repeated."""
    lib = _lib.load()
    device = saved.state.device
    with torch.cuda.device(device):
        signals = _as(signals, torch.float64, device)
        if signals.dim() == 2:
            signals = signals.unsqueeze(0)
        if tuple(signals.shape) != (saved.B, saved.K, saved.N):
            raise ValueError(f"C3:ERROR: signals have shape {tuple(signals.shape)}, the saved forward pass had {(saved.B, saved.K, saved.N)}")
        Ubar = _as(Ubar, torch.complex128, device)
        if tuple(Ubar.shape) != (saved.B, saved.d, saved.d):
            raise ValueError(f"C3:ERROR: Ubar has shape {tuple(Ubar.shape)}, expected {(saved.B, saved.d, saved.d)}")
        grad = torch.empty((saved.B, saved.K, saved.N), dtype=torch.float64, device=device)
        # the shape was served by a fused kernel in the forward call; keep the library's choice the same here
        set_flag = _tuning_value("grad_unitary") == -1
        if set_flag:
            set_tuning("grad_unitary", 1)
        try:
            _lib.check(lib.c3b_pwc_closed_bwd_saved(_ptr(signals), saved.B, saved.K, saved.N, saved.d, saved.Q, saved.CL, _ptr(Ubar),
                                                    _ptr(grad), _ptr(saved.state), saved.state.numel(), _stream()))
        finally:
            if set_flag:
                set_tuning("grad_unitary", -1)
    return grad


def pwc_lindblad_grad(h0, hks, col_ops, signals, dt: float, Ubar, max_workspace_bytes: int = 6 << 30, device=None):
    """Forward U [B,D,D] (D = d^2) and the gradient [B,K,N] of a real scalar loss w.r.t. the control fields through the
    Lindblad superoperator propagators; ``Ubar`` as in :func:`pwc_closed_grad`.  Any d with d^2 <= 128 (warp kernels up to
    d^2 = 16, CTA kernels on the DMMA product above)."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else default_device()
    with torch.cuda.device(device):
        signals = _as(signals, torch.float64, device)
        if signals.dim() == 2:
            signals = signals.unsqueeze(0)
        B, K, N = signals.shape
        h0 = _as(h0, torch.complex128, device)
        hks = _as(hks, torch.complex128, device)
        if h0.dim() != 2:
            raise ValueError("C3:ERROR: the gradient path needs a shared model h0 [d,d]")
        d = h0.shape[-1]
        D = d * d
        col_ops, C = _collapse_ops(col_ops, 0, d, device)
        Ubar = _as(Ubar, torch.complex128, device)
        if tuple(Ubar.shape) != (B, D, D):
            raise ValueError(f"C3:ERROR: Ubar has shape {tuple(Ubar.shape)}, expected {(B, D, D)}")
        U = torch.empty((B, D, D), dtype=torch.complex128, device=device)
        grad = torch.empty((B, K, N), dtype=torch.float64, device=device)
        chunk = B
        while chunk > 1 and lib.c3b_pwc_lindblad_grad_workspace_bytes(B, K, N, d, chunk) > max_workspace_bytes:
            chunk = (chunk + 1) // 2
        nbytes = lib.c3b_pwc_lindblad_grad_workspace_bytes(B, K, N, d, chunk)
        ws = _workspace(nbytes, device)
        _lib.check(lib.c3b_pwc_lindblad_grad(_ptr(h0), _ptr(hks), _ptr(col_ops), C, _ptr(signals), float(dt), B, K, N, d,
                                             _ptr(Ubar), _ptr(grad), _ptr(U), chunk, _ptr(ws), ws.numel(), _stream()))
    return U, grad


def pwc_closed_hlist(Hs, dt: float, return_dUs: bool = False, device=None):
    """Same with explicit Hamiltonians Hs [B,N,d,d] (or [N,d,d]); the reference's
    ``signals is None`` mode (c3/libraries/propagation.py:294-308, 437-438)."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else default_device()
    with torch.cuda.device(device):
        Hs = _as(Hs, torch.complex128, device)
        if Hs.dim() == 3:
            Hs = Hs.unsqueeze(0)
        B, N, d, _ = Hs.shape
        U = torch.empty((B, d, d), dtype=torch.complex128, device=device)
        dUs = torch.empty((B, N, d, d), dtype=torch.complex128, device=device) if return_dUs else None
        nbytes = lib.c3b_pwc_workspace_bytes(B, 0, N, d, 0, 0)
        ws = _workspace(nbytes, device)
        _lib.check(lib.c3b_pwc_closed_hlist(_ptr(Hs), float(dt), B, N, d, _ptr(U), _ptr(dUs), _ptr(ws), ws.numel(),
                                            _stream()))
    return (U, dUs) if return_dUs else U


def _collapse_ops(col_ops, n_models: int, d: int, device):
    """Collapse operators in the layout the C ABI reads: [C,d,d] for a shared model (``n_models == 0``), [B,C,d,d] for
    per-sample models.  Accepts a list of [d,d] (or, per sample, of [B,d,d]) or a stacked tensor; operators shared by all
    samples of a batched model are expanded."""
    if col_ops is None or (isinstance(col_ops, (list, tuple)) and len(col_ops) == 0):
        return None, 0
    if isinstance(col_ops, (list, tuple)):
        ops = [_as(c, torch.complex128, device) for c in col_ops]
        col_ops = torch.stack(ops, dim=1 if ops[0].dim() == 3 else 0)
    else:
        col_ops = _as(col_ops, torch.complex128, device)
    if col_ops.dim() == 2:
        col_ops = col_ops.unsqueeze(0)
    if col_ops.shape[-2:] != (d, d):
        raise ValueError(f"C3:ERROR: collapse operators have shape {tuple(col_ops.shape)}, expected [..., {d}, {d}]")
    C = col_ops.shape[-3]
    if n_models:
        if col_ops.dim() == 3:
            col_ops = col_ops.unsqueeze(0).expand(n_models, C, d, d)
        if tuple(col_ops.shape) != (n_models, C, d, d):
            raise ValueError(f"C3:ERROR: collapse operators have shape {tuple(col_ops.shape)}, expected {(n_models, C, d, d)}")
    elif col_ops.dim() != 3:
        raise ValueError(f"C3:ERROR: collapse operators have shape {tuple(col_ops.shape)}, expected {(C, d, d)} for a shared model")
    return col_ops.contiguous(), C


def pwc_lindblad(h0, hks, col_ops, signals, dt: float, return_dUs: bool = False, device=None):
    """Lindblad superoperator propagators U [B,d^2,d^2] (c3/libraries/propagation.py:551-585)."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else default_device()
    with torch.cuda.device(device):
        signals = _as(signals, torch.float64, device)
        if signals.dim() == 2:
            signals = signals.unsqueeze(0)
        B, K, N = signals.shape
        h0 = _as(h0, torch.complex128, device)
        batched = h0.dim() == 3
        d = h0.shape[-1]
        if batched and h0.shape[0] != B:
            raise ValueError("C3:ERROR: batched h0 must have the batch size of signals")
        if K > 0:
            hks = _as(hks, torch.complex128, device)
            want = (B, K, d, d) if batched else (K, d, d)
            if tuple(hks.shape) != want:
                raise ValueError(f"C3:ERROR: hks has shape {tuple(hks.shape)}, expected {want}")
        else:
            hks = None
        col_ops, C = _collapse_ops(col_ops, B if batched else 0, d, device)
        D = d * d
        U = torch.empty((B, D, D), dtype=torch.complex128, device=device)
        dUs = torch.empty((B, N, D, D), dtype=torch.complex128, device=device) if return_dUs else None
        nbytes = lib.c3b_pwc_workspace_bytes(B, K, N, d, 1, int(batched))
        ws = _workspace(nbytes, device)
        _lib.check(lib.c3b_pwc_lindblad(_ptr(h0), _ptr(hks), _ptr(col_ops), C, _ptr(signals), float(dt), B, K, N, d,
                                        int(batched), _ptr(U), _ptr(dUs), _ptr(ws), ws.numel(), _stream()))
    return (U, dUs) if return_dUs else U


class PreparedModel:
    """Device-resident generators of one model (or of B per-sample models): what ``c3b_model_prepare`` builds once per model
    update so that every later propagation is one fused launch (the reference rebuilds them inside every
    tf_propagation_* call, c3/libraries/propagation.py:426-440, 551-585)."""

    def __init__(self, blob: torch.Tensor, K: int, d: int, dt: float, lindblad: bool, n_models: int):
        self.blob, self.K, self.d, self.dt, self.lindblad, self.n_models = blob, K, d, dt, lindblad, n_models

    @property
    def D(self) -> int:
        return self.d * self.d if self.lindblad else self.d

    @property
    def device(self) -> torch.device:
        return self.blob.device


def prepare_model(h0, hks, dt: float, col_ops=None, lindblad: bool = False, device=None) -> PreparedModel:
    """Generators for ``h0 [d,d]`` / ``hks [K,d,d]`` (or [B,d,d] / [B,K,d,d] for per-sample models) and slice length ``dt``;
    with ``lindblad`` the superoperator generators of ``col_ops``."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else default_device()
    with torch.cuda.device(device):
        h0 = _as(h0, torch.complex128, device)
        batched = h0.dim() == 3
        d = h0.shape[-1]
        n_models = h0.shape[0] if batched else 1
        K = 0
        if hks is not None:
            hks = _as(hks, torch.complex128, device)
            K = hks.shape[-3] if hks.numel() else 0
        if K > 0:
            want = (n_models, K, d, d) if batched else (K, d, d)
            if tuple(hks.shape) != want:
                raise ValueError(f"C3:ERROR: hks has shape {tuple(hks.shape)}, expected {want}")
        else:
            hks = None
        C = 0
        if lindblad:
            col_ops, C = _collapse_ops(col_ops, n_models if batched else 0, d, device)
        else:
            col_ops = None
        nbytes = lib.c3b_model_bytes(K, d, int(lindblad), n_models)
        blob = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _lib.check(lib.c3b_model_prepare(_ptr(h0), _ptr(hks), _ptr(col_ops), C, float(dt), K, d, int(lindblad), n_models,
                                         _ptr(blob), nbytes, _stream()))
    return PreparedModel(blob, K, d, float(dt), bool(lindblad), n_models)


def pwc_prepared(model: PreparedModel, signals, return_dUs: bool = False, out=None, dUs_out=None):
    """U [B,D,D] (and dUs [B,N,D,D]) for ``signals [B,K,N]`` with a prepared model: one fused launch (+ one fold launch when
    the time axis is segmented), no setup kernels."""
    lib = _lib.load()
    device = model.device
    with torch.cuda.device(device):
        signals = _as(signals, torch.float64, device)
        if signals.dim() == 2:
            signals = signals.unsqueeze(0)
        B, K, N = signals.shape
        if K != model.K:
            raise ValueError(f"C3:ERROR: signals have {K} control lines, the prepared model has {model.K}")
        if model.n_models not in (1, B):
            raise ValueError("C3:ERROR: a per-sample prepared model needs one signal row per model")
        D = model.D
        U = out if out is not None else torch.empty((B, D, D), dtype=torch.complex128, device=device)
        if tuple(U.shape) != (B, D, D) or U.dtype != torch.complex128 or not U.is_contiguous():
            raise ValueError("C3:ERROR: `out` must be a contiguous complex128 [B,D,D] tensor")
        dUs = None
        if return_dUs:
            dUs = dUs_out if dUs_out is not None else torch.empty((B, N, D, D), dtype=torch.complex128, device=device)
        nbytes = lib.c3b_pwc_prepared_workspace_bytes(B, N, model.d, int(model.lindblad), model.n_models)
        ws = _workspace(nbytes, device)
        _lib.check(lib.c3b_pwc_prepared(_ptr(model.blob), _ptr(signals), B, K, N, model.d, int(model.lindblad), model.n_models,
                                        _ptr(U), _ptr(dUs), _ptr(ws), ws.numel(), _stream()))
    return (U, dUs) if return_dUs else U


class GraphedPwc:
    """A prepared-model propagation of fixed shape captured in a CUDA graph: one graph launch per call, for the
    latency-bound small-batch calls of a per-gate optimiser loop (B = 1 ... a gate set).

    ``run(signals)`` copies the control fields into the graph's static input buffer (device-to-device, or from pinned
    host memory) and replays; the returned tensors are the graph's static outputs, valid until the next ``run``."""

    def __init__(self, model: PreparedModel, B: int, N: int, return_dUs: bool = False):
        self.model, self.B, self.N, self.return_dUs = model, int(B), int(N), bool(return_dUs)
        dev = model.device
        D = model.D
        with torch.cuda.device(dev):
            self.signals = torch.zeros((self.B, model.K, self.N), dtype=torch.float64, device=dev)
            self.U = torch.empty((self.B, D, D), dtype=torch.complex128, device=dev)
            self.dUs = torch.empty((self.B, self.N, D, D), dtype=torch.complex128, device=dev) if return_dUs else None
            # a private workspace: the shared per-stream one may be regrown (freed) by other calls between replays
            nbytes = _lib.load().c3b_pwc_prepared_workspace_bytes(self.B, self.N, model.d, int(model.lindblad), model.n_models)
            self._ws = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._launch()                       # warm-up outside capture (cudaFuncSetAttribute, lazy module load)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._launch()

    def _launch(self):
        lib = _lib.load()
        m = self.model
        _lib.check(lib.c3b_pwc_prepared(_ptr(m.blob), _ptr(self.signals), self.B, m.K, self.N, m.d, int(m.lindblad), m.n_models,
                                        _ptr(self.U), _ptr(self.dUs), _ptr(self._ws), self._ws.numel(), _stream()))

    def run(self, signals=None):
        if signals is not None:
            if not isinstance(signals, torch.Tensor):
                signals = torch.as_tensor(signals)
            self.signals.copy_(signals.reshape(self.signals.shape), non_blocking=True)
        self.graph.replay()
        return (self.U, self.dUs) if self.return_dUs else self.U


def ordered_product(mats, device=None) -> torch.Tensor:
    """out[b] = mats[b,M-1] ... mats[b,0]  for mats [B,M,D,D] (or [M,D,D] -> [D,D])."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else default_device()
    with torch.cuda.device(device):
        mats = _as(mats, torch.complex128, device)
        squeeze = mats.dim() == 3
        if squeeze:
            mats = mats.unsqueeze(0)
        B, M, D, _ = mats.shape
        out = torch.empty((B, D, D), dtype=torch.complex128, device=device)
        nbytes = lib.c3b_product_workspace_bytes(B, M, D)
        ws = _workspace(nbytes, device)
        _lib.check(lib.c3b_ordered_product(_ptr(mats), B, M, D, _ptr(out), _ptr(ws), ws.numel(), _stream()))
    return out[0] if squeeze else out


def seq_product(gates, seq_idx, seq_len, device=None) -> torch.Tensor:
    """out[s] = gates[idx[s,len_s-1]] ... gates[idx[s,0]] (identity when len_s == 0)."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else default_device()
    with torch.cuda.device(device):
        gates = _as(gates, torch.complex128, device)
        seq_idx = _as(seq_idx, torch.int32, device)
        seq_len = _as(seq_len, torch.int32, device)
        Gn, D, _ = gates.shape
        S = seq_len.shape[0]
        Lmax = seq_idx.shape[1] if seq_idx.dim() == 2 else 0
        out = torch.empty((S, D, D), dtype=torch.complex128, device=device)
        nbytes = lib.c3b_product_workspace_bytes(S, 1, D)
        ws = _workspace(nbytes, device)
        _lib.check(lib.c3b_seq_product(_ptr(gates), Gn, _ptr(seq_idx) if Lmax > 0 else None, _ptr(seq_len), S, Lmax,
                                       D, _ptr(out), _ptr(ws), ws.numel(), _stream()))
    return out


FID_MODES = {"unitary": 0, "average": 1, "lindbladian_unitary": 2, "lindbladian_average": 3}


def gate_infid(U, ideal, sel, mode: str = "unitary", return_overlap: bool = False, device=None):
    """infid[b] of every propagator U[b] against the ideal gate on the rows/columns ``sel``
    (c3/libraries/fidelities.py:152-183, 221-249, 288-311, 377-399).  U [B,D,D] or [D,D]."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else default_device()
    with torch.cuda.device(device):
        U = _as(U, torch.complex128, device)
        squeeze = U.dim() == 2
        if squeeze:
            U = U.unsqueeze(0)
        B, D, _ = U.shape
        ideal = _as(ideal, torch.complex128, device)
        sel = _as(sel, torch.int32, device)
        C = sel.shape[0]
        if tuple(ideal.shape) != (C, C):
            raise ValueError(f"C3:ERROR: ideal gate has shape {tuple(ideal.shape)}, expected {(C, C)}")
        infid = torch.empty((B,), dtype=torch.float64, device=device)
        ov = torch.empty((B,), dtype=torch.complex128, device=device) if return_overlap else None
        _lib.check(lib.c3b_gate_infid(_ptr(U), B, D, _ptr(ideal), _ptr(sel), C, FID_MODES[mode], _ptr(infid), _ptr(ov),
                                      _stream()))
    if squeeze:
        infid = infid[0]
    return (infid, ov) if return_overlap else infid


def gate_infid_grad(overlap, ideal, sel, gbar, D: int, mode: str = "unitary", device=None) -> torch.Tensor:
    """Cotangent Ubar [B,D,D] of the propagators for the loss sum_b gbar[b] infid[b]."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else default_device()
    with torch.cuda.device(device):
        overlap = _as(overlap, torch.complex128, device)
        ideal = _as(ideal, torch.complex128, device)
        sel = _as(sel, torch.int32, device)
        gbar = _as(gbar, torch.float64, device) if gbar is not None else None
        B, C = overlap.shape[0], sel.shape[0]
        Ubar = torch.empty((B, D, D), dtype=torch.complex128, device=device)
        _lib.check(lib.c3b_gate_infid_grad(_ptr(overlap), _ptr(ideal), _ptr(sel), _ptr(gbar), B, D, C, FID_MODES[mode],
                                           _ptr(Ubar), _stream()))
    return Ubar


def seq_populations(gates, seq_idx, seq_len, psi0=None, lindblad_d: int = 0, return_states: bool = False, device=None):
    """Populations [S,D] (Lindblad: [S,d]) of psi0 after each gate sequence (c3/experiment.py:273-302, 603-624)."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else default_device()
    with torch.cuda.device(device):
        gates = _as(gates, torch.complex128, device)
        seq_idx = _as(seq_idx, torch.int32, device)
        seq_len = _as(seq_len, torch.int32, device)
        Gn, D, _ = gates.shape
        S = seq_len.shape[0]
        Lmax = seq_idx.shape[1] if seq_idx.dim() == 2 else 0
        psi0 = _as(psi0, torch.complex128, device).reshape(-1) if psi0 is not None else None
        if psi0 is not None and psi0.shape[0] != D:
            raise ValueError(f"C3:ERROR: psi0 has {psi0.shape[0]} entries, expected {D}")
        pops = torch.empty((S, lindblad_d if lindblad_d else D), dtype=torch.float64, device=device)
        psi = torch.empty((S, D), dtype=torch.complex128, device=device) if return_states else None
        _lib.check(lib.c3b_seq_populations(_ptr(gates), Gn, _ptr(seq_idx) if Lmax > 0 else None, _ptr(seq_len), S, Lmax, D,
                                           _ptr(psi0), int(lindblad_d), _ptr(pops), _ptr(psi), _stream()))
    return (pops, psi) if return_states else pops


def signal_slice_num(t_start: float, t_end: float, resolution: float) -> int:
    """Device.calc_slice_num (c3/generator/devices.py:73-85)."""
    return int(_lib.load().c3b_signal_slice_num(float(t_start), float(t_end), float(resolution)))


NOISE_KEYS = ("awg_amp", "lo_perc", "add_amp", "dc_amp", "pink_amp", "bfl_num", "dc_offset")
NOISE_TRACES = ("awg_i", "awg_q", "lo_cos", "lo_sin", "add", "dc", "pink")


def generate_signals(env_params, env_shape, env_flags, lo_freq, chain, t_start: float, t_end: float, device=None,
                     out=None, noise=None, seed: int = 0, return_noise: bool = False, env_table=None):
    """signals [B,K,N] from pulse parameters (c3b_generate_signals; see include/c3b200.h for the layouts).

    ``env_table [K,E,T]``: the array parameters of the extended envelope shapes (pwc, fourier_*, slepian_fourier, ...).

    ``noise [K,7]`` or ``[B,K,7]`` (columns NOISE_KEYS) switches the noise devices on: one independent realisation per batch
    row, drawn from the counter-based generator keyed by ``seed`` (same seed -> same realisation).  ``return_noise`` also
    returns the realised traces ``[B,K,7,N]`` (NOISE_TRACES)."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else default_device()
    with torch.cuda.device(device):
        env_params = _as(env_params, torch.float64, device)
        B, K, E, P = env_params.shape
        if P != 9:
            raise ValueError("C3:ERROR: env_params must have 9 entries per envelope")
        env_shape = _as(env_shape, torch.int32, device)
        env_flags = _as(env_flags, torch.int32, device)
        lo_freq = _as(lo_freq, torch.float64, device)
        chain = _as(chain, torch.float64, device)
        batched = chain.dim() == 3
        if tuple(env_shape.shape) != (K, E) or tuple(env_flags.shape) != (K, E) or tuple(lo_freq.shape) != (B, K) \
                or tuple(chain.shape) != ((B, K, 11) if batched else (K, 11)):
            raise ValueError("C3:ERROR: inconsistent shapes in generate_signals")
        sim_res = float(chain.reshape(-1, 11)[0, 0])
        N = signal_slice_num(t_start, t_end, sim_res)
        if N <= 0:
            raise ValueError("C3:ERROR: empty time grid")
        if out is None:
            out = torch.empty((B, K, N), dtype=torch.float64, device=device)
        traces = None
        nbatched = 0
        if noise is not None:
            noise = _as(noise, torch.float64, device)
            nbatched = int(noise.dim() == 3)
            if tuple(noise.shape) != ((B, K, len(NOISE_KEYS)) if nbatched else (K, len(NOISE_KEYS))):
                raise ValueError(f"C3:ERROR: noise has shape {tuple(noise.shape)}, expected [K,{len(NOISE_KEYS)}] or [B,K,{len(NOISE_KEYS)}]")
            if return_noise:
                traces = torch.empty((B, K, len(NOISE_TRACES), N), dtype=torch.float64, device=device)
        elif return_noise:
            raise ValueError("C3:ERROR: return_noise needs noise parameters")
        T = 0
        if env_table is not None:
            env_table = _as(env_table, torch.float64, device)
            if env_table.dim() != 3 or tuple(env_table.shape[:2]) != (K, E):
                raise ValueError(f"C3:ERROR: env_table has shape {tuple(env_table.shape)}, expected [K={K},E={E},T]")
            T = int(env_table.shape[2])
        _lib.check(lib.c3b_generate_signals_table(_ptr(env_params), _ptr(env_shape), _ptr(env_flags), _ptr(env_table) if T else None, T,
                                                  _ptr(lo_freq), _ptr(chain), int(batched), float(t_start), float(t_end), B, K, E, N,
                                                  _ptr(noise), nbatched, int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(out), _ptr(traces),
                                                  _stream()))
    return (out, traces) if return_noise else out


def crosstalk(signals: torch.Tensor, chan, matrix) -> torch.Tensor:
    """IN PLACE: mix the drive lines ``chan [C]`` of ``signals [B,K,N]`` by ``matrix [C,C]`` (the Crosstalk device,
    c3/generator/devices.py:225-293).  Returns signals."""
    lib = _lib.load()
    if not (isinstance(signals, torch.Tensor) and signals.is_cuda and signals.dtype == torch.float64 and signals.is_contiguous()
            and signals.dim() == 3):
        raise ValueError("C3:ERROR: crosstalk needs a contiguous float64 CUDA tensor [B,K,N]")
    with torch.cuda.device(signals.device):
        B, K, N = signals.shape
        chan = _as(chan, torch.int32, signals.device).reshape(-1)
        C = int(chan.shape[0])
        matrix = _as(matrix, torch.float64, signals.device)
        if tuple(matrix.shape) != (C, C):
            raise ValueError(f"C3:ERROR: crosstalk matrix has shape {tuple(matrix.shape)}, expected {(C, C)}")
        if len(set(chan.tolist())) != C or not all(0 <= int(c) < K for c in chan.tolist()):
            raise ValueError("C3:ERROR: crosstalk channels must be distinct line indices")
        _lib.check(lib.c3b_crosstalk(_ptr(signals), B, K, N, _ptr(chan), C, _ptr(matrix), _stream()))
    return signals


def generate_signals_grad(env_params, env_shape, env_flags, lo_freq, chain, t_start: float, t_end: float, gsignals,
                          device=None):
    """(grad_env [B,K,E,9], grad_lo [B,K], grad_v2hz [B,K]) from dL/dsignals [B,K,N] (c3b_generate_signals_grad)."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else default_device()
    with torch.cuda.device(device):
        env_params = _as(env_params, torch.float64, device)
        B, K, E, _ = env_params.shape
        env_shape = _as(env_shape, torch.int32, device)
        env_flags = _as(env_flags, torch.int32, device)
        lo_freq = _as(lo_freq, torch.float64, device)
        chain = _as(chain, torch.float64, device)
        gsignals = _as(gsignals, torch.float64, device)
        batched = chain.dim() == 3
        N = gsignals.shape[-1]
        if tuple(gsignals.shape) != (B, K, N):
            raise ValueError("C3:ERROR: gsignals must be [B,K,N]")
        ch = chain.reshape(-1, 11)
        n_awg_max = int((abs(float(t_start) - float(t_end)) * ch[:, 1]).max().item()) + 1
        genv = torch.empty((B, K, E, 9), dtype=torch.float64, device=device)
        glo = torch.empty((B, K), dtype=torch.float64, device=device)
        gv = torch.empty((B, K), dtype=torch.float64, device=device)
        _lib.check(lib.c3b_generate_signals_grad(_ptr(env_params), _ptr(env_shape), _ptr(env_flags), _ptr(lo_freq),
                                                 _ptr(chain), int(batched), float(t_start), float(t_end), B, K, E, N,
                                                 n_awg_max, _ptr(gsignals), _ptr(genv), _ptr(glo), _ptr(gv), _stream()))
    return genv, glo, gv


class _SignalChainFn(torch.autograd.Function):
    """signals = chain(env_params, lo_freq), differentiable w.r.t. both."""

    @staticmethod
    def forward(ctx, env_params, lo_freq, env_shape, env_flags, chain, t_start, t_end):
        sig = generate_signals(env_params.detach(), env_shape, env_flags, lo_freq.detach(), chain, t_start, t_end)
        ctx.save_for_backward(env_params.detach(), lo_freq.detach())
        ctx.static = (env_shape, env_flags, chain, t_start, t_end)
        return sig

    @staticmethod
    def backward(ctx, gsig):
        env_params, lo_freq = ctx.saved_tensors
        env_shape, env_flags, chain, t_start, t_end = ctx.static
        genv, glo, _ = generate_signals_grad(env_params, env_shape, env_flags, lo_freq, chain, t_start, t_end, gsig.contiguous())
        return genv, glo, None, None, None, None, None


def generate_signals_autograd(env_params: torch.Tensor, lo_freq: torch.Tensor, env_shape, env_flags, chain,
                              t_start: float, t_end: float) -> torch.Tensor:
    """Differentiable control fields [B,K,N]: gradients of any loss flow back to the pulse parameters
    ``env_params [B,K,E,9]`` and the carrier frequencies ``lo_freq [B,K]`` (CUDA float64 tensors)."""
    if not (env_params.is_cuda and lo_freq.is_cuda):
        raise ValueError("C3:ERROR: generate_signals_autograd needs CUDA tensors")
    return _SignalChainFn.apply(env_params, lo_freq, env_shape, env_flags, chain, float(t_start), float(t_end))


def frame_dephase(U: torch.Tensor, occ, phases=None, probs=None, lindblad: bool = False) -> torch.Tensor:
    """IN PLACE: U[b] <- dephasing_b . FR_b . U[b] for a batch of propagators ``U [B,D,D]`` (or [D,D]), with the frame
    rotation and dephasing channel of c3/model.py:536-578, 597-639 given by their exponents: ``occ [L,d]`` occupation numbers
    of the driven qubits, ``phases [B,L]`` (freq t_final + framechange) and, for Lindblad superoperators, ``probs [B,L]``.
    Returns U."""
    lib = _lib.load()
    if not (isinstance(U, torch.Tensor) and U.is_cuda and U.dtype == torch.complex128 and U.is_contiguous()):
        raise ValueError("C3:ERROR: frame_dephase needs a contiguous complex128 CUDA tensor")
    device = U.device
    with torch.cuda.device(device):
        Uv = U if U.dim() == 3 else U.unsqueeze(0)
        B, D, _ = Uv.shape
        occ = _as(occ, torch.int32, device)
        if occ.dim() != 2:
            raise ValueError("C3:ERROR: occ must be [L,d]")
        L, d = occ.shape
        if D != (d * d if lindblad else d):
            raise ValueError(f"C3:ERROR: propagators of dimension {D} do not match d = {d} (lindblad = {lindblad})")

        def rows(x, name):
            if x is None:
                return None
            x = _as(x, torch.float64, device)
            if x.dim() == 1:
                x = x.unsqueeze(0).expand(B, L).contiguous()
            if tuple(x.shape) != (B, L):
                raise ValueError(f"C3:ERROR: {name} has shape {tuple(x.shape)}, expected {(B, L)}")
            return x

        phases, probs = rows(phases, "phases"), rows(probs, "probs")
        if probs is not None and bool(((probs < 0) | (probs > 1)).any()):
            raise ValueError("Dephasing channel strength is outside [0,1] range")
        _lib.check(lib.c3b_frame_dephase(_ptr(Uv), B, D, d, _ptr(occ), L, _ptr(phases), _ptr(probs), int(bool(lindblad)), _stream()))
    return U


def dress_models(drift, ops=None, ordered: bool = True, device=None):
    """Dressed frame of every drift Hamiltonian ``drift [B,d,d]`` (or [d,d]) and T^dag X T of ``ops [M,d,d]`` /
    ``[B,M,d,d]`` (c3/model.py:453-534).  Returns dict(eigenframe [B,d], transform [B,d,d], drift [B,d,d],
    ops [B,M,d,d] or None, sweeps [B])."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else default_device()
    with torch.cuda.device(device):
        drift = _as(drift, torch.complex128, device)
        squeeze = drift.dim() == 2
        if squeeze:
            drift = drift.unsqueeze(0)
        B, d, _ = drift.shape
        M, batched = 0, False
        if ops is not None:
            ops = _as(ops, torch.complex128, device)
            batched = ops.dim() == 4
            M = ops.shape[-3]
            if batched and ops.shape[0] != B:
                raise ValueError("C3:ERROR: batched operators must have the batch size of the drift")
        ef = torch.empty((B, d), dtype=torch.float64, device=device)
        T = torch.empty((B, d, d), dtype=torch.complex128, device=device)
        dd = torch.empty((B, d, d), dtype=torch.complex128, device=device)
        dops = torch.empty((B, M, d, d), dtype=torch.complex128, device=device) if M > 0 else None
        info = torch.empty((B,), dtype=torch.int32, device=device)
        _lib.check(lib.c3b_dress_models(_ptr(drift), _ptr(ops), int(batched), B, M, d, int(bool(ordered)), _ptr(ef), _ptr(T),
                                        _ptr(dd), _ptr(dops), _ptr(info), _stream()))
    out = {"eigenframe": ef, "transform": T, "drift": dd, "ops": dops, "sweeps": info}
    if squeeze:
        out = {k: (v[0] if v is not None else None) for k, v in out.items()}
    return out


def kron(A, B, device=None) -> torch.Tensor:
    """(Batched) Kronecker product with the row-major convention of tf_kron."""
    lib = _lib.load()
    device = torch.device(device) if device is not None else default_device()
    with torch.cuda.device(device):
        A = _as(A, torch.complex128, device)
        B = _as(B, torch.complex128, device)
        a_b, b_b = A.dim() > 2, B.dim() > 2
        lead = A.shape[:-2] if a_b else (B.shape[:-2] if b_b else ())
        if a_b and b_b and A.shape[:-2] != B.shape[:-2]:
            raise ValueError("C3:ERROR: kron batch shapes differ")
        batch = 1
        for s in lead:
            batch *= s
        ra, ca = A.shape[-2:]
        rb, cb = B.shape[-2:]
        out = torch.empty(tuple(lead) + (ra * rb, ca * cb), dtype=torch.complex128, device=device)
        _lib.check(lib.c3b_kron(_ptr(A), _ptr(B), _ptr(out), batch, ra, ca, rb, cb, int(a_b), int(b_b), _stream()))
    return out


def measure_fp64_peak(kind: str = "dfma", seconds: float = 0.5, device: Optional[int] = None) -> float:
    """Measured fp64 TFLOP/s of this GPU ("dfma": vector pipe, "dmma": mma.sync m8n8k4)."""
    lib = _lib.load()
    dev = torch.cuda.current_device() if device is None else device
    v = lib.c3b_measure_fp64_peak(0 if kind == "dfma" else 1, dev, float(seconds))
    if v < 0:
        _lib.check(int(v))
    return v


def launch_count() -> int:
    """Kernels launched by libc3b200 in this process so far."""
    return int(_lib.load().c3b_launch_count())


def last_kernel_ms() -> float:
    """Duration of the main fused kernel of the last pwc_* call (needs set_tuning("profile", 1))."""
    v = _lib.load().c3b_last_kernel_ms()
    if v < 0:
        _lib.check(int(v))
    return v
