"""CPU oracle for gradients through the PWC propagator (TEST INFRASTRUCTURE ONLY, see c3_oracle.py).

The reference differentiates its TensorFlow graph with tf.GradientTape
(c3/optimizers/optimizer.py:210-215): U = prod_n expm(-i (h0 + sum_k c_k[n] h_k) dt), loss = f(U).
TensorFlow is not installable here; torch's CPU autograd through torch.linalg.matrix_exp computes the
same derivative (both differentiate the exact matrix exponential), complex128."""
import numpy as np
import torch


def propagate_torch(h0, hks, signals, dt):
    """signals [B,K,N] float64 tensor (may require grad) -> U [B,d,d] complex128 (CPU)."""
    h0 = torch.as_tensor(h0, dtype=torch.complex128)
    hks = torch.as_tensor(hks, dtype=torch.complex128)
    c = signals.to(torch.complex128)
    H = h0[None, None] + torch.einsum("bkn,kij->bnij", c, hks)
    dU = torch.linalg.matrix_exp(-1j * dt * H)
    B, N = dU.shape[:2]
    U = dU[:, 0]
    for n in range(1, N):
        U = dU[:, n] @ U
    return U


def loss_and_grad(h0, hks, signals_np, dt, target):
    """L = sum_b (1 - |tr(T_b^dag U_b)|^2 / d^2)  (a unitary-overlap infidelity) and dL/dsignals."""
    sig = torch.tensor(np.asarray(signals_np), dtype=torch.float64, requires_grad=True)
    U = propagate_torch(h0, hks, sig, dt)
    T = torch.as_tensor(target, dtype=torch.complex128)
    d = U.shape[-1]
    ov = torch.einsum("bij,bij->b", T.conj(), U)
    L = (1.0 - (ov.abs() ** 2) / d ** 2).sum()
    L.backward()
    return float(L), sig.grad.numpy(), U.detach().numpy()


def propagate_lindblad_torch(h0, hks, col_ops, signals, dt):
    """Lindblad superoperator propagators [B,D,D] on torch-CPU (autograd-able in ``signals``): the operator of
    c3/libraries/propagation.py:563-582 (row-major Kronecker convention of tf_kron), matrix_exp per slice,
    later slices on the left."""
    h0 = torch.as_tensor(h0, dtype=torch.complex128)
    hks = torch.as_tensor(hks, dtype=torch.complex128)
    cols = torch.as_tensor(np.asarray(col_ops), dtype=torch.complex128)
    d = h0.shape[-1]
    eye = torch.eye(d, dtype=torch.complex128)
    c = signals.to(torch.complex128)
    H = h0[None, None] + torch.einsum("bkn,kij->bnij", c, hks)

    def kron(a, b):
        return torch.einsum("...ij,...kl->...ikjl", a, b).reshape(a.shape[:-2] + (d * d, d * d))

    lind = -1j * (kron(H, eye.expand_as(H)) - kron(eye.expand_as(H), H.transpose(-1, -2)))
    diss = torch.zeros((d * d, d * d), dtype=torch.complex128)
    for L in cols:
        m = L.conj().T @ L
        diss = diss + kron(L, L.conj()) - 0.5 * kron(m, eye) - 0.5 * kron(eye, m.T)
    dU = torch.linalg.matrix_exp((lind + diss) * dt)
    U = dU[:, 0]
    for n in range(1, dU.shape[1]):
        U = dU[:, n] @ U
    return U
