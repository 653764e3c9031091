"""Mirror of the hot-path helpers of ``c3.utils.tf_utils`` on the B200 engine:

  tf_matmul_left(dUs)                       c3/utils/tf_utils.py:120-129
  tf_matmul_right(dUs)                      :132-141
  tf_matmul_n(tensor_list, folding_stack)   :144-163   (+ _tf_matmul_n_even/_odd :166-193)
  Id_like, tf_kron, tf_spre, tf_spost, tf_super        :240-289

The ordered products run as one device kernel; the reference's pairwise tree and its
sequential fold differ only by fp64 re-association ("hard parts").
"""
from __future__ import annotations

from typing import Callable, List

import numpy as np
import torch

from . import engine


def _t(x) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (list, tuple)) and len(x) and isinstance(x[0], torch.Tensor):
        return torch.stack(list(x))
    if hasattr(x, "numpy") and not isinstance(x, np.ndarray):
        x = x.numpy()
    return torch.as_tensor(np.asarray(x))


def tf_matmul_left(dUs) -> torch.Tensor:
    """dU_{N-1} ... dU_1 dU_0 for dUs [N,n,n] (later matrices multiply from the left)."""
    return engine.ordered_product(_t(dUs))


def tf_matmul_right(dUs) -> torch.Tensor:
    """dU_0 dU_1 ... dU_{N-1}: the left product of the reversed list."""
    return engine.ordered_product(torch.flip(_t(dUs), dims=[-3]))


def _tf_matmul_n_even(odd, even):
    """Marker for one tree level with equal halves (c3/utils/tf_utils.py:166-178)."""
    return torch.matmul(odd, even)


def _tf_matmul_n_odd(odd, even):
    """Marker for one tree level that carries the unpaired last element (:181-193)."""
    return torch.cat([torch.matmul(odd, even[:-1]), even[-1:]], 0)


def compute_folding_stack(n_steps: int) -> List[Callable]:
    """The per-level even/odd list of Experiment._compute_folding_stack (c3/experiment.py:93-107)."""
    stack = []
    n = int(n_steps)
    while n > 1:
        stack.append(_tf_matmul_n_even if n % 2 == 0 else _tf_matmul_n_odd)
        n = int(np.ceil(n / 2))
    return stack


def tf_matmul_n(tensor_list, folding_stack=None) -> torch.Tensor:
    """Ordered product of ``tensor_list`` [N,n,n].  ``folding_stack`` is accepted for call
    compatibility and ignored: the fold happens in one kernel instead of ceil(log2 N)
    batched-matmul levels."""
    del folding_stack
    return engine.ordered_product(_t(tensor_list))


def Id_like(A) -> torch.Tensor:
    """Identity of the same (batched) shape as A."""
    A = _t(A)
    eye = torch.eye(A.shape[-1], dtype=A.dtype, device=A.device)
    return eye.expand(A.shape).clone()


def tf_kron(A, B) -> torch.Tensor:
    """Kronecker product of 2 matrices, with optional leading batch dimensions."""
    return engine.kron(_t(A), _t(B))


def tf_spre(A) -> torch.Tensor:
    """Superoperator on the left of matrix A:  A (x) I."""
    A = _t(A)
    return engine.kron(A, torch.eye(A.shape[-1], dtype=torch.complex128))


def tf_spost(A) -> torch.Tensor:
    """Superoperator on the right of matrix A:  I (x) A^T."""
    A = _t(A)
    return engine.kron(torch.eye(A.shape[-1], dtype=torch.complex128), A.transpose(-1, -2).contiguous())


def tf_super(A) -> torch.Tensor:
    """Superoperator from both sides of matrix A: spre(A) @ spost(A^dagger) = A (x) A^*."""
    A = _t(A).to(torch.complex128)
    return engine.kron(A, A.conj().resolve_conj())
