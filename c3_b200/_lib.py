"""ctypes binding of libc3b200.so (include/c3b200.h).  There is NO fallback: if the CUDA
library is missing, importing the engine raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libc3b200.so")

#: every symbol include/c3b200.h declares (checked by tests/test_cabi.py)
SYMBOLS = [
    "c3b_version", "c3b_last_error", "c3b_pwc_workspace_bytes", "c3b_pwc_closed", "c3b_pwc_closed_hlist",
    "c3b_pwc_lindblad", "c3b_product_workspace_bytes", "c3b_ordered_product", "c3b_seq_product", "c3b_kron",
    "c3b_set_tuning", "c3b_pwc_path", "c3b_measure_fp64_peak", "c3b_launch_count",
    "c3b_last_kernel_ms", "c3b_pwc_grad_workspace_bytes", "c3b_pwc_closed_grad",
    "c3b_gate_infid", "c3b_gate_infid_grad", "c3b_seq_populations", "c3b_signal_slice_num", "c3b_generate_signals",
    "c3b_generate_signals_grad", "c3b_pwc_lindblad_grad_workspace_bytes", "c3b_pwc_lindblad_grad",
    "c3b_dress_models", "c3b_pwc_closed_gated", "c3b_pwc_gated_supported",
    "c3b_generate_signals_noisy", "c3b_generate_signals_table", "c3b_crosstalk", "c3b_pwc_closed_saved_bytes", "c3b_pwc_closed_saved_chunks",
    "c3b_pwc_closed_fwd_saved", "c3b_pwc_closed_bwd_saved", "c3b_frame_dephase", "c3b_model_bytes", "c3b_model_prepare", "c3b_pwc_prepared_workspace_bytes", "c3b_pwc_prepared",
]

_lib = None


class C3BError(Exception):
    """Raised when a C-ABI call returns a non-zero status; message starts with 'C3:ERROR:'
    like the reference's own exceptions (c3/experiment.py:464-468)."""


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"C3:ERROR: {LIB_PATH} is missing. Build it with `python -m c3_b200.build` "
            "(needs nvcc; there is no CPU fallback for the propagator path)."
        )
    lib = C.CDLL(LIB_PATH)
    vp, sz, i, d = C.c_void_p, C.c_size_t, C.c_int, C.c_double
    lib.c3b_version.restype = i
    lib.c3b_last_error.restype = C.c_char_p
    lib.c3b_pwc_workspace_bytes.restype = sz
    lib.c3b_pwc_workspace_bytes.argtypes = [i, i, i, i, i, i]
    lib.c3b_pwc_closed.restype = i
    lib.c3b_pwc_closed.argtypes = [vp, vp, vp, d, i, i, i, i, i, vp, vp, vp, sz, vp]
    lib.c3b_pwc_closed_gated.restype = i
    lib.c3b_pwc_closed_gated.argtypes = [vp, vp, vp, d, i, i, i, i, vp, vp, vp, sz, vp]
    lib.c3b_pwc_gated_supported.restype = i
    lib.c3b_pwc_gated_supported.argtypes = [i]
    lib.c3b_pwc_closed_hlist.restype = i
    lib.c3b_pwc_closed_hlist.argtypes = [vp, d, i, i, i, vp, vp, vp, sz, vp]
    lib.c3b_pwc_lindblad.restype = i
    lib.c3b_pwc_lindblad.argtypes = [vp, vp, vp, i, vp, d, i, i, i, i, i, vp, vp, vp, sz, vp]
    lib.c3b_pwc_grad_workspace_bytes.restype = sz
    lib.c3b_pwc_grad_workspace_bytes.argtypes = [i, i, i, i, i]
    lib.c3b_pwc_closed_grad.restype = i
    lib.c3b_pwc_closed_grad.argtypes = [vp, vp, vp, d, i, i, i, i, vp, vp, vp, i, vp, sz, vp]
    lib.c3b_pwc_lindblad_grad_workspace_bytes.restype = sz
    lib.c3b_pwc_lindblad_grad_workspace_bytes.argtypes = [i, i, i, i, i]
    lib.c3b_pwc_lindblad_grad.restype = i
    lib.c3b_pwc_lindblad_grad.argtypes = [vp, vp, vp, i, vp, d, i, i, i, i, vp, vp, vp, i, vp, sz, vp]
    lib.c3b_frame_dephase.restype = i
    lib.c3b_frame_dephase.argtypes = [vp, i, i, i, vp, i, vp, vp, i, vp]
    lib.c3b_dress_models.restype = i
    lib.c3b_dress_models.argtypes = [vp, vp, i, i, i, i, i, vp, vp, vp, vp, vp, vp]
    lib.c3b_gate_infid.restype = i
    lib.c3b_gate_infid.argtypes = [vp, i, i, vp, vp, i, i, vp, vp, vp]
    lib.c3b_gate_infid_grad.restype = i
    lib.c3b_gate_infid_grad.argtypes = [vp, vp, vp, vp, i, i, i, i, vp, vp]
    lib.c3b_seq_populations.restype = i
    lib.c3b_seq_populations.argtypes = [vp, i, vp, vp, i, i, i, vp, i, vp, vp, vp]
    lib.c3b_signal_slice_num.restype = i
    lib.c3b_signal_slice_num.argtypes = [d, d, d]
    lib.c3b_generate_signals.restype = i
    lib.c3b_generate_signals.argtypes = [vp, vp, vp, vp, vp, i, d, d, i, i, i, i, vp, vp]
    lib.c3b_generate_signals_noisy.restype = i
    lib.c3b_generate_signals_noisy.argtypes = [vp, vp, vp, vp, vp, i, d, d, i, i, i, i, vp, i, C.c_ulonglong, vp, vp, vp]
    lib.c3b_pwc_closed_saved_bytes.restype = C.c_size_t
    lib.c3b_pwc_closed_saved_bytes.argtypes = [i, i, i, i]
    lib.c3b_pwc_closed_saved_chunks.restype = i
    lib.c3b_pwc_closed_saved_chunks.argtypes = [i, i, i, i, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.c3b_pwc_closed_fwd_saved.restype = i
    lib.c3b_pwc_closed_fwd_saved.argtypes = [vp, vp, vp, d, i, i, i, i, vp, vp, C.c_size_t, vp]
    lib.c3b_pwc_closed_bwd_saved.restype = i
    lib.c3b_pwc_closed_bwd_saved.argtypes = [vp, i, i, i, i, i, i, vp, vp, vp, C.c_size_t, vp]
    lib.c3b_crosstalk.restype = i
    lib.c3b_crosstalk.argtypes = [vp, i, i, i, vp, i, vp, vp]
    lib.c3b_generate_signals_table.restype = i
    lib.c3b_generate_signals_table.argtypes = [vp, vp, vp, vp, i, vp, vp, i, d, d, i, i, i, i, vp, i, C.c_ulonglong, vp, vp, vp]
    lib.c3b_generate_signals_grad.restype = i
    lib.c3b_generate_signals_grad.argtypes = [vp, vp, vp, vp, vp, i, d, d, i, i, i, i, i, vp, vp, vp, vp, vp]
    lib.c3b_product_workspace_bytes.restype = sz
    lib.c3b_product_workspace_bytes.argtypes = [i, i, i]
    lib.c3b_ordered_product.restype = i
    lib.c3b_ordered_product.argtypes = [vp, i, i, i, vp, vp, sz, vp]
    lib.c3b_seq_product.restype = i
    lib.c3b_seq_product.argtypes = [vp, i, vp, vp, i, i, i, vp, vp, sz, vp]
    lib.c3b_kron.restype = i
    lib.c3b_kron.argtypes = [vp, vp, vp, i, i, i, i, i, i, i, vp]
    lib.c3b_set_tuning.restype = i
    lib.c3b_set_tuning.argtypes = [C.c_char_p, C.c_longlong]
    lib.c3b_pwc_path.restype = i
    lib.c3b_pwc_path.argtypes = [i, i, i]
    lib.c3b_measure_fp64_peak.restype = d
    lib.c3b_measure_fp64_peak.argtypes = [i, i, d]
    lib.c3b_launch_count.restype = C.c_longlong
    lib.c3b_last_kernel_ms.restype = d
    lib.c3b_model_bytes.restype = sz
    lib.c3b_model_bytes.argtypes = [i, i, i, i]
    lib.c3b_model_prepare.restype = i
    lib.c3b_model_prepare.argtypes = [vp, vp, vp, i, d, i, i, i, i, vp, sz, vp]
    lib.c3b_pwc_prepared_workspace_bytes.restype = sz
    lib.c3b_pwc_prepared_workspace_bytes.argtypes = [i, i, i, i, i]
    lib.c3b_pwc_prepared.restype = i
    lib.c3b_pwc_prepared.argtypes = [vp, vp, i, i, i, i, i, i, vp, vp, vp, sz, vp]
    _lib = lib
    return lib


C3B_EUNSUPPORTED = -4      # include/c3b200.h


def check(rc: int) -> None:
    if rc != 0:
        msg = load().c3b_last_error().decode("utf-8", "replace")
        raise C3BError(msg or f"C3:ERROR: libc3b200 call failed with status {rc}")
