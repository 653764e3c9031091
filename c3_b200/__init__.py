"""c3_b200: a B200-native engine for the piecewise-constant propagator path of q-optimize/c3
(``Experiment.compute_propagators`` / ``c3.libraries.propagation``).

Importing the package is cheap; the CUDA library (c3_b200/libc3b200.so) is loaded on first use
and its absence is an error -- there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
