"""CPU check of the polynomial evaluation schemes the CUDA kernels use for the matrix exponential (c3b_common.cuh): the
literals in the header, composed with exact rational arithmetic, reproduce the Taylor coefficients 1/k! -- degree 18 for the
5-product scheme (Bader-Blanes-Casas T18), degree 15 for the 4-product scheme -- and both agree with scipy's expm on matrices
in their norm range."""
import math
import os
import re
from fractions import Fraction

import numpy as np
import scipy.linalg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _literals(prefix):
    src = open(os.path.join(ROOT, "c3_b200", "csrc", "c3b_common.cuh")).read()
    out = {}
    for name, val in re.findall(r"#define\s+(%s\w+)\s+\(?(-?[0-9.]+(?:e-?[0-9]+)?)\)?" % prefix, src):
        out[name[len(prefix):]] = Fraction(val)
    return out


def _mul(p, q):
    r = [Fraction(0)] * (len(p) + len(q) - 1)
    for i, a in enumerate(p):
        for j, b in enumerate(q):
            r[i + j] += a * b
    return r


def _add(*ps):
    r = [Fraction(0)] * max(len(p) for p in ps)
    for p in ps:
        for i, a in enumerate(p):
            r[i] += a
    return r


def _sc(c, p):
    return [c * a for a in p]


X = [Fraction(0), Fraction(1)]
ONE = [Fraction(1)]


def test_t15_scheme_matches_taylor_to_degree_15():
    c = _literals("C3B_T15_")
    x2 = _mul(X, X)
    p0 = _mul(x2, _add(_sc(c["A1"], x2), _sc(c["A2"], X)))
    p1 = _add(_mul(_add(p0, _sc(c["B1"], x2), _sc(c["B2"], X)), _add(p0, _sc(c["B3"], x2), _sc(c["B4"], ONE))), _sc(c["B5"], p0))
    t = _add(_mul(_add(p1, _sc(c["C1"], x2), _sc(c["C2"], X)), _add(p1, _sc(c["C3"], p0), _sc(c["C4"], X))),
             _sc(c["C9"], p1), _sc(c["C5"], p0), _sc(c["C6"], x2), _sc(c["C7"], X), _sc(c["C8"], ONE))
    assert len(t) == 17
    for k in range(16):
        assert abs(float(t[k] * math.factorial(k) - 1)) < 1e-17, (k, float(t[k] * math.factorial(k)))
    assert abs(float(t[16] * math.factorial(16)) - 0.5457435022) < 1e-9          # the "+": degree 16 is not the Taylor term


def test_t18_scheme_matches_taylor_to_degree_18():
    c = _literals("C3B_T18_")
    a2 = _mul(X, X)
    a3 = _mul(a2, X)
    a6 = _mul(a3, a3)
    b1 = _add(_sc(c["A11"], X), _sc(c["A21"], a2), _sc(c["A31"], a3))
    b2 = _add(_sc(c["B11"], X), _sc(c["B21"], a2), _sc(c["B31"], a3), _sc(c["B61"], a6))
    b3 = _add(_sc(c["B02"], ONE), _sc(c["B12"], X), _sc(c["B22"], a2), _sc(c["B32"], a3), _sc(c["B62"], a6))
    b4 = _add(_sc(c["B03"], ONE), _sc(c["B13"], X), _sc(c["B23"], a2), _sc(c["B33"], a3), _sc(c["B63"], a6))
    b5 = _add(_sc(c["B24"], a2), _sc(c["B34"], a3), _sc(c["B64"], a6))
    a9 = _add(_mul(b1, b5), b4)
    t = _add(b2, _mul(_add(b3, a9), a9))
    for k in range(19):
        assert abs(float(t[k] * math.factorial(k) - 1)) < 1e-15, (k, float(t[k] * math.factorial(k)))


def _t15(A):
    c = {k: float(v) for k, v in _literals("C3B_T15_").items()}
    I = np.eye(A.shape[0])
    A2 = A @ A
    p0 = A2 @ (c["A1"] * A2 + c["A2"] * A)
    p1 = (p0 + c["B1"] * A2 + c["B2"] * A) @ (p0 + c["B3"] * A2 + c["B4"] * I) + c["B5"] * p0
    return (p1 + c["C1"] * A2 + c["C2"] * A) @ (p1 + c["C3"] * p0 + c["C4"] * A) + c["C9"] * p1 + c["C5"] * p0 + c["C6"] * A2 \
        + c["C7"] * A + c["C8"] * I


def test_t15_accuracy_within_theta():
    """Forward error against scipy for anti-Hermitian and general matrices with inf-norm up to theta_15 = 0.8 (the kernels scale
    and square above it): a few 1e-15 at most, i.e. 1e-12 after the 1000 products of a gate -- two orders inside the 1e-10
    tolerance of the path."""
    rng = np.random.default_rng(0)
    worst = 0.0
    for trial in range(60):
        d = int(rng.integers(2, 30))
        H = rng.normal(size=(d, d)) + 1j * rng.normal(size=(d, d))
        A = -1j * (H + H.conj().T) if trial % 2 == 0 else H
        if trial % 3 == 0:      # physical shape: large diagonal, weak couplings
            A = 1j * np.diag(rng.uniform(-1, 1, d)) + 0.02 * A
        A = A * (rng.uniform(0.05, 0.8) / np.abs(A).sum(axis=1).max())
        E = scipy.linalg.expm(A)
        worst = max(worst, np.linalg.norm(_t15(A) - E) / np.linalg.norm(E))
    assert worst < 5e-15, worst
