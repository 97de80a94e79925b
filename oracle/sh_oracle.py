"""TEST INFRASTRUCTURE ONLY -- CPU (numpy, float64) restatement of the reference SH encoder.

The reference hard-codes the 64 real spherical-harmonics polynomials of degree < 8 and their partial derivatives
(shencoder/src/shencoder.cu:43-121 values, :130-350 d/dx, d/dy, d/dz; coefficient index i = l*(l+1)+m, Condon-Shortley
phase, e.g. outputs[1] = -0.4886 y, outputs[2] = 0.4886 z, outputs[3] = -0.4886 x).  Those polynomials are

    Y_lm = k_lm * {Re, Im}((x + i y)^|m|) * d^|m| P_l(z) / dz^|m|,

evaluated here from numpy's Legendre polynomials and complex powers, independently of the generator that writes
the CUDA basis (raw_ngp_b200/csrc/gen_sh_basis.py).  Pinned by tests/golden/sh_deg8.npz (outputs of the
reference's own kernel) and tests/golden/sh_ref_source.npz (the reference's source expressions evaluated verbatim
by tools/eval_reference_sh_source.py).
"""
from math import factorial, pi, sqrt

import numpy as np
from numpy.polynomial import legendre as npleg
from numpy.polynomial import polynomial as nppoly


def _q(l, a):
    """power-series coefficients of d^a P_l / dz^a"""
    c = npleg.leg2poly([0] * l + [1])
    return nppoly.polyder(c, a) if a > 0 else c


def sh_basis(dirs, degree, jacobian=False):
    """dirs [B,3] -> Y [B, degree^2] (float64); with jacobian=True also dY [B, 3, degree^2] (reference dy_dx layout)."""
    d = np.asarray(dirs, dtype=np.float64)
    x, y, z = d[:, 0], d[:, 1], d[:, 2]
    w = x + 1j * y
    B = d.shape[0]
    n = degree * degree
    Y = np.zeros((B, n))
    J = np.zeros((B, 3, n)) if jacobian else None
    for l in range(degree):
        for m in range(-l, l + 1):
            a = abs(m)
            K = sqrt((2 * l + 1) / (4 * pi) * factorial(l - a) / factorial(l + a))
            k = K if m == 0 else (-1) ** a * sqrt(2.0) * K
            q = nppoly.polyval(z, _q(l, a))
            wa = w ** a
            A = wa.real if m >= 0 else wa.imag
            i = l * (l + 1) + m
            Y[:, i] = k * A * q
            if jacobian:
                dq = nppoly.polyval(z, _q(l, a + 1)) if a + 1 <= l else np.zeros_like(z)
                if a == 0:
                    dAx = dAy = np.zeros_like(z)
                else:
                    wp = a * w ** (a - 1)          # d(w^a)/dx = a w^(a-1) ; d(w^a)/dy = i a w^(a-1)
                    dAx = wp.real if m >= 0 else wp.imag
                    dAy = (1j * wp).real if m >= 0 else (1j * wp).imag
                J[:, 0, i] = k * dAx * q
                J[:, 1, i] = k * dAy * q
                J[:, 2, i] = k * A * dq
    return (Y, J) if jacobian else Y


def sh_backward(grad, dirs, degree):
    """grad_inputs[b,d] = sum_ch grad[b,ch] * dY[b,d,ch]   (shencoder.cu:358-382)."""
    _, J = sh_basis(dirs, degree, jacobian=True)
    return np.einsum("bc,bdc->bd", np.asarray(grad, dtype=np.float64), J)
