#!/usr/bin/env python
"""Derivation and float32 check of the polynomial the MLP epilogues use for Phi(x) = (1 + erf(x / sqrt 2)) / 2
(slak_b200/csrc/mlp_tc.cu: phi2): odd Chebyshev fit of Phi(x) - 0.5 on [-L, L], evaluated in fp32 Horner form as
sat(0.5 + x Q(x^2)); prints the coefficients and the maximum error over [-12, 12]."""
import numpy as np
from numpy.polynomial import chebyshev as Ch
from scipy.special import erf

L, N = 4.25, 9
Phi = lambda x: 0.5 * (1 + erf(x / np.sqrt(2)))
xs = np.cos(np.linspace(0, np.pi, 8001)) * L
c = Ch.chebfit(xs / L, Phi(xs) - 0.5, 2 * N - 1)
c[0::2] = 0
p = Ch.cheb2poly(c)
coef = [np.float32(p[2 * k + 1] / L ** (2 * k + 1)) for k in range(N)]
print("coefficients of x^(2k+1), k = 0..%d:" % (N - 1), ", ".join("%.9ef" % v for v in coef))
x = np.linspace(-12, 12, 2400001).astype(np.float32)
u = (x * x).astype(np.float32)
q = np.full_like(u, coef[-1])
for ck in coef[-2::-1]:
    q = (q * u + ck).astype(np.float32)
val = np.clip(np.float32(0.5) + x * q, 0, 1)
err = np.abs(val - Phi(x.astype(np.float64)))
print("max |Phi_poly - Phi| over [-12, 12] in float32: %.3e at x = %.3f" % (err.max(), x[err.argmax()]))
g = np.abs(x * val - x * Phi(x.astype(np.float64)))
print("max |gelu_poly - gelu|: %.3e" % g.max())
