"""TEST INFRASTRUCTURE ONLY -- CPU restatement of `compute_invN_lm`
(commander3/src/comm_N_mod.f90:127-197): the diagonal N^-1_{lm,lm} of the inverse noise covariance in
harmonic space, as a sum of Wigner-3j products over the m=0 coefficients of the inverse-noise map.

The reference calls SLATEC's DRC3JJ (a third-party routine that is not under /root/reference) twice per (l,m):
    DRC3JJ(l, l, 0, 0, ...)   -> (L l l; 0 0 0)    for all allowed L   (:159-160)
    DRC3JJ(l, l, -m, m, ...)  -> (L l l; 0 -m m)   for all allowed L   (:161-162)
and accumulates  a_L0 sqrt(2L+1) (L l l;0 -m m)(L l l;0 0 0)  for L <= min(2l, lmax)  (:165-171), then scales by
(2l+1)/sqrt(4pi) * npix/(4pi) and flips the sign for odd m (:172-173).  DRC3JJ evaluates the 3j symbols by the
Schulten-Gordon recursion; this restatement evaluates the same symbols from the Racah formula in exact integer
arithmetic (slow, any l), which pins them independently of any recursion.  Parity: unpinned by the reference (no
test or fixture holds N_lm values); pinned here against sympy's wigner_3j, the closed forms of tests/test_oracle.py
and a direct numerical quadrature of |Y_lm|^2 Y_L0.

Only tests/ may import this module; the product computes N_lm on the GPU (commander_b200/csrc/invn.cu).
"""
from __future__ import annotations

from fractions import Fraction
from functools import lru_cache
from math import factorial, isqrt, pi, sqrt

import numpy as np


def _sqrt_fraction(q: Fraction) -> float:
    """sqrt of a non-negative rational to double precision without overflow: scale to a big integer square root."""
    if q == 0:
        return 0.0
    shift = 240   # bits of extra precision
    num = q.numerator << (2 * shift)
    r = isqrt(num // q.denominator)
    return r / float(1 << shift) if r.bit_length() < 1000 else float(Fraction(r, 1 << shift))


@lru_cache(maxsize=None)
def wigner3j(j1: int, j2: int, j3: int, m1: int, m2: int, m3: int) -> float:
    """(j1 j2 j3; m1 m2 m3) for integer arguments, Racah's formula in exact arithmetic."""
    if m1 + m2 + m3 != 0 or j3 < abs(j1 - j2) or j3 > j1 + j2:
        return 0.0
    if abs(m1) > j1 or abs(m2) > j2 or abs(m3) > j3:
        return 0.0
    f = factorial
    delta = Fraction(f(j1 + j2 - j3) * f(j1 - j2 + j3) * f(-j1 + j2 + j3), f(j1 + j2 + j3 + 1))
    pref = delta * f(j1 + m1) * f(j1 - m1) * f(j2 + m2) * f(j2 - m2) * f(j3 + m3) * f(j3 - m3)
    tmin = max(0, j2 - j3 - m1, j1 - j3 + m2)
    tmax = min(j1 + j2 - j3, j1 - m1, j2 + m2)
    s = Fraction(0)
    for t in range(tmin, tmax + 1):
        den = f(t) * f(j3 - j2 + t + m1) * f(j3 - j1 + t - m2) * f(j1 + j2 - j3 - t) * f(j1 - t - m1) * f(j2 - t + m2)
        s += Fraction((-1) ** t, den)
    sign = -1 if (j1 - j2 - m3) & 1 else 1
    # value = sign * sqrt(pref) * s ; keep s exact and take one square root
    val2 = pref * s * s
    return sign * (1 if s > 0 else -1) * _sqrt_fraction(val2)


def compute_invN_lm(a_l0: np.ndarray, lmax: int, ms, npix: float) -> np.ndarray:
    """commander3/src/comm_N_mod.f90:150-184.  a_l0 (nmaps, lmax+1); returns N_lm (nmaps, nalm) in the local
    real-packed order of the m's in `ms` (m=0: l=0..lmax; m>0: the (+m, -m) pair for l=m..lmax, both equal)."""
    a_l0 = np.asarray(a_l0, dtype=np.float64)
    nmaps = a_l0.shape[0]
    cols = []
    for m in ms:
        for l in range(m, lmax + 1):
            val = np.zeros(nmaps)
            for lp in range(0, min(2 * l, lmax) + 1):          # l1min = 0, l1max = 2l ; exit above lmax (:167)
                t = wigner3j(lp, l, l, 0, -m, m) * wigner3j(lp, l, l, 0, 0, 0)
                if t != 0.0:
                    val += a_l0[:, lp] * sqrt(2.0 * lp + 1.0) * t
            val *= (2 * l + 1) / sqrt(4.0 * pi) * npix / (4.0 * pi)
            if m & 1:
                val = -val
            cols.append(val)
            if m > 0:
                cols.append(val)
    return np.array(cols).T.copy() if cols else np.zeros((nmaps, 0))


def gaunt_quadrature(l: int, m: int, L: int, n: int = 400) -> float:
    """int |Y_lm|^2 Y_L0 dOmega by Gauss-Legendre quadrature of scipy's spherical harmonics (pins the 3j product)."""
    from scipy.special import sph_harm_y
    x, w = np.polynomial.legendre.leggauss(n)
    th = np.arccos(x)
    ylm = sph_harm_y(l, m, th, 0.0)
    yL0 = sph_harm_y(L, 0, th, 0.0).real
    return float(2.0 * pi * np.sum(w * (ylm.real ** 2 + ylm.imag ** 2) * yL0))
