"""CPU restatement of comm_conviqt's convolution cube -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; nothing under
commander_b200/ does.  Follows commander3/src/comm_conviqt_mod.f90 line by line (plain Python loops: small
cases only):

    beam_table       constructor, :94-115      beam a_lm -> single-precision complex alm_beam(nmaps, ntri)
    get_alms         :294-357                  sky x beam products -> the a_lm columns of beam index m_b
    precompute_sky   :207-292                  spin-m_b syntheses + one FFTW c2r (length 2*bmax) per pixel
    interp           :155-205                  psi lookup / linear interpolation, single precision

The spin-j synthesis itself is oracle/sht_cpu.c (the restatement of what libsharp2 does for
`sharp_execute(SHARP_Y, j, 2, ...)`, :254).  Parity unpinned: the reference has no test or golden vector for
this routine; what pins this file is numpy's irfft (the same c2r definition as FFTW's: backward sign,
unnormalised, imaginary parts of the DC and Nyquist inputs ignored) and the closed form
v1 + conj(v2) mfac = 2 sum_c s_c Re(b_c), v1 - conj(v2) mfac = 2i sum_c s_c Im(b_c) (tests/test_conviqt.py).
"""
from __future__ import annotations

import math

import numpy as np


def lm_table(lmax, ms=None):
    """(l, m) of every local real-packed entry, commander3/src/comm_map_mod.f90:242-261."""
    ms = range(lmax + 1) if ms is None else ms
    lm = []
    for m in ms:
        for l in range(m, lmax + 1):
            if m == 0:
                lm.append((l, 0))
            else:
                lm.append((l, m))
                lm.append((l, -m))
    return lm


def lm2i(lm):
    return {t: i for i, t in enumerate(lm)}


def beam_table(lmax, nmaps, lm, beam_alm, beam_lmax=None):
    """alm_beam of the constructor, :94-115; returned as (ntri, nmaps) complex64 (0-based triangular index
    l(l+1)/2 + m).  beam_alm (nmaps, nalm) real-packed on the layout `lm`; entries with l > beam_lmax stay 0."""
    ntri = (lmax + 1) * (lmax + 2) // 2
    out = np.zeros((ntri, nmaps), dtype=np.complex64)
    idx = lm2i(lm)
    s2 = np.float32(math.sqrt(np.float32(2.0)))          # sqrt(2.0) in single precision
    for (l, m), k in idx.items():
        if m < 0:
            continue
        if beam_lmax is not None and l > beam_lmax:
            continue
        j = l * (l + 1) // 2 + m
        for c in range(nmaps):
            if m == 0:
                out[j, c] = np.complex64(complex(np.float32(beam_alm[c, k]), 0.0))
            else:
                re = np.float32(beam_alm[c, k]) / s2
                im = np.float32(beam_alm[c, idx[(l, -m)]]) / s2
                out[j, c] = np.complex64(complex(re, im))
    return out


def get_alms(m_b, lmax, lm, sky_alm, alm_beam):
    """:294-357.  Returns alm (2, nalm) float64 (column 2 stays 0 for m_b = 0)."""
    nmaps = sky_alm.shape[0]
    idx = lm2i(lm)
    spinsign = -1.0 if m_b != 0 else 1.0
    sqrt_two = math.sqrt(2.0)
    alm = np.zeros((2, len(lm)))
    for i, (l, m) in enumerate(lm):
        if m < 0:
            continue
        if l < m_b:
            continue
        mfac = -1.0 if (m & 1) else 1.0
        lnorm = 0.5 * math.sqrt(4.0 * math.pi / (2.0 * l + 1.0))
        j = l * (l + 1) // 2 + m_b
        alm_b = [complex(alm_beam[j, c]) for c in range(nmaps)]
        if m == 0:                                        # get_alm_TEB, comm_map_mod.f90:1523-1546
            alm_s = [complex(sky_alm[c, i], 0.0) for c in range(nmaps)]
        else:
            ineg = idx[(l, -m)]
            alm_s = [1.0 / math.sqrt(2.0) * complex(sky_alm[c, i], sky_alm[c, ineg]) for c in range(nmaps)]
        v1 = sum(s * b for s, b in zip(alm_s, alm_b))
        v2 = sum(s.conjugate() * b for s, b in zip(alm_s, alm_b)) * mfac
        almc = spinsign * lnorm * (v1 + v2.conjugate() * mfac)
        if m == 0:
            alm[0, i] = almc.real
        else:
            alm[0, i] = almc.real * sqrt_two
            alm[0, ineg] = almc.imag * sqrt_two
        if m_b > 0:
            almc = -1j * spinsign * lnorm * (v1 - v2.conjugate() * mfac)
            if m == 0:
                alm[1, i] = almc.real
            else:
                alm[1, i] = almc.real * sqrt_two
                alm[1, ineg] = almc.imag * sqrt_two
    return alm


def c2r(dv, n):
    """FFTW's c2r for one pixel as a plain sum: dt_k = sum over the Hermitian-completed spectrum of
    X_j e^{+2 pi i j k / n}; the imaginary parts of X_0 and X_{n/2} do not enter."""
    h = n // 2
    dt = np.empty(n)
    for k in range(n):
        acc = dv[0].real + dv[h].real * (-1.0) ** k
        for j in range(1, h):
            w = 2.0 * math.pi * ((j * k) % n) / n
            acc += 2.0 * (dv[j].real * math.cos(w) - dv[j].imag * math.sin(w))
        dt[k] = acc
    return dt


def precompute_sky(S, nside, lmax, bmax, sky_alm, alm_beam, rings=None, ms=None, vectorised=True):
    """:207-292 with the spin-j synthesis done by oracle/sht_cpu (S).  Returns the cube in double precision,
    shape (2*bmax, np): c(pix, psi) of the reference before its real(., sp) rounding (:281)."""
    lm = lm_table(lmax, ms)
    npix = S.map_size(nside, rings)
    marray = {}
    for j in range(bmax + 1):
        alm = get_alms(j, lmax, lm, sky_alm, alm_beam)
        if j == 0:
            marray[0] = S.execute(S.Y, 0, nside, lmax, alm=alm[0:1], rings=rings, ms=ms)[0].copy()
        else:
            mout = S.execute(S.Y, j, nside, lmax, alm=alm, rings=rings, ms=ms)
            marray[j] = mout[0].copy()
            marray[-j] = mout[1].copy()
    n = 2 * bmax
    dv = np.empty((npix, bmax + 1), dtype=np.complex128)
    dv[:, 0] = marray[0]
    for j in range(1, bmax + 1):
        dv[:, j] = marray[j] + 1j * marray[-j]
    if vectorised:
        cube = np.fft.irfft(dv, n=n, axis=1) * n
    else:
        cube = np.stack([c2r(dv[i], n) for i in range(npix)])
    return np.ascontiguousarray(cube.T)


def interp(c, psisteps, pixnum, psi, optim=0):
    """:155-205.  c is the cube (psisteps, npix) in single precision; arithmetic in single precision as in
    the reference (psires, unwrap, x0, x1 are real(sp))."""
    f = np.float32
    psires = f(2.0 * math.pi / psisteps)
    twopi = f(2.0 * math.pi)
    unwrap = f(np.fmod(f(-f(psi)), twopi))
    if unwrap < 0:
        unwrap = f(unwrap + twopi)                        # Fortran modulo: result has the sign of the divisor
    if optim == 2:
        bpsi = max(int(np.rint(unwrap / psires)), 0)
        if bpsi == psisteps:
            bpsi = 0
        return f(c[bpsi, pixnum])
    psii = int(unwrap / psires)
    psiu = psii + 1
    if psiu >= psisteps:
        psiu = 0
    x0 = f(psii * psires)
    x1 = f(psiu * psires)
    with np.errstate(divide="ignore", invalid="ignore"):
        return f((f(c[psii, pixnum]) * f(x1 - unwrap) + f(c[psiu, pixnum]) * f(unwrap - x0)) / f(x1 - x0))
