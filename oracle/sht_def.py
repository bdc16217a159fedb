"""Definitional oracle for the Commander3 SHT hot path -- TEST INFRASTRUCTURE ONLY.

This file is a checker.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import it; the product (commander_b200/) never does.

PARITY STATUS: "parity unpinned" -- the reference (ArtemBasyrov/Commander) has no
tests or golden vectors for this path and its arithmetic lives in libsharp2
(HEALPix 3.70 bundle, cmake/project_instructions.cmake:102,106), which is not
under /root/reference and cannot be fetched.  This oracle therefore restates the
published mathematics libsharp2 implements, and is pinned against
  (1) scipy.special.sph_harm_y (independent third-party spin-0 harmonics),
  (2) closed-form known answers whose signs the reference itself fixes:
      dipole      commander3/src/comm_cmb_comp_mod.f90:145-156
      real-packed commander3/src/comm_map_mod.f90:1497-1520 (a_lm=(alm[+m]+i alm[-m])/sqrt2)
      COSMO pol.  commander3/src/comm_map_mod.f90:1002
  (3) the reference's own in-repo Legendre recurrence, restated in
      `comp_normalised_Plm` below (commander3/src/math_tools.f90:926-1028).

What it builds: the dense real matrix Y (pixels x real-packed alm) for a HEALPix
ring subset and an m subset, for spin 0 and spin 2, straight from the definition

    sY_lm(theta,phi) = (-1)^m sqrt((2l+1)/4pi) d^l_{-m,s}(theta) e^{i m phi}

with the Wigner small-d evaluated by its explicit finite sum in 50-digit
arithmetic.  Then (commander3/src/sharp.f90:8-14 job types)
    SHARP_Y   : map = Y alm            SHARP_Yt  : alm = Y^T map
    SHARP_WY  : map = diag(w) Y alm    SHARP_YtW : alm = Y^T diag(w) map
with w_ring = 4 pi / npix * weight[northring-1]   (libsharp2 healpix geometry).
"""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np

try:  # mpmath is only needed for the dense definitional matrices
    import mpmath as mp
except Exception:  # pragma: no cover
    mp = None


# ----------------------------------------------------------------------------
# HEALPix ring geometry (what sharp_make_subset_healpix_geom_info computes;
# called from commander3/src/sharp.f90:145-166, rings chosen at
# commander3/src/comm_map_mod.f90:193-226)
# ----------------------------------------------------------------------------
def healpix_ring(nside: int, ring: int):
    """ring in 1..4*nside-1 -> (cos theta, sin theta, nph, phi0, global RING pixel offset)."""
    npix = 12 * nside * nside
    north = 4 * nside - ring if ring > 2 * nside else ring
    if north < nside:  # polar cap
        # theta = 2 asin(north / (sqrt(6) nside)); use cos = 1 - north^2/(3 nside^2)
        cth = 1.0 - north * north / (3.0 * nside * nside)
        sth = math.sin(2.0 * math.asin(north / (math.sqrt(6.0) * nside)))
        nph = 4 * north
        phi0 = math.pi / nph
        ofs = 2 * north * (north - 1)
    else:  # equatorial belt
        cth = (2 * nside - north) * 8.0 * nside / npix
        sth = math.sqrt((1.0 - cth) * (1.0 + cth))
        nph = 4 * nside
        phi0 = 0.0 if ((north - nside) & 1) else math.pi / nph
        ofs = 2 * nside * (nside - 1) + (north - nside) * 4 * nside
    if north != ring:  # southern hemisphere mirror
        cth = -cth
        ofs = npix - nph - ofs
    return cth, sth, nph, phi0, ofs


def ring_weight(nside: int, ring: int, weight=None) -> float:
    north = 4 * nside - ring if ring > 2 * nside else ring
    w = 4.0 * math.pi / (12 * nside * nside)
    if weight is not None:
        w *= float(weight[north - 1])
    return w


def mapinfo_rings(nside: int, rank: int = 0, nprocs: int = 1):
    """Ring list of one rank, commander3/src/comm_map_mod.f90:197-225."""
    rings = []
    for i in range(1 + rank, 2 * nside + 1, nprocs):
        rings.append(i)
        if i < 2 * nside:
            rings.append(4 * nside - i)
    return sorted(rings)


def mapinfo_ms(lmax: int, rank: int = 0, nprocs: int = 1):
    """m list of one rank, commander3/src/comm_map_mod.f90:231-247."""
    return list(range(rank, lmax + 1, nprocs))


def map_size(nside: int, rings) -> int:
    return sum(healpix_ring(nside, r)[2] for r in rings)


def alm_count(lmax: int, ms) -> int:
    return sum((lmax + 1 - m) * (1 if m == 0 else 2) for m in ms)


def alm_index(lmax: int, ms):
    """Real-packed m-major local index, commander3/src/comm_map_mod.f90:248-260.

    Returns list of (l, m_signed) per local slot."""
    out = []
    for m in ms:
        for l in range(m, lmax + 1):
            out.append((l, m))
            if m > 0:
                out.append((l, -m))
    return out


# ----------------------------------------------------------------------------
# Wigner small-d by the explicit sum, high precision
# ----------------------------------------------------------------------------
@lru_cache(maxsize=None)
def _fact(n: int) -> int:
    return math.factorial(n)


def wigner_d(j: int, m1: int, m2: int, cth_half, sth_half):
    """d^j_{m1,m2}(beta) with cos(beta/2), sin(beta/2) given as mpmath numbers.

    Convention of sympy.physics.quantum.spin.Rotation.d (checked in tests)."""
    smin = max(0, m2 - m1)
    smax = min(j + m2, j - m1)
    pref = mp.sqrt(mp.mpf(_fact(j + m1) * _fact(j - m1) * _fact(j + m2) * _fact(j - m2)))
    tot = mp.mpf(0)
    for s in range(smin, smax + 1):
        den = _fact(j + m2 - s) * _fact(s) * _fact(m1 - m2 + s) * _fact(j - m1 - s)
        term = mp.mpf((-1) ** (m1 - m2 + s)) / den
        term *= cth_half ** (2 * j + m2 - m1 - 2 * s) * sth_half ** (m1 - m2 + 2 * s)
        tot += term
    return pref * tot


def slam(l: int, m: int, s: int, cth: float, sth: float):
    """s-lambda_lm(theta) = sY_lm(theta, 0), m may be negative.  mpmath number."""
    if l < abs(s) or l < abs(m):
        return mp.mpf(0)
    theta = mp.atan2(mp.mpf(sth), mp.mpf(cth))
    ch, sh = mp.cos(theta / 2), mp.sin(theta / 2)
    return (-1) ** (m % 2) * mp.sqrt(mp.mpf(2 * l + 1) / (4 * mp.pi)) * wigner_d(l, -m, s, ch, sh)


def Y_matrix(nside: int, lmax: int, spin: int, rings=None, ms=None, dps: int = 40):
    """Dense real synthesis matrix for the given local rings / m's.

    spin 0: shape (npix_local, nalm_local)
    spin 2: shape (2*npix_local, 2*nalm_local)  rows [Q;U], columns [E;B]
    (column/row blocks are the Fortran columns 2,3 of alm/map,
     commander3/src/comm_map_mod.f90:447)."""
    assert mp is not None
    mp.mp.dps = dps
    rings = list(range(1, 4 * nside)) if rings is None else list(rings)
    ms = list(range(lmax + 1)) if ms is None else list(ms)
    idx = alm_index(lmax, ms)
    nalm = len(idx)
    npix = map_size(nside, rings)
    ncomp = 1 if spin == 0 else 2
    Y = np.zeros((ncomp * npix, ncomp * nalm))
    rt2 = math.sqrt(2.0)
    pofs = 0
    for r in rings:
        cth, sth, nph, phi0, _ = healpix_ring(nside, r)
        phi = phi0 + 2.0 * math.pi * np.arange(nph) / nph
        for k, (l, m) in enumerate(idx):
            am = abs(m)
            # complex a_lm = (alm[+m] + i alm[-m]) / sqrt2 for m>0 ; real for m=0
            # unit real-packed coefficient -> complex coefficient c of a_{l,am}
            if m == 0:
                c = 1.0 + 0.0j
            elif m > 0:
                c = 1.0 / rt2 + 0.0j
            else:
                c = 1.0j / rt2
            e = np.exp(1j * am * phi)
            if spin == 0:
                lam = float(slam(l, am, 0, cth, sth))
                # T = sum_m a_lm Y_lm over all m = a_l0 Y_l0 + 2 Re sum_{m>0} a_lm Y_lm
                col = (c * lam * e).real * (1.0 if am == 0 else 2.0)
                Y[pofs:pofs + nph, k] = col
            else:
                if l < spin:
                    continue
                # (Q +- iU) = ssg sum_{l, all m} (aE_lm +- i aB_lm) (+-s)Y_lm   (ssg = -1 for spin 2)
                lp_p = float(slam(l, am, +spin, cth, sth))    # +2 lambda_{l,+m}
                lm_p = float(slam(l, am, -spin, cth, sth))    # -2 lambda_{l,+m}
                lp_n = float(slam(l, -am, +spin, cth, sth))   # +2 lambda_{l,-m}
                lm_n = float(slam(l, -am, -spin, cth, sth))   # -2 lambda_{l,-m}
                # (+s)a_lm = ssg (E + iB)_lm, (-s)a_lm = ssg (-1)^s (E - iB)_lm = -(E - iB)_lm, with
                # a^X_{l,-m} = (-1)^m conj(a^X_lm) as for T: the two maps are real for every spin (this is the
                # W/X formulation of HEALPix / libsharp2 written with complex coefficients).
                # ssg = -1 for even spin (HEALPix COSMO at spin 2), +1 for odd spin (libsharp2's normalisation as
                # recorded in SURVEY's appendix; unpinned for spin != 2).
                sg = (-1) ** (am % 2)
                ssg = 1.0 if (spin & 1) else -1.0
                for comp_in, (cE, cB) in enumerate(((c, 0.0), (0.0, c))):
                    # positive-m term
                    P = ssg * (cE + 1j * cB) * lp_p * e       # contributes to Q+iU
                    M = -(cE - 1j * cB) * lm_p * e            # contributes to Q-iU
                    if am > 0:
                        cEn, cBn = sg * np.conj(cE), sg * np.conj(cB)
                        P = P + ssg * (cEn + 1j * cBn) * lp_n * np.conj(e)
                        M = M - (cEn - 1j * cBn) * lm_n * np.conj(e)
                    Q = 0.5 * (P + M)
                    U = (P - M) / (2.0j)
                    assert np.max(np.abs(Q.imag)) < 1e-12 and np.max(np.abs(U.imag)) < 1e-12
                    Y[pofs:pofs + nph, comp_in * nalm + k] = Q.real
                    Y[npix + pofs:npix + pofs + nph, comp_in * nalm + k] = U.real
        pofs += nph
    return Y


def ring_weights_vector(nside: int, rings, weight=None, ncomp: int = 1):
    w = []
    for r in rings:
        nph = healpix_ring(nside, r)[2]
        w.append(np.full(nph, ring_weight(nside, r, weight)))
    w = np.concatenate(w) if w else np.zeros(0)
    return np.tile(w, ncomp)


# ----------------------------------------------------------------------------
# The reference's only in-repo Legendre arithmetic, restated
# (commander3/src/math_tools.f90:926-1028, comp_normalised_Plm)
# ----------------------------------------------------------------------------
def comp_normalised_Plm(nlmax: int, m: int, theta: float) -> np.ndarray:
    """lambda_lm(theta) for l=0..nlmax (zeros below m).  HEALPix-style scaled recurrence."""
    plm = np.zeros(nlmax + 1)
    bignorm = 1e-20 * np.finfo(np.float64).max
    cth, sth = math.cos(theta), math.sin(theta)
    lam_mm = bignorm / math.sqrt(4.0 * math.pi)
    for mm in range(1, m + 1):
        f2m = 2.0 * mm
        lam_mm = -lam_mm * sth * math.sqrt((f2m + 1.0) / f2m)
    plm[m] = lam_mm / bignorm
    fm2 = float(m) ** 2
    lam_0, lam_1 = 0.0, 1.0 / bignorm
    fl2 = float(m + 1) ** 2
    a_rec = math.sqrt((4.0 * fl2 - 1.0) / (fl2 - fm2))
    lam_2 = cth * lam_1 * a_rec
    for l in range(m + 1, nlmax + 1):
        plm[l] = lam_2 * lam_mm
        lam_0 = lam_1 / a_rec
        lam_1 = lam_2
        fl2 = float(l + 1) ** 2
        a_rec = math.sqrt((4.0 * fl2 - 1.0) / (fl2 - fm2))
        lam_2 = (cth * lam_1 - lam_0) * a_rec
    return plm
