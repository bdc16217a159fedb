"""CPU tests of the conviqt restatement (oracle/conviqt.py) and of the host logic of the mirror
(commander_b200/comm_conviqt.py; commander3/src/comm_conviqt_mod.f90).  No compute call into the CUDA library."""
import math

import numpy as np

from oracle import conviqt as O


def test_c2r_is_fftw_c2r():
    """The plain-sum c2r equals numpy's irfft * n (same definition as FFTW c2r: backward sign, unnormalised,
    imaginary parts of the DC and Nyquist inputs ignored, comm_conviqt_mod.f90:267-279)."""
    rng = np.random.default_rng(0)
    for bmax in (1, 2, 3, 5, 8):
        n = 2 * bmax
        dv = rng.standard_normal(bmax + 1) + 1j * rng.standard_normal(bmax + 1)
        ref = np.fft.irfft(dv, n=n) * n
        assert np.allclose(O.c2r(dv, n), ref, rtol=0, atol=1e-13 * np.abs(ref).max())


def test_get_alms_closed_form():
    """v1 + conj(v2) mfac = 2 sum_c s_c Re b_c and v1 - conj(v2) mfac = 2i sum_c s_c Im b_c, so the two columns are
    spinsign * lnorm * 2 * sum_c s_c {Re, Im}(b_c) in the complex basis."""
    lmax, nmaps = 9, 3
    rng = np.random.default_rng(1)
    lm = O.lm_table(lmax)
    idx = O.lm2i(lm)
    sky = rng.standard_normal((nmaps, len(lm)))
    ntri = (lmax + 1) * (lmax + 2) // 2
    beam = (rng.standard_normal((ntri, nmaps)) + 1j * rng.standard_normal((ntri, nmaps))).astype(np.complex64)
    for m_b in (0, 1, 4):
        alm = O.get_alms(m_b, lmax, lm, sky, beam)
        for (l, m), i in idx.items():
            if m < 0:
                continue
            if l < m_b:
                assert alm[0, i] == 0 and alm[1, i] == 0
                continue
            s = sky[:, i] + 0j if m == 0 else (sky[:, i] + 1j * sky[:, idx[(l, -m)]]) / math.sqrt(2.0)
            b = beam[l * (l + 1) // 2 + m_b].astype(np.complex128)
            f = (-1.0 if m_b else 1.0) * 0.5 * math.sqrt(4 * math.pi / (2 * l + 1)) * 2.0
            pos, neg = f * np.sum(s * b.real), f * np.sum(s * b.imag)
            got_p = alm[0, i] + 0j if m == 0 else (alm[0, i] + 1j * alm[0, idx[(l, -m)]]) / math.sqrt(2.0)
            got_n = alm[1, i] + 0j if m == 0 else (alm[1, i] + 1j * alm[1, idx[(l, -m)]]) / math.sqrt(2.0)
            if m == 0:
                pos, neg = pos.real, neg.real
            assert abs(got_p - pos) <= 1e-13 * (1 + abs(pos))
            if m_b:
                assert abs(got_n - neg) <= 1e-13 * (1 + abs(neg))
            else:
                assert got_n == 0


def test_oracle_cube_axisymmetric_beam_is_plain_convolution(cpu_oracle):
    """Beam with m = 0 coefficients only, b_l0 = sqrt((2l+1)/4pi) B_l: every psi plane of the cube is the
    map smoothed with B_l (convolution theorem), through both c2r implementations."""
    S = cpu_oracle
    nside, lmax, bmax = 4, 8, 3
    rng = np.random.default_rng(2)
    lm = O.lm_table(lmax)
    sky = rng.standard_normal((1, len(lm)))
    B = np.exp(-0.02 * np.arange(lmax + 1) * (np.arange(lmax + 1) + 1))
    balm = np.zeros((1, len(lm)))
    for i, (l, m) in enumerate(lm):
        if m == 0:
            balm[0, i] = math.sqrt((2 * l + 1) / (4 * math.pi)) * B[l]
    beam = O.beam_table(lmax, 1, lm, balm)
    c1 = O.precompute_sky(S, nside, lmax, bmax, sky, beam, vectorised=True)
    c2 = O.precompute_sky(S, nside, lmax, bmax, sky, beam, vectorised=False)
    assert np.allclose(c1, c2, rtol=0, atol=1e-13 * np.abs(c1).max())
    l = np.array([t[0] for t in lm])
    bl32 = np.array([float(np.float32(math.sqrt((2 * ll + 1) / (4 * math.pi)) * B[ll])) / math.sqrt((2 * ll + 1) / (4 * math.pi))
                     for ll in range(lmax + 1)])          # the beam table is single precision
    ref = S.execute(S.Y, 0, nside, lmax, alm=sky * bl32[l])[0]
    for k in range(2 * bmax):
        assert np.linalg.norm(c1[k] - ref) <= 1e-12 * np.linalg.norm(ref)


def test_mirror_host_logic_matches_oracle(shtlib):
    """Beam table (bitwise), vectorised get_alms and interp of the mirror against the loop restatement."""
    from commander_b200 import comm_map, comm_mapinfo
    from commander_b200.comm_conviqt import comm_conviqt
    nside, lmax, lmax_beam, bmax, nmaps = 4, 10, 7, 3, 3
    rng = np.random.default_rng(3)
    info = comm_mapinfo(None, nside, lmax, nmaps, True)
    binfo = comm_mapinfo(None, nside, lmax_beam, nmaps, True)
    sky, beam = comm_map(info), comm_map(binfo)
    sky.alm[...] = rng.standard_normal(sky.alm.shape)
    beam.alm[...] = rng.standard_normal(beam.alm.shape)
    cv = comm_conviqt(nside, lmax, nmaps, bmax, beam, sky, precompute=False)
    lm = [tuple(int(x) for x in info.lm[:, i]) for i in range(info.nalm)]
    assert lm == O.lm_table(lmax)
    # the oracle's beam table wants the beam on the sky layout: re-pack with the reference's own rule (alm_equal)
    b2 = comm_map(info)
    beam.alm_equal(b2)
    tab = O.beam_table(lmax, nmaps, lm, b2.alm, beam_lmax=lmax_beam)
    assert cv.alm_beam.dtype == np.complex64 and np.array_equal(cv.alm_beam.view(np.float32), tab.view(np.float32))
    for m_b in range(bmax + 1):
        ref = O.get_alms(m_b, lmax, lm, sky.alm, tab)
        got = cv.get_alms(m_b, sky)
        assert np.allclose(got, ref, rtol=0, atol=1e-14 * np.abs(ref).max())
    cv.c[...] = rng.standard_normal(cv.c.shape).astype(np.float32)
    pix = rng.integers(0, info.npix, 200)
    psi = rng.uniform(-10, 10, 200)
    for optim in (0, 2):
        cv.optim = optim
        got = cv.interp(pix, psi)
        ref = np.array([O.interp(cv.c, cv.psisteps, int(p), float(a), optim) for p, a in zip(pix, psi)], dtype=np.float32)
        assert got.dtype == np.float32 and np.array_equal(got, ref, equal_nan=True)


import glob
import os

import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "conviqt_*.npz"))))
def test_oracle_cube_vs_golden(cpu_oracle, path):
    """The committed cubes were made with the dense definitional spin-j matrices (tests/golden/make_golden.py);
    the restatement with the fast CPU transforms must reproduce them."""
    g = np.load(path)
    nside, lmax, bmax = int(g["nside"]), int(g["lmax"]), int(g["bmax"])
    lm = O.lm_table(lmax)
    tab = O.beam_table(lmax, 3, lm, g["beam_alm"])
    for vec in (True, False):
        cube = O.precompute_sky(cpu_oracle, nside, lmax, bmax, g["sky_alm"], tab, vectorised=vec)
        assert np.linalg.norm(cube - g["cube"]) <= 1e-12 * np.linalg.norm(g["cube"])
