"""CPU tests of the oracle itself (no GPU): the definitional oracle against scipy and the
closed forms whose signs the reference fixes; the fast CPU restatement against the
definitional one and against the committed golden vectors."""
import glob
import math
import os

import numpy as np
import pytest

from oracle import sht_def as D

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _grid(nside):
    th, ph = [], []
    for r in range(1, 4 * nside):
        cth, sth, nph, phi0, _ = D.healpix_ring(nside, r)
        th += [math.atan2(sth, cth)] * nph
        ph += list(phi0 + 2 * np.pi * np.arange(nph) / nph)
    return np.array(th), np.array(ph)


def test_ring_offsets_tile_the_sphere():
    for nside in (1, 2, 3, 4, 6, 8, 12, 16):
        pos = 0
        for o, n in sorted((D.healpix_ring(nside, r)[4], D.healpix_ring(nside, r)[2]) for r in range(1, 4 * nside)):
            assert o == pos
            pos += n
        assert pos == 12 * nside ** 2


def test_dipole_convention():
    """map = d.n  <=>  alm(1,+1) = -sqrt(4pi/3) dx, alm(1,0) = sqrt(4pi/3) dz, alm(1,-1) = sqrt(4pi/3) dy
    (commander3/src/comm_cmb_comp_mod.f90:145-156)."""
    nside, lmax = 4, 5
    Y = D.Y_matrix(nside, lmax, 0)
    idx = D.alm_index(lmax, range(lmax + 1))
    d = np.array([0.3, -0.7, 0.5]); c = math.sqrt(4 * np.pi / 3)
    alm = np.zeros(len(idx))
    alm[idx.index((1, 1))] = -c * d[0]; alm[idx.index((1, 0))] = c * d[2]; alm[idx.index((1, -1))] = c * d[1]
    th, ph = _grid(nside)
    ref = d[0] * np.sin(th) * np.cos(ph) + d[1] * np.sin(th) * np.sin(ph) + d[2] * np.cos(th)
    assert np.max(np.abs(Y @ alm - ref)) < 1e-14


def test_monopole_and_spin2_closed_forms():
    """a_00=1 -> 1/sqrt(4pi);  E20, B20, E22 closed forms (COSMO convention,
    commander3/src/comm_map_mod.f90:1002)."""
    nside, lmax = 4, 6
    idx = D.alm_index(lmax, range(lmax + 1)); nalm = len(idx); npix = 12 * nside ** 2
    Y0 = D.Y_matrix(nside, lmax, 0)
    a = np.zeros(nalm); a[0] = 1.0
    assert np.max(np.abs(Y0 @ a - 1 / math.sqrt(4 * np.pi))) < 1e-15
    Y2 = D.Y_matrix(nside, lmax, 2)
    th, ph = _grid(nside)
    k = math.sqrt(15 / (32 * np.pi))

    def synth(E=None, B=None):
        v = np.zeros(2 * nalm)
        for (l, m), x in (E or {}).items(): v[idx.index((l, m))] = x
        for (l, m), x in (B or {}).items(): v[nalm + idx.index((l, m))] = x
        out = Y2 @ v
        return out[:npix], out[npix:]
    Q, U = synth(E={(2, 0): 1.0})
    assert np.max(np.abs(Q + k * np.sin(th) ** 2)) < 1e-14 and np.max(np.abs(U)) < 1e-14
    Q, U = synth(B={(2, 0): 1.0})
    assert np.max(np.abs(U + k * np.sin(th) ** 2)) < 1e-14 and np.max(np.abs(Q)) < 1e-14
    Q, U = synth(E={(2, 2): math.sqrt(2.0)})
    assert np.max(np.abs(Q + 0.25 * math.sqrt(5 / np.pi) * (1 + np.cos(th) ** 2) * np.cos(2 * ph))) < 1e-14
    assert np.max(np.abs(U - 0.5 * math.sqrt(5 / np.pi) * np.cos(th) * np.sin(2 * ph))) < 1e-14
    # l < 2 carries no polarisation
    Q, U = synth(E={(1, 0): 1.0, (1, 1): 1.0, (0, 0): 1.0}, B={(1, -1): 1.0})
    assert np.max(np.abs(Q)) == 0 and np.max(np.abs(U)) == 0


def test_scipy_golden_columns():
    g = np.load(os.path.join(GOLD, "scipy_cols_n8_l12.npz"))
    Y = D.Y_matrix(int(g["nside"]), int(g["lmax"]), 0)
    for slot, col in zip(g["slots"], g["cols"]):
        assert np.max(np.abs(Y[:, slot] - col)) < 5e-15


def test_reference_recurrence_matches_definition():
    """comp_normalised_Plm (commander3/src/math_tools.f90:926-1028) == definition."""
    nside, lmax = 4, 12
    for m in (0, 1, 3, 7, 12):
        for r in (1, 3, 7, 8, 11):
            cth, sth, *_ = D.healpix_ring(nside, r)
            p = D.comp_normalised_Plm(lmax, m, math.atan2(sth, cth))
            ref = np.array([float(D.slam(l, m, 0, cth, sth)) if l >= m else 0 for l in range(lmax + 1)])
            assert np.max(np.abs(p - ref)) < 1e-13


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "dense_*.npz"))))
def test_cpu_restatement_vs_golden(cpu_oracle, path):
    S = cpu_oracle
    g = np.load(path)
    nside, lmax = int(g["nside"]), int(g["lmax"])
    kw = dict(rings=g["rings"], ms=g["ms"], weight=g["weight"])
    for spin in (0, 2):
        alm, mp = g[f"s{spin}_alm"], g[f"s{spin}_map"]
        assert rel(S.execute(S.Y, spin, nside, lmax, alm=alm, **kw), g[f"s{spin}_Y"]) < 1e-13
        assert rel(S.execute(S.WY, spin, nside, lmax, alm=alm, **kw), g[f"s{spin}_WY"]) < 1e-13
        assert rel(S.execute(S.Yt, spin, nside, lmax, map=mp, **kw), g[f"s{spin}_Yt"]) < 1e-13
        assert rel(S.execute(S.YtW, spin, nside, lmax, map=mp, **kw), g[f"s{spin}_YtW"]) < 1e-13


@pytest.mark.parametrize("nside,lmax", [(1, 3), (2, 5), (4, 11), (4, 14), (8, 10), (3, 8), (6, 13)])
def test_cpu_restatement_vs_dense(cpu_oracle, nside, lmax):
    S = cpu_oracle
    rng = np.random.default_rng(nside * 100 + lmax)
    for spin in (0, 2):
        nc = 1 if spin == 0 else 2
        Y = D.Y_matrix(nside, lmax, spin, dps=30)
        alm = rng.standard_normal((nc, S.alm_count(lmax)))
        mp = rng.standard_normal((nc, S.map_size(nside)))
        assert rel(S.execute(S.Y, spin, nside, lmax, alm=alm), (Y @ alm.ravel()).reshape(nc, -1)) < 1e-13
        assert rel(S.execute(S.Yt, spin, nside, lmax, map=mp), (Y.T @ mp.ravel()).reshape(nc, -1)) < 1e-13


@pytest.mark.parametrize("spin", [1, 3, 4])
def test_cpu_restatement_vs_dense_arbitrary_spin(cpu_oracle, spin):
    """Spins other than 0 and 2 (conviqt path): the fast CPU restatement against the definitional matrices."""
    S = cpu_oracle
    nside, lmax = 4, 9
    rng = np.random.default_rng(40 + spin)
    Y = D.Y_matrix(nside, lmax, spin, dps=30)
    alm = rng.standard_normal((2, S.alm_count(lmax)))
    mp = rng.standard_normal((2, S.map_size(nside)))
    assert rel(S.execute(S.Y, spin, nside, lmax, alm=alm), (Y @ alm.ravel()).reshape(2, -1)) < 1e-13
    assert rel(S.execute(S.Yt, spin, nside, lmax, map=mp), (Y.T @ mp.ravel()).reshape(2, -1)) < 1e-13


def test_cpu_restatement_scaled_range(cpu_oracle):
    """Large m near the poles exercises the 2^-800 rescaling; spin-0 columns are compared
    with the reference's own recurrence, mlim skipping with the unskipped result."""
    S = cpu_oracle
    nside, lmax = 128, 380
    nalm = S.alm_count(lmax)

    def off(l, m):
        o = sum((lmax + 1 - mm) * (1 if mm == 0 else 2) for mm in range(m))
        return o + (l - m) * (1 if m == 0 else 2)
    for (l, m) in ((380, 370), (350, 200), (380, 2), (300, 300)):
        alm = np.zeros((1, nalm)); alm[0, off(l, m)] = 1.0
        out = S.execute(S.Y, 0, nside, lmax, alm=alm)
        for r in (2, 40, 128, 200, 255):
            cth, sth, nph, phi0, ofs = D.healpix_ring(nside, r)
            lam = D.comp_normalised_Plm(lmax, m, math.atan2(sth, cth))[l]
            ref = (1.0 if m == 0 else math.sqrt(2.0)) * lam * math.cos(m * (phi0 + 2 * np.pi * 3 / nph))
            # the restatement does not accumulate |lambda| < 2^-100
            assert abs(out[0, ofs + 3] - ref) <= 1e-11 * abs(lam) + 2e-30
    rng = np.random.default_rng(5)
    alm = rng.standard_normal((2, nalm))
    a = S.execute(S.Y, 2, nside, lmax, alm=alm)
    b = S.execute(S.Y, 2, nside, lmax, alm=alm, mlim_skip=True)
    assert rel(b, a) < 1e-13


def test_cpu_adjointness(cpu_oracle):
    S = cpu_oracle
    nside, lmax = 32, 80
    rng = np.random.default_rng(9)
    for spin in (0, 2):
        nc = 1 if spin == 0 else 2
        alm = rng.standard_normal((nc, S.alm_count(lmax)))
        mp = rng.standard_normal((nc, S.map_size(nside)))
        lhs = np.sum(S.execute(S.Y, spin, nside, lmax, alm=alm) * mp)
        rhs = np.sum(alm * S.execute(S.Yt, spin, nside, lmax, map=mp))
        assert abs(lhs - rhs) < 1e-12 * np.linalg.norm(alm) * np.linalg.norm(mp)
