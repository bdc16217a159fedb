"""The constrained-realisation CG behind the C ABI (cmdr_cr_*, commander_b200/csrc/cr.cu) against the operator composed
from ORACLE transforms and the reference's PCG written out in numpy (commander3/src/comm_cr_mod.f90:201-348, 771-1024).
Tolerances: operator 1e-10 relative L2 (north_star); CG solution within the solver's own convergence tolerance."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _noise(nside, lmax, Cl, scale=1.0):
    from oracle import sht_def as D
    sth2 = np.concatenate([np.full(D.healpix_ring(nside, r)[2], D.healpix_ring(nside, r)[1] ** 2) for r in range(1, 4 * nside)])
    sigma0 = scale * math.sqrt(Cl[lmax // 2, 0] * 12 * nside ** 2 / (4 * math.pi))
    return np.stack([1.0 / (sigma0 * (1 + 0.5 * sth2))] * 3)


def _oracle_ops(S, nside, lmax):
    def Y(alm):
        return np.concatenate([S.execute(S.Y, 0, nside, lmax, alm=alm[0:1]), S.execute(S.Y, 2, nside, lmax, alm=alm[1:3])])

    def Yt(mp):
        return np.concatenate([S.execute(S.Yt, 0, nside, lmax, map=mp[0:1]), S.execute(S.Yt, 2, nside, lmax, map=mp[1:3])])
    return Y, Yt


def _factors(info, Cl, bls):
    l = info.lm[0].astype(np.int64)
    sS = np.stack([np.sqrt(Cl[l, j]) for j in range(3)])
    sS[1:, l < 2] = 0.0
    return sS, [np.stack([b[l, j] for j in range(3)]) for b in bls]


@pytest.mark.parametrize("device", [None, "cuda"])
def test_native_operator_rhs_precond_vs_oracle(shtlib, cpu_oracle, device):
    """two bands: A x, the RHS and the diagonal preconditioner against the oracle composition / the torch mirror"""
    import torch
    from commander_b200 import comm_mapinfo
    from commander_b200.comm_cr import cr_cmb_system, cr_native_system, gaussian_beam
    S = cpu_oracle
    nside, lmax = 32, 64
    rng = np.random.default_rng(41)
    l = np.arange(lmax + 1, dtype=np.float64)
    Cl = np.stack([1.0 / (l * (l + 1) + 1.0)] * 3, axis=1)
    info = comm_mapinfo(None, nside, lmax, 3, True)
    siN = [_noise(nside, lmax, Cl), _noise(nside, lmax, Cl, 1.7) * (1.0 + 0.3 * rng.uniform(size=(3, info.np)))]
    siN[1][:, ::11] = 0.0                                              # masked pixels: N^-1 = 0
    invN = [s * s for s in siN]
    bls = [gaussian_beam(lmax, 200.0), 0.8 * gaussian_beam(lmax, 330.0)]   # the second with a mixing scalar folded in
    put = (lambda a: a) if device is None else (lambda a: torch.as_tensor(a, device=device))
    get = (lambda a: a) if device is None else (lambda a: a.cpu().numpy())
    sysn = cr_native_system(info, [put(v) for v in invN], bls, Cl)
    Y, Yt = _oracle_ops(S, nside, lmax)
    sS, blv = _factors(info, Cl, bls)
    x = rng.standard_normal((3, info.nalm))
    want = x.copy()
    for q in range(2):
        want += sS * blv[q] * Yt(invN[q] * Y(sS * blv[q] * x))
    got = get(sysn.matmulA(put(x)))
    assert np.linalg.norm(got - want) <= 1e-10 * np.linalg.norm(want)
    # RHS with both fluctuation terms
    data = [rng.standard_normal((3, info.np)) for _ in range(2)]
    eta = [rng.standard_normal((3, info.np)) for _ in range(2)]
    eta0 = rng.standard_normal((3, info.nalm))
    wantb = eta0.copy()
    for q in range(2):
        wantb += sS * blv[q] * Yt(invN[q] * data[q] + siN[q] * eta[q])
    gotb = get(sysn.computeRHS([put(d) for d in data], [put(e) for e in eta], put(eta0)))
    assert np.linalg.norm(gotb - wantb) <= 1e-10 * np.linalg.norm(wantb)
    # diagonal preconditioner = 1 / (1 + sum_nu S b^2 N_lm,nu) with N_lm as the torch mirror computes it per band
    dev = torch.device("cuda")
    acc = np.ones((3, info.nalm))
    for q in range(2):
        m = cr_cmb_system(info, torch.as_tensor(siN[q], device=dev), bls[q], Cl)
        acc += (sS * blv[q]) ** 2 * m.invN_lm.cpu().numpy()
    assert np.linalg.norm(sysn.Minv() - 1.0 / acc) <= 1e-10 * np.linalg.norm(1.0 / acc)
    r = rng.standard_normal((3, info.nalm))
    assert np.linalg.norm(get(sysn.invM(put(r))) - r / acc) <= 1e-10 * np.linalg.norm(r / acc)
    # without a prior: A = sum_nu B^t Y^t N^-1 Y B, no unit term
    sys0 = cr_native_system(info, [put(v) for v in invN], bls, None, precond="none")
    l_loc = info.lm[0]
    want0 = np.zeros_like(x)
    for q in range(2):
        f = blv[q].copy(); f[1:, l_loc < 2] = 0.0
        want0 += f * Yt(invN[q] * Y(f * x))
    got0 = get(sys0.matmulA(put(x)))
    assert np.linalg.norm(got0 - want0) <= 1e-10 * np.linalg.norm(want0)


def _numpy_pcg(A, Minv, b, maxiter, tol, crit, miniter=5):
    """solve_cr_eqn_by_CG, commander3/src/comm_cr_mod.f90:201-348, written out"""
    x = np.zeros_like(b); r = b.copy(); d = Minv * r
    dn = float(np.sum(r * d)); d0 = float(np.sum(b * (Minv * b))); lim = tol * d0
    hist, it = [dn], 0
    for i in range(1, maxiter + 1):
        if dn < lim and (i >= miniter or dn <= 1e-30 * d0) and crit != "fixed_iter":
            break
        q = A(d); alpha = dn / float(np.sum(d * q)); x = x + alpha * d; r = r - alpha * q
        s = Minv * r; do = dn; dn = float(np.sum(r * s)); d = s + dn / do * d
        hist.append(dn); it = i
    return x, it, hist


def test_native_cg_vs_oracle_pcg(shtlib, cpu_oracle):
    """residual criterion: same iteration count (+-1) and the same solution as the PCG driven by the oracle SHT"""
    import torch
    from commander_b200 import comm_mapinfo
    from commander_b200.comm_cr import cr_native_system, gaussian_beam
    S = cpu_oracle
    nside, lmax = 32, 64
    rng = np.random.default_rng(5)
    l = np.arange(lmax + 1, dtype=np.float64)
    Cl = np.stack([1.0 / (l * (l + 1) + 1.0)] * 3, axis=1)
    info = comm_mapinfo(None, nside, lmax, 3, True)
    siN = _noise(nside, lmax, Cl)
    bl = gaussian_beam(lmax, 4 * 60.0 * 58.6 / nside / 16)
    sysn = cr_native_system(info, [torch.as_tensor(siN * siN, device="cuda")], [bl], Cl)
    Y, Yt = _oracle_ops(S, nside, lmax)
    sS, blv = _factors(info, Cl, [bl])
    f = sS * blv[0]
    data = Y(f * rng.standard_normal((3, info.nalm))) + rng.standard_normal((3, info.np)) / siN
    b = sysn.computeRHS([data], [rng.standard_normal((3, info.np))], rng.standard_normal((3, info.nalm)))
    x, it, hist = sysn.solve(b, maxiter=200, cg_tol=1e-8, cg_conv_crit="residual")
    assert 3 < it < 200 and hist[-1] < 1e-8 * hist[0] * 10
    Minv = sysn.Minv()
    xo, ito, histo = _numpy_pcg(lambda v: v + f * Yt(siN * siN * Y(f * v)), Minv, b, 200, 1e-8, "residual")
    assert abs(ito - it) <= 1, (it, ito)
    assert np.linalg.norm(x - xo) <= 1e-6 * np.linalg.norm(xo)
    k = min(len(hist), len(histo), 10)
    assert np.allclose(hist[:k], histo[:k], rtol=1e-8)
    # fixed_iter never exits early; an initial guess is honoured
    n0 = sysn.n_matmul
    x2, it2, h2 = sysn.solve(b, maxiter=7, cg_conv_crit="fixed_iter")
    assert it2 == 7 and sysn.n_matmul - n0 == 8 and len(h2) == 8
    x3, it3, h3 = sysn.solve(b, x0=x, maxiter=50, cg_tol=1e-8, cg_conv_crit="residual", cg_miniter=1)
    assert it3 <= 2 and np.linalg.norm(x3 - xo) <= 1e-6 * np.linalg.norm(xo)


def test_native_cg_config3_size_three_iterations(shtlib, cpu_oracle):
    """BASELINE.json configs[2] at its real size (nside 1024, lmax 2000, IQU, diagonal N^-1, 10' Gaussian beam):
    three `fixed_iter` iterations of the native CG against the same three iterations driven by the oracle SHT.
    Iterate-by-iterate agreement: residual history to 1e-9 relative, x to 1e-9 relative L2."""
    import torch
    from commander_b200 import comm_mapinfo
    from commander_b200.comm_cr import cr_native_system, gaussian_beam
    S = cpu_oracle
    nside, lmax = 1024, 2000
    rng = np.random.default_rng(7)
    l = np.arange(lmax + 1, dtype=np.float64)
    Cl = np.stack([1.0 / (l * (l + 1) + 1.0)] * 3, axis=1)
    info = comm_mapinfo(None, nside, lmax, 3, True)
    north = np.minimum(info.rings, 4 * nside - info.rings).astype(np.float64)
    z = np.where(north < nside, 1.0 - north ** 2 / (3.0 * nside ** 2), (2.0 * nside - north) * 2.0 / (3.0 * nside))
    counts = np.where(north < nside, 4 * north, 4 * nside).astype(np.int64)
    sth2 = np.repeat(1.0 - z * z, counts)
    sigma0 = math.sqrt(Cl[1000, 0] * 12 * nside ** 2 / (4 * math.pi))
    siN = np.stack([1.0 / (sigma0 * (1 + 0.5 * sth2))] * 3)
    bl = gaussian_beam(lmax, 10.0)
    sysn = cr_native_system(info, [torch.as_tensor(siN * siN, device="cuda")], [bl], Cl)
    Y, Yt = _oracle_ops(S, nside, lmax)
    sS, blv = _factors(info, Cl, [bl])
    f = sS * blv[0]
    data = rng.standard_normal((3, info.np)) / siN
    b = sysn.computeRHS([data])
    bo = f * Yt(siN * siN * data)
    assert np.linalg.norm(b - bo) <= 1e-10 * np.linalg.norm(bo)
    x, it, hist = sysn.solve(b, maxiter=3, cg_conv_crit="fixed_iter")
    xo, ito, histo = _numpy_pcg(lambda v: v + f * Yt(siN * siN * Y(f * v)), sysn.Minv(), bo, 3, 1e-8, "fixed_iter")
    assert it == ito == 3
    assert np.allclose(hist, histo, rtol=1e-9), (hist, histo)
    assert np.linalg.norm(x - xo) <= 1e-9 * np.linalg.norm(xo)
