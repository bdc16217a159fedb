"""GPU test of the constrained-realisation CG (BASELINE.json configs[2], scaled down so the CPU
oracle finishes in seconds): the same PCG (commander3/src/comm_cr_mod.f90:201-348) run once with
the CUDA SHT and once with the CPU oracle SHT must agree within the solver tolerance and take the
same number of iterations (+-1)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup(nside, lmax, seed):
    from oracle import sht_def as D
    rng = np.random.default_rng(seed)
    l = np.arange(lmax + 1, dtype=np.float64)
    Cl = np.stack([1.0 / (l * (l + 1) + 1.0)] * 3, axis=1)
    # noise: sigma_p = sigma0 (1 + 0.5 sin^2 theta), S/N ~ 1 at l ~ lmax/2
    sth2 = np.concatenate([np.full(D.healpix_ring(nside, r)[2], D.healpix_ring(nside, r)[1] ** 2)
                           for r in range(1, 4 * nside)])
    lh = lmax // 2
    sigma0 = math.sqrt(Cl[lh, 0] * 12 * nside ** 2 / (4 * math.pi))
    siN = np.stack([1.0 / (sigma0 * (1 + 0.5 * sth2))] * 3)
    return rng, Cl, siN


def test_cg_gpu_vs_oracle(shtlib, cpu_oracle):
    import torch
    from commander_b200 import comm_map, comm_mapinfo
    from commander_b200.comm_cr import cr_cmb_system, gaussian_beam, solve_cr_eqn_by_CG
    S = cpu_oracle
    nside, lmax = 32, 64
    rng, Cl, siN = _setup(nside, lmax, 5)
    info = comm_mapinfo(None, nside, lmax, 3, True)
    bl = gaussian_beam(lmax, 4 * 60.0 * 58.6 / nside / 16)   # a few pixels
    dev = torch.device("cuda")
    sysm = cr_cmb_system(info, torch.as_tensor(siN, device=dev), bl, Cl)
    # data = Y B sqrt(S) xi + sigma eta
    xi = rng.standard_normal((3, info.nalm))
    sig = comm_map(info, device=dev)
    sig.alm.copy_(torch.as_tensor(xi, device=dev) * sysm.sqrtS * sysm.bl)
    sig.Y()
    data = sig.map + torch.as_tensor(rng.standard_normal((3, info.np)) / siN, device=dev)
    eta_pix = torch.as_tensor(rng.standard_normal((3, info.np)), device=dev)
    eta_alm = torch.as_tensor(rng.standard_normal((3, info.nalm)), device=dev)
    b = sysm.computeRHS(data, eta_pix, eta_alm)
    x, it, hist = solve_cr_eqn_by_CG(sysm, b, maxiter=200, cg_tol=1e-8, cg_conv_crit="residual")
    assert 3 < it < 200 and hist[-1] < 1e-8 * hist[0] * 10

    # the same solver with the CPU oracle as SHT engine (numpy)
    sqrtS, blv, invN, Minv = (t.cpu().numpy() for t in (sysm.sqrtS, sysm.bl, sysm.invN, sysm.Minv))

    def Y(alm):
        return np.concatenate([S.execute(S.Y, 0, nside, lmax, alm=alm[0:1]), S.execute(S.Y, 2, nside, lmax, alm=alm[1:3])])

    def Yt(mp):
        return np.concatenate([S.execute(S.Yt, 0, nside, lmax, map=mp[0:1]), S.execute(S.Yt, 2, nside, lmax, map=mp[1:3])])

    def A(v):
        return v + sqrtS * blv * Yt(invN * Y(sqrtS * blv * v))
    bb = sqrtS * blv * Yt(data.cpu().numpy() * invN + siN * eta_pix.cpu().numpy()) + eta_alm.cpu().numpy()
    assert np.linalg.norm(bb - b.cpu().numpy()) <= 1e-10 * np.linalg.norm(bb)
    xo = np.zeros_like(bb); r = bb - A(xo); d = Minv * r
    dn = float(np.sum(r * d)); d0 = float(np.sum(bb * (Minv * bb))); lim = 1e-8 * d0
    ito = 0
    for i in range(1, 201):
        if dn < lim and i >= 5:
            break
        q = A(d); alpha = dn / float(np.sum(d * q)); xo = xo + alpha * d; r = r - alpha * q
        s = Minv * r; do = dn; dn = float(np.sum(r * s)); d = s + dn / do * d; ito = i
    assert abs(ito - it) <= 1
    err = np.linalg.norm(x.cpu().numpy() - xo) / np.linalg.norm(xo)
    assert err <= 1e-6, err     # both are converged to cg_tol = 1e-8 on the preconditioned residual


def test_cg_fixed_iter_runs_maxiter(shtlib):
    """criterion 'fixed_iter' (the shipped production setting) never exits early."""
    import torch
    from commander_b200 import comm_mapinfo
    from commander_b200.comm_cr import cr_cmb_system, gaussian_beam, solve_cr_eqn_by_CG
    nside, lmax = 64, 128
    rng, Cl, siN = _setup(nside, lmax, 6)
    info = comm_mapinfo(None, nside, lmax, 3, True)
    dev = torch.device("cuda")
    sysm = cr_cmb_system(info, torch.as_tensor(siN, device=dev), gaussian_beam(lmax, 60.0), Cl)
    b = sysm.computeRHS(torch.as_tensor(rng.standard_normal((3, info.np)), device=dev))
    n0 = sysm.n_matmul
    x, it, hist = solve_cr_eqn_by_CG(sysm, b, maxiter=12, cg_conv_crit="fixed_iter")
    assert it == 12 and sysm.n_matmul - n0 == 13 and all(np.isfinite(hist))


def test_cg_pseudoinv_preconditioner(shtlib):
    """precond_type 'pseudoinv' (commander3/src/comm_diffuse_comp_mod.f90:2237-2380): same solution as with the
    diagonal preconditioner, in fewer iterations (its point: T^+ = YtW N WY resolves the spatially varying noise)."""
    import torch
    from commander_b200 import comm_map, comm_mapinfo
    from commander_b200.comm_cr import cr_cmb_system, gaussian_beam, solve_cr_eqn_by_CG
    nside, lmax = 64, 128
    rng, Cl, siN = _setup(nside, lmax, 11)
    siN = siN * (1.0 + 3.0 * (np.arange(siN.shape[1]) % 7 == 0))     # strongly inhomogeneous noise
    info = comm_mapinfo(None, nside, lmax, 3, True)
    dev = torch.device("cuda")
    bl = gaussian_beam(lmax, 40.0)
    res = {}
    for pc in ("diagonal", "pseudoinv"):
        sysm = cr_cmb_system(info, torch.as_tensor(siN, device=dev), bl, Cl, precond=pc)
        xi = np.random.default_rng(12).standard_normal((3, info.nalm))
        sig = comm_map(info, device=dev)
        sig.alm.copy_(torch.as_tensor(xi, device=dev) * sysm.sqrtS * sysm.bl)
        sig.Y()
        data = sig.map + torch.as_tensor(np.random.default_rng(13).standard_normal((3, info.np)) / siN, device=dev)
        b = sysm.computeRHS(data)
        x, it, hist = solve_cr_eqn_by_CG(sysm, b, maxiter=300, cg_tol=1e-12, cg_conv_crit="residual")
        res[pc] = (x, it)
    xd, itd = res["diagonal"]
    xp, itp = res["pseudoinv"]
    assert float((xd - xp).norm() / xd.norm()) <= 2e-5     # both converged to 1e-12 on their own preconditioned residual
    assert itp < itd, (itp, itd)


def test_cg_full_invN_lm_preconditioner(shtlib):
    """Diagonal preconditioner with the full compute_invN_lm (commander3/src/comm_N_mod.f90:127-197, GPU kernel
    cmdr_sht_invn_diag) against its monopole-only approximation, for noise that varies with latitude (the term
    the 3j sum captures): same solution, comparable iteration count (each run stops on its own preconditioned
    residual, so the counts are not ordered: 39 vs 36 here on B200)."""
    import torch
    from commander_b200 import comm_map, comm_mapinfo
    from commander_b200.comm_cr import cr_cmb_system, gaussian_beam, solve_cr_eqn_by_CG
    nside, lmax = 64, 128
    rng, Cl, siN = _setup(nside, lmax, 21)
    info = comm_mapinfo(None, nside, lmax, 3, True)
    # strongly latitude-dependent depth: 4x deeper at the poles than on the equator
    z = np.concatenate([np.full(4 * min(r, 4 * nside - r, nside), 1.0 - (r / (2.0 * nside)) if r <= 2 * nside else (r / (2.0 * nside)) - 1.0)
                        for r in info.rings])
    siN = siN * (1.0 + 3.0 * z ** 2)[None, :]
    dev = torch.device("cuda")
    bl = gaussian_beam(lmax, 40.0)
    res = {}
    for mode in ("monopole", "wigner"):
        sysm = cr_cmb_system(info, torch.as_tensor(siN, device=dev), bl, Cl, invN_lm=mode)
        xi = np.random.default_rng(22).standard_normal((3, info.nalm))
        sig = comm_map(info, device=dev)
        sig.alm.copy_(torch.as_tensor(xi, device=dev) * sysm.sqrtS * sysm.bl)
        sig.Y()
        data = sig.map + torch.as_tensor(np.random.default_rng(23).standard_normal((3, info.np)) / siN, device=dev)
        b = sysm.computeRHS(data)
        x, it, hist = solve_cr_eqn_by_CG(sysm, b, maxiter=400, cg_tol=1e-12, cg_conv_crit="residual")
        res[mode] = (x, it)
        if mode == "wigner":
            # N_lm is positive and its (l,m)-average equals the monopole term
            assert float(sysm.invN_lm.min()) > 0.0
    xm, itm = res["monopole"]
    xw, itw = res["wigner"]
    assert float((xm - xw).norm() / xm.norm()) <= 2e-5
    assert itw <= 1.25 * itm + 2, (itw, itm)
    print("CG iterations: monopole", itm, "full N_lm", itw)
