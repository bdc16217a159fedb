"""GPU tests of the general constrained-realisation system (several diffuse components x several bands):
commander_b200/comm_cr.py::cr_system, the mirror of cr_matmulA / cr_computeRHS / the diagonal preconditioner
(commander3/src/comm_cr_mod.f90:771-1024, 542-769; comm_diffuse_comp_mod.f90:1167-1557, 2186-2235), against the same
operator composed from CPU-oracle transforms in numpy.  Tolerances: operator and right-hand side 1e-10 relative L2;
CG solution within the solver tolerance (stated in the test)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _lm(lmax):
    """(l, m) per slot of the single-rank real-packed layout and its inverse lookup."""
    l, m = [], []
    for mm in range(lmax + 1):
        ls = np.arange(mm, lmax + 1)
        if mm == 0:
            l.append(ls); m.append(np.zeros_like(ls))
        else:
            l.append(np.repeat(ls, 2)); m.append(np.tile([mm, -mm], ls.size))
    l, m = np.concatenate(l), np.concatenate(m)
    return l, m, {(int(a), int(b)): i for i, (a, b) in enumerate(zip(l, m))}


def _repack(a, lmax_from, lmax_to, nmaps_to=None):
    """alm_equal of the reference (comm_map_mod.f90:1148-1176) written independently: copy the common (l, m)."""
    lf, mf, _ = _lm(lmax_from)
    lt, mt, pos = _lm(lmax_to)
    nm = a.shape[0] if nmaps_to is None else nmaps_to
    out = np.zeros((nm, lt.size))
    q = min(nm, a.shape[0])
    sel = lf <= lmax_to
    idx = np.array([pos[(int(x), int(y))] for x, y in zip(lf[sel], mf[sel])], dtype=np.int64)
    out[:q, idx] = a[:q, sel]
    return out


def _Y(S, nside, lmax, alm):
    if alm.shape[0] == 1:
        return S.execute(S.Y, 0, nside, lmax, alm=alm)
    return np.concatenate([S.execute(S.Y, 0, nside, lmax, alm=alm[0:1]), S.execute(S.Y, 2, nside, lmax, alm=alm[1:3])])


def _Yt(S, nside, lmax, mp, job=None):
    job = S.Yt if job is None else job
    if mp.shape[0] == 1:
        return S.execute(job, 0, nside, lmax, map=mp)
    return np.concatenate([S.execute(job, 0, nside, lmax, map=mp[0:1]), S.execute(job, 2, nside, lmax, map=mp[1:3])])


def _problem(seed=3, pixel_F=False):
    """3 bands (different nside / lmax / beam / noise), 3 components (CMB with prior; dust with prior, lower lmax,
    band-dependent mixing; a temperature-only low-l component without prior)."""
    from commander_b200.comm_cr import gaussian_beam
    rng = np.random.default_rng(seed)
    bands = [dict(nside=16, lmax=32, fwhm=600.0, sig=1.0), dict(nside=16, lmax=40, fwhm=420.0, sig=0.7),
             dict(nside=32, lmax=48, fwhm=300.0, sig=1.3)]
    for b in bands:
        npix = 12 * b["nside"] ** 2
        b["b_l"] = gaussian_beam(b["lmax"], b["fwhm"], 3)
        b["invN"] = (1.0 / (b["sig"] * (1.0 + 0.5 * rng.uniform(size=(3, npix)))) ** 2)
    comps = []
    l = np.arange(49, dtype=np.float64)
    comps.append(dict(nside=32, lmax=40, nmaps=3, Cl=np.stack([40.0 / (l * (l + 1) + 1.0)] * 3, axis=1)[:41],
                      F_mean=np.ones((3, 3))))
    comps.append(dict(nside=32, lmax=32, nmaps=3, Cl=np.stack([8.0 / (l + 1.0) ** 2.5] * 3, axis=1)[:33],
                      F_mean=np.array([[0.3] * 3, [1.0] * 3, [2.5] * 3])))
    comps.append(dict(nside=32, lmax=8, nmaps=1, Cl=None, F_mean=np.array([[1.5], [1.0], [0.5]])))
    for c in comps:
        c["F"] = None
    if pixel_F:      # spatially varying mixing of the dust component in the last band (Y, x F, YtW)
        comps[1]["F"] = [None, None, 2.5 * (1.0 + 0.2 * rng.uniform(size=(3, 12 * 32 ** 2)))]
    return bands, comps, rng


def _oracle_A(S, bands, comps, x_parts, with_prior=True):
    lmax_all = max(max(c["lmax"] for c in comps), 2)
    sx = []
    for c, x in zip(comps, x_parts):
        if c["Cl"] is None:
            sx.append(x.copy())
        else:
            l, _, _ = _lm(c["lmax"])
            sS = np.sqrt(c["Cl"][l, :c["nmaps"]]).T.copy()
            sS[1:, l < 2] = 0.0
            sx.append(x * sS)
    y = [np.zeros_like(x) for x in x_parts]
    for q, b in enumerate(bands):
        lb, _, _ = _lm(b["lmax"])
        band_alm = np.zeros((3, lb.size))
        for c, s in zip(comps, sx):
            nm = min(3, c["nmaps"])
            a = _repack(s, c["lmax"], b["lmax"])[:nm]
            if c["F"] is not None and c["F"][q] is not None:
                mp = _Y(S, b["nside"], b["lmax"], a) * c["F"][q][:nm]
                a = _Yt(S, b["nside"], b["lmax"], mp, job=S.YtW)
            else:
                a = a * c["F_mean"][q, :nm, None]
            band_alm[:nm] += a * b["b_l"][lb, :nm].T
        mp = _Y(S, b["nside"], lmax_all, _repack(band_alm, b["lmax"], lmax_all)) * b["invN"]
        back = _repack(_Yt(S, b["nside"], lmax_all, mp), lmax_all, b["lmax"])
        for i, c in enumerate(comps):
            nm = min(3, c["nmaps"])
            a = back[:nm] * b["b_l"][lb, :nm].T
            if c["F"] is not None and c["F"][q] is not None:
                mp2 = _Y(S, b["nside"], b["lmax"], a) * c["F"][q][:nm]
                a = _Yt(S, b["nside"], b["lmax"], mp2, job=S.YtW)
            else:
                a = a * c["F_mean"][q, :nm, None]
            y[i] += _repack(a, b["lmax"], c["lmax"], nmaps_to=c["nmaps"])
    out = []
    for c, x, yy in zip(comps, x_parts, y):
        if c["Cl"] is None:
            out.append(yy)
        else:
            l, _, _ = _lm(c["lmax"])
            sS = np.sqrt(c["Cl"][l, :c["nmaps"]]).T.copy()
            sS[1:, l < 2] = 0.0
            out.append(yy * sS + (x if with_prior else 0.0))
    return out


def _build(bands, comps, dev):
    import torch
    from commander_b200 import comm_mapinfo
    from commander_b200.comm_cr import cr_band, cr_component, cr_system
    cb = [cr_band(comm_mapinfo(None, b["nside"], b["lmax"], 3, True), torch.as_tensor(b["invN"], device=dev), b["b_l"])
          for b in bands]
    cc = [cr_component(comm_mapinfo(None, c["nside"], c["lmax"], c["nmaps"], c["nmaps"] == 3), Cl=c["Cl"],
                       F_mean=c["F_mean"], F=c["F"]) for c in comps]
    return cr_system(cc, cb, dev)


@pytest.mark.parametrize("pixel_F", [False, True])
def test_matmulA_and_rhs_vs_oracle(shtlib, cpu_oracle, pixel_F):
    import torch
    S = cpu_oracle
    dev = torch.device("cuda", 0)
    bands, comps, rng = _problem(3, pixel_F)
    sysm = _build(bands, comps, dev)
    xs = [rng.standard_normal((c["nmaps"], _lm(c["lmax"])[0].size)) for c in comps]
    x = torch.as_tensor(np.concatenate([v.ravel() for v in xs]), device=dev)
    assert x.numel() == sysm.ncr
    got = sysm.matmulA(x).cpu().numpy()
    ref = np.concatenate([v.ravel() for v in _oracle_A(S, bands, comps, xs)])
    assert rel(got, ref) <= 1e-10, rel(got, ref)
    # symmetry of A
    ys = torch.as_tensor(rng.standard_normal(sysm.ncr), device=dev)
    a, b = float(torch.dot(ys, sysm.matmulA(x))), float(torch.dot(x, sysm.matmulA(ys)))
    assert abs(a - b) <= 1e-11 * max(abs(a), abs(b))
    # right-hand side, mean-field term: sqrt(S) sum_nu F^t B^t Y^t N^-1 d
    data = [rng.standard_normal((3, 12 * b["nside"] ** 2)) for b in bands]
    rhs = sysm.computeRHS([torch.as_tensor(d, device=dev) for d in data]).cpu().numpy()
    lmax_all = max(c["lmax"] for c in comps)
    parts = [np.zeros_like(v) for v in xs]
    for q, b in enumerate(bands):
        lb = _lm(b["lmax"])[0]
        back = _repack(_Yt(S, b["nside"], lmax_all, data[q] * b["invN"]), lmax_all, b["lmax"])
        for i, c in enumerate(comps):
            nm = min(3, c["nmaps"])
            a = back[:nm] * b["b_l"][lb, :nm].T
            if c["F"] is not None and c["F"][q] is not None:
                a = _Yt(S, b["nside"], b["lmax"], _Y(S, b["nside"], b["lmax"], a) * c["F"][q][:nm], job=S.YtW)
            else:
                a = a * c["F_mean"][q, :nm, None]
            parts[i] += _repack(a, b["lmax"], c["lmax"], nmaps_to=c["nmaps"])
    for c, p in zip(comps, parts):
        if c["Cl"] is not None:
            l = _lm(c["lmax"])[0]
            sS = np.sqrt(c["Cl"][l, :c["nmaps"]]).T.copy()
            sS[1:, l < 2] = 0.0
            p *= sS
    ref_rhs = np.concatenate([p.ravel() for p in parts])
    assert rel(rhs, ref_rhs) <= 1e-10, rel(rhs, ref_rhs)


def test_multi_component_cg(shtlib, cpu_oracle):
    """PCG on the 3-band / 3-component system: converges, the block-diagonal preconditioner beats none, and the solution
    satisfies the oracle-composed normal equations to the solver tolerance."""
    import torch
    from commander_b200.comm_cr import solve_cr_eqn_by_CG
    S = cpu_oracle
    dev = torch.device("cuda", 0)
    bands, comps, rng = _problem(4, False)
    sysm = _build(bands, comps, dev)
    data = [torch.as_tensor(rng.standard_normal((3, 12 * b["nside"] ** 2)) / np.sqrt(b["invN"]), device=dev) for b in bands]
    b = sysm.computeRHS(data)
    x, it, hist = solve_cr_eqn_by_CG(sysm, b, maxiter=500, cg_tol=1e-10, cg_conv_crit="residual")
    assert 3 < it < 500 and hist[-1] <= 1e-10 * hist[0] * 10

    class NoPre:
        def __getattr__(self, k):
            return getattr(sysm, k)

        def invM(self, r):
            return r.clone()
    x2, it2, _ = solve_cr_eqn_by_CG(NoPre(), b, maxiter=3000, cg_tol=1e-10, cg_conv_crit="residual")
    assert it <= it2, (it, it2)
    assert float((x - x2).norm() / x.norm()) <= 1e-3
    # residual of the oracle-composed system at the GPU solution
    xs, o = [], 0
    for c in comps:
        n = c["nmaps"] * _lm(c["lmax"])[0].size
        xs.append(x[o:o + n].cpu().numpy().reshape(c["nmaps"], -1)); o += n
    Ax = np.concatenate([v.ravel() for v in _oracle_A(S, bands, comps, xs)])
    bb = b.cpu().numpy()
    assert np.linalg.norm(Ax - bb) <= 1e-4 * np.linalg.norm(bb)


def _oracle_pseudoinv(S, bands, comps, r_parts):
    """applyDiffPrecond_pseudoinv (commander3/src/comm_diffuse_comp_mod.f90:2237-2380) with alpha_nu
    (comm_N_rms_mod.f90:218-247) and the per-(l, pol) pseudo-inverse (:1560-1658), composed from oracle transforms."""
    lmax_pre = max(max(c["lmax"] for c in comps), 2)
    nb, npre = len(bands), len(comps)
    lp, mp_, pos_pre = _lm(lmax_pre)
    alpha = []
    for b in bands:
        tau = _Y(S, b["nside"], b["lmax"], _Yt(S, b["nside"], b["lmax"], b["invN"]))
        a = np.zeros(3)
        a[0] = np.sqrt((tau[0] ** 2).sum() / tau[0].sum())
        a[1:] = np.sqrt((tau[1:] ** 2).sum() / tau[1:].sum())
        alpha.append(a)
    M = np.zeros((3, lmax_pre + 1, npre, nb + npre))
    for j in range(3):
        for l in range(lmax_pre + 1):
            V = np.zeros((nb + npre, npre))
            for q, b in enumerate(bands):
                if l > b["lmax"]:
                    continue
                for k, c in enumerate(comps):
                    if l > c["lmax"] or j >= c["nmaps"]:
                        continue
                    V[q, k] = alpha[q][j] * b["b_l"][l, j] * c["F_mean"][q, j]
                    if c["Cl"] is not None:
                        V[q, k] *= 0.0 if (j > 0 and l < 2) else np.sqrt(c["Cl"][l, j])
            for k, c in enumerate(comps):
                if c["Cl"] is not None and l <= c["lmax"] and j < c["nmaps"]:
                    V[nb + k, k] = 1.0
            M[j, l] = np.linalg.pinv(V)
    y = np.zeros((npre, 3, lp.size))
    for i, (c, r) in enumerate(zip(comps, r_parts)):
        y[i] = _repack(r, c["lmax"], lmax_pre, nmaps_to=3)
    z = np.zeros_like(y)
    for q, b in enumerate(bands):
        lb, mb, _ = _lm(b["lmax"])
        sel = lb <= lmax_pre
        jpre = np.array([pos_pre[(int(a), int(c_))] for a, c_ in zip(lb[sel], mb[sel])])
        alm = np.zeros((3, lb.size))
        for p_ in range(3):
            alm[p_, sel] = np.einsum("nk,kn->n", M[p_, lb[sel], :, q], y[:, p_, jpre])
        wpix = 4 * np.pi / (12 * b["nside"] ** 2)
        mp = _Y(S, b["nside"], b["lmax"], alm) * wpix                     # WY with unit ring weights
        mp = np.where(b["invN"] > 0, mp / np.where(b["invN"] > 0, b["invN"], 1.0), 0.0)
        back = _Yt(S, b["nside"], b["lmax"], mp, job=S.YtW) * (alpha[q] ** 2)[:, None]
        for p_ in range(3):
            z[:, p_, jpre] += (M[p_, lb[sel], :, q] * back[p_, sel][:, None]).T
    Mp = M[:, lp][..., nb:]
    w2 = np.einsum("jakb,kja->bja", Mp, y)
    z += np.einsum("jakb,bja->kja", Mp, w2)
    return [_repack(z[i], lmax_pre, c["lmax"], nmaps_to=c["nmaps"]) for i, c in enumerate(comps)]


def test_pseudoinv_preconditioner_vs_oracle(shtlib, cpu_oracle):
    """precond_type 'pseudoinv' for the general system: one application against the same operator composed from oracle
    transforms (1e-9: two pinv implementations), symmetry, and PCG with it reaching the solution of the diagonal run."""
    import torch
    from commander_b200.comm_cr import solve_cr_eqn_by_CG
    S = cpu_oracle
    dev = torch.device("cuda", 0)
    bands, comps, rng = _problem(6, False)
    sysd = _build(bands, comps, dev)
    sysp = _build(bands, comps, dev)
    sysp.precond = "pseudoinv"
    rs = [rng.standard_normal((c["nmaps"], _lm(c["lmax"])[0].size)) for c in comps]
    r = torch.as_tensor(np.concatenate([v.ravel() for v in rs]), device=dev)
    got = sysp.invM(r).cpu().numpy()
    ref = np.concatenate([v.ravel() for v in _oracle_pseudoinv(S, bands, comps, rs)])
    assert rel(got, ref) <= 1e-9, rel(got, ref)
    r2 = torch.as_tensor(rng.standard_normal(sysp.ncr), device=dev)
    a, b = float(torch.dot(r2, sysp.invM(r))), float(torch.dot(r, sysp.invM(r2)))
    assert abs(a - b) <= 1e-10 * max(abs(a), abs(b))
    data = [torch.as_tensor(rng.standard_normal((3, 12 * b_["nside"] ** 2)) / np.sqrt(b_["invN"]), device=dev) for b_ in bands]
    rhs = sysd.computeRHS(data)
    xd, itd, _ = solve_cr_eqn_by_CG(sysd, rhs, maxiter=800, cg_tol=1e-12, cg_conv_crit="residual")
    xp, itp, _ = solve_cr_eqn_by_CG(sysp, rhs, maxiter=800, cg_tol=1e-12, cg_conv_crit="residual")
    assert float((xd - xp).norm() / xd.norm()) <= 1e-4, (itd, itp)
    assert itp < 800 and itd < 800
    print("CG iterations: diagonal", itd, "pseudoinv", itp)
