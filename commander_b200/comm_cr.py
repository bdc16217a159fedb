"""Constrained-realisation CG for a CMB-only signal model: the mirror of the part of
`comm_cr_mod` that drives the SHT hot path (commander3/src/comm_cr_mod.f90).

    A = 1 + sqrt(S) B^T Y^T N^-1 Y B sqrt(S)            cr_matmulA, :771-1024 (one band, F = 1)
    b = sqrt(S) B^T Y^T N^-1 d  [+ sqrt(S) B^T Y^T N^-1/2 eta1 + eta0]   cr_computeRHS, :542-769
    PCG with M^-1 = diag                                  solve_cr_eqn_by_CG, :201-348; cr_invM :1026-1077

with N^-1 = siN^2 per pixel (commander3/src/comm_N_rms_mod.f90:264-273), B = b_l per l and
component (commander3/src/comm_B_bl_mod.f90:108-127), S = diagonal C_l (TT, EE, BB;
commander3/src/comm_Cl_mod.f90:588-637), and the diagonal preconditioner of
commander3/src/comm_diffuse_comp_mod.f90:2186-2235 for npre = 1.

All vectors live on the GPU as torch tensors of shape (nmaps, nalm) in the real-packed basis; the
two SHTs per A application are `comm_map%Y` and `comm_map%Yt` on device pointers; the two dot products
per iteration go through `mpi_dot_product` = local dot + NCCL all-reduce
(commander3/src/comm_utils.f90:599-614).

The preconditioner's N^-1_{lm,lm} is `compute_invN_lm` (commander3/src/comm_N_mod.f90:127-197), on the GPU
(`cmdr_sht_invn_diag`: the Wigner-3j sum of the reference evaluated as an exact quadrature).  `invN_lm="monopole"`
keeps only its L = 0 term, the sky mean of siN^2 times npix/4pi (what round 1 started with).  Any SPD
preconditioner gives the same solution; only the iteration count differs.
"""
from __future__ import annotations

import math

import numpy as np

from .comm_map import comm_map, comm_mapinfo


def gaussian_beam(lmax: int, fwhm_arcmin: float, nmaps: int = 3) -> np.ndarray:
    """b_l(l, j): Gaussian beam with the spin-2 factor of commander3/src/comm_map_mod.f90:1122-1135."""
    sigma = fwhm_arcmin * math.pi / 180.0 / 60.0 / math.sqrt(8.0 * math.log(2.0))
    l = np.arange(lmax + 1, dtype=np.float64)
    bl = np.empty((lmax + 1, nmaps))
    bl[:, 0] = np.exp(-0.5 * l * (l + 1) * sigma ** 2)
    for j in range(1, nmaps):
        bl[:, j] = np.exp(-0.5 * l * (l + 1) * sigma ** 2) * math.exp(2.0 * sigma ** 2)
    return bl


def compute_invN_lm(invN_diag: comm_map) -> None:
    """commander3/src/comm_N_mod.f90:127-197.  On entry invN_diag%map holds N^-1 per pixel (all columns treated
    as scalars, :146); on exit invN_diag%alm holds N_lm: `YtW_scalar`, the m = 0 coefficients from the rank that
    owns them to everybody (:147-152), then the 3j sum per local (l,m) (:154-184) -- here one GPU kernel."""
    from . import sharp
    info = invN_diag.info
    invN_diag.YtW_scalar()
    lmax, nm = info.lmax, info.nmaps
    i0 = info.lm2i(0, 0)
    if invN_diag.device is None:
        a_l0 = np.zeros((nm, lmax + 1))
        if i0 >= 0:
            a_l0[:] = invN_diag.alm[:, i0:i0 + lmax + 1]
        if info.dist and info.comm.size > 1:
            raise NotImplementedError("host-resident comm_map on several ranks: use device maps")
    else:
        import torch
        a = torch.zeros((nm, lmax + 1), dtype=torch.float64, device=invN_diag.alm.device)
        if i0 >= 0:
            a.copy_(invN_diag.alm[:, i0:i0 + lmax + 1])
        if info.dist and info.comm.size > 1:
            info.comm.allreduce_sum_(a)          # mpi_bcast from the owner of m = 0: zero everywhere else
        a_l0 = a.cpu().numpy()
    sharp.invN_diag(a_l0, float(info.npix), info.alm_info, invN_diag.alm)


class cr_cmb_system:
    """One band, one component (CMB, F = 1) constrained-realisation system."""

    def __init__(self, info: comm_mapinfo, siN, b_l: np.ndarray, Cl: np.ndarray, mask=None, mb_eff: float = 1.0,
                 precond: str = "diagonal", invN_lm: str = "wigner"):
        import torch
        self.torch = torch
        self.info = info
        dev = siN.device
        self.dev = dev
        self.siN = siN
        self.invN = siN * siN
        if mask is not None:
            self.invN = self.invN * mask      # samp_group mask, comm_N_rms_mod.f90:270-272
        l = info.lm[0].astype(np.int64)
        nm = info.nmaps
        # per-(alm slot, component) factors
        bl = np.stack([b_l[l, j] for j in range(nm)]) * mb_eff                 # matmulB
        sS = np.stack([np.sqrt(np.maximum(Cl[l, j], 0.0)) for j in range(nm)])  # sqrtS, diagonal C_l
        for j in range(1, nm):
            sS[j, l < 2] = 0.0
        self.bl = torch.as_tensor(bl, device=dev)
        self.sqrtS = torch.as_tensor(sS, device=dev)
        self.buf = comm_map(info, device=dev)
        # diagonal preconditioner: 1 + C_l b_l^2 N^-1_{lm,lm}, monopole term of compute_invN_lm
        mean_invN = self.invN.sum(dim=1)
        self._allreduce(mean_invN)
        mean_invN = mean_invN / float(info.npix)
        invN_diag = mean_invN * float(info.npix) / (4.0 * math.pi)
        if invN_lm == "wigner":
            # P_cr%invM_diff for npre = 1 with the full N_lm (comm_diffuse_comp_mod.f90:1167-1252)
            self.buf.map.copy_(self.invN)
            compute_invN_lm(self.buf)
            self.invN_lm = self.buf.alm.clone()
            self.Minv = 1.0 / (1.0 + (self.sqrtS * self.bl) ** 2 * self.invN_lm)
        elif invN_lm == "monopole":
            self.invN_lm = None
            self.Minv = 1.0 / (1.0 + (self.sqrtS * self.bl) ** 2 * invN_diag[:, None])
        else:
            raise ValueError("invN_lm must be 'wigner' or 'monopole'")
        # pseudo-inverse preconditioner (precond_type = 'pseudoinv')
        if precond not in ("diagonal", "pseudoinv"):
            raise ValueError("Preconditioner type not supported: " + precond)
        self.precond = precond
        # alpha_nu^2 = mean inverse noise variance in harmonic units (N%alpha_nu, :2293-2295): with it T / alpha^2 is
        # O(1), which the pseudo-inverse needs (V = [alpha U; 1])
        self.alpha2 = invN_diag[:, None]
        U = self.sqrtS * self.bl * torch.sqrt(self.alpha2)
        self.Uplus = U / (U * U + 1.0)
        self.Pplus2 = 1.0 / (U * U + 1.0) ** 2
        # N%N: sigma^2 per pixel; masked / unobserved pixels (N^-1 = 0) carry no information and are left out
        self.N = torch.where(self.invN > 0, 1.0 / self.invN.clamp_min(1e-300), torch.zeros_like(self.invN))
        self.n_matmul = 0

    # -- communicator helpers
    def _allreduce(self, t):
        c = self.info.comm
        if self.info.dist and c.size > 1:
            c.allreduce_sum_(t)
        return t

    def mpi_dot_product(self, a, b):
        """commander3/src/comm_utils.f90:599-614"""
        s = (a * b).sum().reshape(1)
        self._allreduce(s)
        return float(s.item())

    # -- operators
    def matmulA(self, x):
        """cr_matmulA, commander3/src/comm_cr_mod.f90:771-1024 for one band / one diffuse component."""
        m = self.buf
        m.alm.copy_(x)
        m.alm.mul_(self.sqrtS)        # sqrtS_x, :797-836
        m.alm.mul_(self.bl)           # evalDiffuseBand: F_mean = 1, beam :2089
        m.Y()                         # :888-892
        m.map.mul_(self.invN)         # N%invN, :905
        m.Yt()                        # :913-918
        m.alm.mul_(self.bl)           # projectDiffuseBand (B^T)
        m.alm.mul_(self.sqrtS)        # :957-1008
        self.n_matmul += 1
        return x + m.alm

    def invM(self, r):
        """cr_invM, :1026-1077: 'diagonal' (default) or 'pseudoinv' preconditioner."""
        if self.precond == "pseudoinv":
            return self.invM_pseudoinv(r)
        return r * self.Minv

    def invM_pseudoinv(self, r):
        """applyDiffPrecond_pseudoinv, commander3/src/comm_diffuse_comp_mod.f90:2237-2380, for one band and one
        component.  The system is A = V^T diag(T, 1) V with V = [U; 1], U = alpha b_l sqrt(C_l) per (l, pol) and
        T = Y^T N^-1 Y / alpha^2; its pseudo-inverse preconditioner is V^+ diag(T^+, 1) V^+T with V^+ = [U, 1] / (U^2 + 1)
        (the columns of P_cr%invM_diff, :1313-1557) and T^+ = alpha^2 YtW N WY (:2288-2292: `call invN_x%WY`,
        `call data(k)%N%N(invN_x)`, `call invN_x%YtW`)."""
        m = self.buf
        m.alm.copy_(r)
        m.alm.mul_(self.Uplus)                 # sum over (U^plus)^t, :2270-2285
        m.WY()                                 # :2288
        m.map.mul_(self.N)                     # N%N: noise covariance, :2290
        m.YtW()                                # :2292
        m.alm.mul_(self.alpha2)                # alpha_nu^2, :2293-2295
        z = m.alm * self.Uplus                 # sum over U^plus, :2300-2315
        return z + r * self.Pplus2             # prior terms, :2322-2346

    def computeRHS(self, data, eta_pix=None, eta_alm=None):
        """cr_computeRHS, :542-769: mean-field term, plus the two fluctuation terms when the white
        noise draws are given ('sample' branch)."""
        m = self.buf
        m.map.copy_(data * self.invN)
        if eta_pix is not None:
            m.map.add_(self.siN * eta_pix)          # N^-1/2 eta
        m.Yt()                                       # :615
        m.alm.mul_(self.bl)                          # B^T, :616
        m.alm.mul_(self.sqrtS)                       # :652
        b = m.alm.clone()
        if eta_alm is not None:
            b += eta_alm
        return b

    def x2amp(self, x):
        """cr_x2amp after multiplying with sqrt(S) (comm_cr_mod.f90:350-390): the signal a_lm."""
        return x * self.sqrtS


def solve_cr_eqn_by_CG(sys: cr_cmb_system, b, maxiter=300, cg_tol=1e-8, cg_conv_crit="residual",
                       cg_miniter=5, cg_check_conv_freq=1, x0=None, verbose=False):
    """solve_cr_eqn_by_CG, commander3/src/comm_cr_mod.f90:201-348 (same update order, same
    convergence test).  Returns (x, iterations done, residual history)."""
    torch = sys.torch
    x = torch.zeros_like(b) if x0 is None else x0.clone()
    r = b - sys.matmulA(x)
    d = sys.invM(r)
    delta_new = sys.mpi_dot_product(r, d)
    delta0 = sys.mpi_dot_product(b, sys.invM(b))
    if cg_conv_crit not in ("residual", "fixed_iter"):
        raise ValueError("Unsupported convergence criterion = " + cg_conv_crit)
    lim_convergence = cg_tol * delta0
    val_convergence = 1e2 * lim_convergence
    hist = [delta_new]
    it = 0
    for i in range(1, maxiter + 1):
        if i % cg_check_conv_freq == 0:
            val_convergence = delta_new
            if (val_convergence < lim_convergence and (i >= cg_miniter or delta_new <= 1e-30 * delta0)
                    and cg_conv_crit != "fixed_iter"):
                break
        q = sys.matmulA(d)
        alpha = delta_new / sys.mpi_dot_product(d, q)
        x = x + alpha * d
        r = r - alpha * q
        s = sys.invM(r)
        delta_old = delta_new
        delta_new = sys.mpi_dot_product(r, s)
        beta = delta_new / delta_old
        d = s + beta * d
        hist.append(delta_new)
        it = i
        if verbose:
            print(f"  CG iter. {i:5d} -- res = {delta_new:13.5e}, tol = {lim_convergence:13.5e}")
    return x, it, hist
