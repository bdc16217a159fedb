"""compute_invN_lm (commander3/src/comm_N_mod.f90:127-197): the oracle's Wigner-3j restatement against
independent references (CPU), and the GPU quadrature kernel `cmdr_sht_invn_diag` against the oracle and
closed forms.  Tolerance 1e-10 relative (FP64)."""
import math

import numpy as np
import pytest

from oracle import invn_lm


def test_wigner3j_vs_sympy():
    sympy = pytest.importorskip("sympy")
    from sympy.physics.wigner import wigner_3j
    rng = np.random.default_rng(0)
    for _ in range(60):
        l = int(rng.integers(0, 13)); L = int(rng.integers(0, 2 * l + 1)); m = int(rng.integers(0, l + 1))
        for args in ((L, l, l, 0, -m, m), (L, l, l, 0, 0, 0)):
            assert abs(invn_lm.wigner3j(*args) - float(wigner_3j(*args))) <= 1e-14


def test_gaunt_identity():
    """(-1)^m (2l+1) sqrt((2L+1)/4pi) (L l l;0 -m m)(L l l;0 0 0) = int |Y_lm|^2 Y_L0 (what the GPU kernel integrates)."""
    for l, m, L in ((3, 1, 2), (5, 5, 4), (7, 2, 10), (12, 7, 24), (12, 0, 6), (9, 4, 3)):
        lhs = (-1) ** m * (2 * l + 1) * math.sqrt((2 * L + 1) / (4 * math.pi)) * \
            invn_lm.wigner3j(L, l, l, 0, -m, m) * invn_lm.wigner3j(L, l, l, 0, 0, 0)
        assert abs(lhs - invn_lm.gaunt_quadrature(l, m, L)) <= 1e-12


def _closed_form(lmax, ms, npix, c0, c2):
    """N_lm for nbar = c0 Y_00 + c2 Y_20:  int|Y_lm|^2 Y_00 = 1/sqrt(4pi);
    int|Y_lm|^2 Y_20 = sqrt(5/4pi) (l(l+1) - 3m^2) / ((2l-1)(2l+3))."""
    out = []
    for m in ms:
        for l in range(m, lmax + 1):
            v = npix / (4 * math.pi) * (c0 / math.sqrt(4 * math.pi) +
                                        c2 * math.sqrt(5 / (4 * math.pi)) * (l * (l + 1) - 3 * m * m) / ((2 * l - 1) * (2 * l + 3)))
            out += [v] if m == 0 else [v, v]
    return np.array(out)


def test_oracle_closed_forms():
    lmax, ms, npix = 14, [0, 1, 4, 9, 14], 12 * 8 ** 2
    a = np.zeros((1, lmax + 1)); a[0, 0] = 1.3; a[0, 2] = -0.4
    got = invn_lm.compute_invN_lm(a, lmax, ms, npix)[0]
    ref = _closed_form(lmax, ms, npix, 1.3, -0.4)
    assert np.max(np.abs(got - ref)) <= 1e-11 * np.max(np.abs(ref))


@pytest.mark.gpu
@pytest.mark.parametrize("lmax,ms", [(8, None), (24, None), (37, [1, 4, 7, 10, 13, 16, 19, 22, 25, 28, 31, 34, 37]), (40, None)])
@pytest.mark.parametrize("nmaps", [1, 3])
def test_gpu_invN_diag_vs_oracle(shtlib, lmax, ms, nmaps):
    sharp = shtlib
    rng = np.random.default_rng(lmax + nmaps)
    msl = list(range(lmax + 1)) if ms is None else ms
    ai = sharp.sharp_make_mmajor_real_packed_alm_info(lmax, ms=None if ms is None else np.array(ms, dtype=np.int32))
    a_l0 = rng.standard_normal((nmaps, lmax + 1)) / (1.0 + np.arange(lmax + 1))
    a_l0[:, 0] += 5.0
    npix = 12 * 16 ** 2
    out = np.full((nmaps, ai.n_local), np.nan)
    sharp.invN_diag(a_l0, npix, ai, out)
    ref = invn_lm.compute_invN_lm(a_l0, lmax, msl, npix)
    assert out.shape == ref.shape
    err = np.linalg.norm(out - ref) / np.linalg.norm(ref)
    assert err <= 1e-10, err
    assert np.max(np.abs(out - ref)) <= 1e-10 * np.max(np.abs(ref))
    sharp.sharp_destroy_alm_info(ai)


@pytest.mark.gpu
def test_gpu_invN_diag_closed_form_large(shtlib):
    """lmax 2000 (the CG configuration), device output: monopole + quadrupole profile against the closed form."""
    import torch
    sharp = shtlib
    lmax, npix = 2000, 12 * 1024 ** 2
    ai = sharp.sharp_make_mmajor_real_packed_alm_info(lmax)
    a = np.zeros((2, lmax + 1)); a[0, 0] = 2.0; a[0, 2] = 0.7; a[1, 0] = 1.0
    out = torch.full((2, ai.n_local), float("nan"), dtype=torch.float64, device="cuda")
    sharp.invN_diag(a, npix, ai, out)
    got = out.cpu().numpy()
    ms = list(range(lmax + 1))
    ref0 = _closed_form(lmax, ms, npix, 2.0, 0.7)
    ref1 = _closed_form(lmax, ms, npix, 1.0, 0.0)
    assert np.max(np.abs(got[0] - ref0)) <= 1e-10 * np.max(np.abs(ref0))
    assert np.max(np.abs(got[1] - ref1)) <= 1e-10 * np.max(np.abs(ref1))
    sharp.sharp_destroy_alm_info(ai)
