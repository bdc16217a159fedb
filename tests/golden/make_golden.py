"""Generates tests/golden/*.npz from the DEFINITIONAL oracle (oracle/sht_def.py: dense Y
matrices from 50-digit Wigner-d sums, independent of the fast CPU path and of the CUDA
product) plus scipy.special.sph_harm_y columns as a third-party pin.

The reference ships no golden vectors for this path and its implementation
(libsharp2) is not available here, so these are the committed fixtures.  Re-run with
    python tests/golden/make_golden.py
"""
import math
import os
import sys

import numpy as np
import scipy.special as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import sht_def as D  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def case(nside, lmax, seed, rank=0, nprocs=1):
    rng = np.random.default_rng(seed)
    rings = D.mapinfo_rings(nside, rank, nprocs)
    ms = D.mapinfo_ms(lmax, rank, nprocs)
    w = rng.uniform(0.9, 1.1, 2 * nside)
    out = {"nside": nside, "lmax": lmax, "rings": np.array(rings, dtype=np.int32),
           "ms": np.array(ms, dtype=np.int32), "weight": w}
    for spin in (0, 2):
        nc = 1 if spin == 0 else 2
        Y = D.Y_matrix(nside, lmax, spin, rings=rings, ms=ms)
        nalm, npix = D.alm_count(lmax, ms), D.map_size(nside, rings)
        alm = rng.standard_normal(nc * nalm)
        mp = rng.standard_normal(nc * npix)
        wv = D.ring_weights_vector(nside, rings, w, nc)
        out[f"s{spin}_alm"] = alm.reshape(nc, nalm)
        out[f"s{spin}_map"] = mp.reshape(nc, npix)
        out[f"s{spin}_Y"] = (Y @ alm).reshape(nc, npix)
        out[f"s{spin}_WY"] = (wv * (Y @ alm)).reshape(nc, npix)
        out[f"s{spin}_Yt"] = (Y.T @ mp).reshape(nc, nalm)
        out[f"s{spin}_YtW"] = (Y.T @ (wv * mp)).reshape(nc, nalm)
    return out


def scipy_columns(nside, lmax):
    """T map of single unit real-packed coefficients, straight from scipy."""
    idx = D.alm_index(lmax, range(lmax + 1))
    picks = [(0, 0), (1, 0), (1, 1), (1, -1), (2, 2), (3, -2), (lmax, lmax), (lmax, -1), (lmax, 0)]
    cols = []
    for (l, m) in picks:
        col = []
        for r in range(1, 4 * nside):
            cth, sth, nph, phi0, _ = D.healpix_ring(nside, r)
            phi = phi0 + 2 * np.pi * np.arange(nph) / nph
            y = sp.sph_harm_y(l, abs(m), math.atan2(sth, cth), phi)
            col.append(y.real if m == 0 else (math.sqrt(2) * y.real if m > 0 else -math.sqrt(2) * y.imag))
        cols.append(np.concatenate(col))
    return {"nside": nside, "lmax": lmax, "picks": np.array(picks, dtype=np.int32),
            "slots": np.array([idx.index(p) for p in picks], dtype=np.int32), "cols": np.array(cols)}


def conviqt_case(nside, lmax, bmax, seed):
    """The conviqt cube (commander3/src/comm_conviqt_mod.f90:207-357) with the spin-j syntheses done by the DENSE
    definitional matrices (not by the fast CPU path): sky and beam a_lm in, double-precision cube out."""
    from oracle import conviqt as O
    rng = np.random.default_rng(seed)
    lm = O.lm_table(lmax)
    sky = rng.standard_normal((3, len(lm)))
    beam = rng.standard_normal((3, len(lm))) / (1.0 + np.array([t[0] for t in lm]))
    tab = O.beam_table(lmax, 3, lm, beam)
    npix = D.map_size(nside, list(range(1, 4 * nside)))
    marr = {}
    for j in range(bmax + 1):
        alm = O.get_alms(j, lmax, lm, sky, tab)
        if j == 0:
            marr[0] = D.Y_matrix(nside, lmax, 0) @ alm[0]
        else:
            mm = (D.Y_matrix(nside, lmax, j) @ alm.ravel()).reshape(2, npix)
            marr[j], marr[-j] = mm[0], mm[1]
    n = 2 * bmax
    cube = np.empty((n, npix))
    for i in range(npix):
        dv = np.array([marr[0][i]] + [marr[j][i] + 1j * marr[-j][i] for j in range(1, bmax + 1)])
        cube[:, i] = O.c2r(dv, n)
    return {"nside": nside, "lmax": lmax, "bmax": bmax, "sky_alm": sky, "beam_alm": beam, "cube": cube}


if __name__ == "__main__":
    np.savez_compressed(os.path.join(OUT, "conviqt_n2_l5_b2.npz"), **conviqt_case(2, 5, 2, 21))
    np.savez_compressed(os.path.join(OUT, "conviqt_n4_l6_b3.npz"), **conviqt_case(4, 6, 3, 22))
    np.savez_compressed(os.path.join(OUT, "dense_n4_l9.npz"), **case(4, 9, 11))
    np.savez_compressed(os.path.join(OUT, "dense_n2_l7.npz"), **case(2, 7, 12))        # lmax > 3 nside - 1
    np.savez_compressed(os.path.join(OUT, "dense_n4_l8_rank1of3.npz"), **case(4, 8, 13, rank=1, nprocs=3))
    np.savez_compressed(os.path.join(OUT, "scipy_cols_n8_l12.npz"), **scipy_columns(8, 12))
    print("golden vectors written to", OUT)
