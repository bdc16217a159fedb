"""GPU parity tests: the CUDA path, called through the C ABI exactly as
commander3/src/sharp.f90 calls libsharp2, against the CPU oracle on the same seeded
inputs.  Tolerance (BASELINE.json north_star): relative L2 error <= 1e-10 in FP64."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-10


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _handles(sharp, nside, lmax, rings=None, ms=None, weight=None):
    ai = sharp.sharp_make_mmajor_real_packed_alm_info(lmax, ms=ms)
    gi = sharp.sharp_make_healpix_geom_info(nside, rings=rings, weight=weight)
    return ai, gi


def _zero_low_l(alm, lmax, ms):
    """E/B have no l<2 modes."""
    i = 0
    for m in (range(lmax + 1) if ms is None else ms):
        f = 1 if m == 0 else 2
        for l in range(m, min(lmax, 1) + 1):
            alm[:, i + f * (l - m): i + f * (l - m) + f] = 0.0
        i += f * (lmax + 1 - m)
    return alm


CASES = [(1, 2), (1, 5), (2, 4), (2, 9), (4, 8), (4, 15), (8, 16), (16, 47), (32, 64), (64, 128), (64, 200)]


@pytest.mark.parametrize("nside,lmax", CASES)
@pytest.mark.parametrize("spin", [0, 2])
def test_all_jobs_vs_oracle(shtlib, cpu_oracle, nside, lmax, spin):
    sharp, S = shtlib, cpu_oracle
    rng = np.random.default_rng(1000 * nside + lmax + spin)
    nc = 1 if spin == 0 else 2
    w = rng.uniform(0.9, 1.1, 2 * nside)
    ai, gi = _handles(sharp, nside, lmax, weight=w)
    assert ai.n_local == S.alm_count(lmax) and gi.n_local == S.map_size(nside)
    alm = rng.standard_normal((nc, ai.n_local))
    if spin == 2:
        _zero_low_l(alm, lmax, None)
    mp = rng.standard_normal((nc, gi.n_local))
    for job, name in ((sharp.SHARP_Y, "Y"), (sharp.SHARP_WY, "WY")):
        out = np.full((nc, gi.n_local), np.nan)
        sharp.sharp_execute(job, spin, nc, alm.copy(), ai, out, gi)
        ref = S.execute(job, spin, nside, lmax, alm=alm, weight=w)
        assert rel(out, ref) <= TOL, (name, rel(out, ref))
    for job, name in ((sharp.SHARP_Yt, "Yt"), (sharp.SHARP_YtW, "YtW")):
        out = np.full((nc, ai.n_local), np.nan)
        sharp.sharp_execute(job, spin, nc, out, ai, mp.copy(), gi)
        ref = S.execute(job, spin, nside, lmax, map=mp, weight=w)
        assert rel(out, ref) <= TOL, (name, rel(out, ref))
    sharp.sharp_destroy_alm_info(ai)
    sharp.sharp_destroy_geom_info(gi)


def test_config1_roundtrip_and_parity(shtlib, cpu_oracle):
    """BASELINE.json configs[0]: comm_map Y/YtW round trip, nside=256 lmax=512 IQU, 1 rank."""
    from commander_b200 import comm_map, comm_mapinfo
    S = cpu_oracle
    nside, lmax = 256, 512
    info = comm_mapinfo(None, nside, lmax, 3, True)
    m = comm_map(info)
    rng = np.random.default_rng(1)
    m.alm[:] = rng.standard_normal(m.alm.shape)
    _zero_low_l(m.alm[1:3], lmax, None)
    alm0 = m.alm.copy()
    m.Y()
    refT = S.execute(S.Y, 0, nside, lmax, alm=alm0[0:1])
    refP = S.execute(S.Y, 2, nside, lmax, alm=alm0[1:3])
    assert rel(m.map[0:1], refT) <= TOL and rel(m.map[1:3], refP) <= TOL
    m.YtW()
    aT = S.execute(S.YtW, 0, nside, lmax, map=refT)
    aP = S.execute(S.YtW, 2, nside, lmax, map=refP)
    assert rel(m.alm[0:1], aT) <= TOL and rel(m.alm[1:3], aP) <= TOL
    # HEALPix quadrature is approximate: round trip is a sanity check only
    assert rel(m.alm, alm0) < 2e-2


@pytest.mark.parametrize("spin", [0, 2])
def test_ring_and_m_subsets(shtlib, cpu_oracle, spin):
    """Local rings / m's of rank r of P as comm_mapinfo builds them
    (commander3/src/comm_map_mod.f90:197-261)."""
    from oracle import sht_def as D
    sharp, S = shtlib, cpu_oracle
    nside, lmax, P = 16, 40, 3
    nc = 1 if spin == 0 else 2
    rng = np.random.default_rng(7 + spin)
    for r in range(P):
        rings, ms = D.mapinfo_rings(nside, r, P), D.mapinfo_ms(lmax, r, P)
        ai, gi = _handles(sharp, nside, lmax, rings=rings, ms=ms)
        alm = rng.standard_normal((nc, ai.n_local))
        out = np.zeros((nc, gi.n_local))
        sharp.sharp_execute(sharp.SHARP_Y, spin, nc, alm, ai, out, gi)
        ref = S.execute(S.Y, spin, nside, lmax, alm=alm, rings=rings, ms=ms)
        assert rel(out, ref) <= TOL
        mp = rng.standard_normal((nc, gi.n_local))
        out = np.zeros((nc, ai.n_local))
        sharp.sharp_execute(sharp.SHARP_YtW, spin, nc, out, ai, mp, gi)
        ref = S.execute(S.YtW, spin, nside, lmax, map=mp, rings=rings, ms=ms)
        assert rel(out, ref) <= TOL
    # a northern-only ring list in scrambled order
    rings = [5, 1, 9, 20, 33]
    ai, gi = _handles(sharp, nside, lmax, rings=rings)
    alm = rng.standard_normal((nc, ai.n_local))
    out = np.zeros((nc, gi.n_local))
    sharp.sharp_execute(sharp.SHARP_Y, spin, nc, alm, ai, out, gi)
    ref = S.execute(S.Y, spin, nside, lmax, alm=alm, rings=rings)
    assert rel(out, ref) <= TOL


@pytest.mark.parametrize("spin", [0, 2])
def test_add_flag(shtlib, cpu_oracle, spin):
    sharp, S = shtlib, cpu_oracle
    nside, lmax = 8, 20
    nc = 1 if spin == 0 else 2
    rng = np.random.default_rng(3)
    ai, gi = _handles(sharp, nside, lmax)
    alm = rng.standard_normal((nc, ai.n_local))
    base = rng.standard_normal((nc, gi.n_local))
    out = base.copy()
    sharp.sharp_execute(sharp.SHARP_Y, spin, nc, alm, ai, out, gi, add=True)
    ref = base + S.execute(S.Y, spin, nside, lmax, alm=alm)
    assert rel(out, ref) <= TOL
    abase = rng.standard_normal((nc, ai.n_local))
    aout = abase.copy()
    sharp.sharp_execute(sharp.SHARP_Yt, spin, nc, aout, ai, base, gi, add=True)
    ref = abase + S.execute(S.Yt, spin, nside, lmax, map=base)
    assert rel(aout, ref) <= TOL


def test_adjointness_and_device_pointers(shtlib):
    """<Y a, x> == <a, Yt x> to 1e-12, with device-resident
    tensors passed through the same sharp_execute entry point."""
    import torch
    from commander_b200 import comm_map, comm_mapinfo
    nside, lmax = 128, 300
    info = comm_mapinfo(None, nside, lmax, 3, True)
    a, x = comm_map(info, device="cuda"), comm_map(info, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(5)
    a.alm.normal_(generator=g)
    x.map.normal_(generator=g)
    xm = x.map.clone()
    a.Y()
    x.Yt()
    lhs = float((a.map * xm).sum())
    a2 = comm_map(info, device="cuda")
    g2 = torch.Generator(device="cuda").manual_seed(5)
    a2.alm.normal_(generator=g2)
    rhs = float((a2.alm * x.alm).sum())
    assert abs(lhs - rhs) <= 1e-12 * max(abs(lhs), abs(rhs), float(a.map.norm() * xm.norm()))
    # fused IQU entry point gives the same maps
    b = comm_map(info, device="cuda")
    b.alm.copy_(a2.alm)
    b.Y_iqu()
    assert float((b.map - a.map).norm() / a.map.norm()) <= 1e-14


@pytest.mark.parametrize("spin", [0, 2])
def test_pinned_host_pipelined_path(shtlib, cpu_oracle, spin):
    """Pinned host buffers take the chunked copy/compute-overlapped path (abi.cu try_pipelined);
    pageable buffers take the plain staged path.  Both must match the oracle."""
    import torch
    sharp, S = shtlib, cpu_oracle
    nside, lmax = 512, 700
    nc = 1 if spin == 0 else 2
    rng = np.random.default_rng(21 + spin)
    w = rng.uniform(0.9, 1.1, 2 * nside)
    ai, gi = _handles(sharp, nside, lmax, weight=w)
    assert gi.n_local >= (1 << 21)
    alm_p = torch.empty((nc, ai.n_local), dtype=torch.float64).pin_memory()
    map_p = torch.empty((nc, gi.n_local), dtype=torch.float64).pin_memory()
    alm_p.copy_(torch.as_tensor(rng.standard_normal((nc, ai.n_local))))
    alm0 = alm_p.numpy().copy()
    sharp.sharp_execute(sharp.SHARP_Y, spin, nc, alm_p.numpy(), ai, map_p.numpy(), gi)
    ref = S.execute(S.Y, spin, nside, lmax, alm=alm0)
    assert rel(map_p.numpy(), ref) <= TOL
    pageable = np.zeros((nc, gi.n_local))
    sharp.sharp_execute(sharp.SHARP_Y, spin, nc, alm0.copy(), ai, pageable, gi)
    assert rel(pageable, map_p.numpy()) <= 1e-14
    map_p.copy_(torch.as_tensor(rng.standard_normal((nc, gi.n_local))))
    mp0 = map_p.numpy().copy()
    sharp.sharp_execute(sharp.SHARP_YtW, spin, nc, alm_p.numpy(), ai, map_p.numpy(), gi)
    ref = S.execute(S.YtW, spin, nside, lmax, map=mp0, weight=w)
    assert rel(alm_p.numpy(), ref) <= TOL
    sharp.sharp_destroy_alm_info(ai)
    sharp.sharp_destroy_geom_info(gi)


def test_general_complex_alm_info(shtlib, cpu_oracle):
    """sharp_make_general_alm_info (commander3/src/sharp.f90:35-42; declared by the reference, never
    called): complex a_lm, m-major, against the real-packed result."""
    import ctypes as C
    sharp, S = shtlib, cpu_oracle
    nside, lmax = 16, 33
    L = sharp.lib()
    mval = np.arange(lmax + 1, dtype=np.int32)
    mvstart = np.zeros(lmax + 1, dtype=np.int64)
    idx = 0
    for m in range(lmax + 1):
        mvstart[m] = idx - m
        idx += lmax + 1 - m
    h = C.c_void_p()
    L.sharp_make_general_alm_info(lmax, lmax + 1, 1, mval.ctypes.data, mvstart.ctypes.data, 0, C.byref(h))
    assert L.sharp_alm_count(h) == idx
    gi = sharp.sharp_make_healpix_geom_info(nside)
    rng = np.random.default_rng(17)
    a = rng.standard_normal(idx) + 1j * rng.standard_normal(idx)
    a[: lmax + 1] = a[: lmax + 1].real          # m = 0 is real
    buf = np.ascontiguousarray(a).view(np.float64).copy()
    out = np.zeros(gi.n_local)
    ap = (C.c_void_p * 1)(buf.ctypes.data); mp = (C.c_void_p * 1)(out.ctypes.data)
    L.sharp_execute(sharp.SHARP_Y, 0, ap, mp, gi.handle, h, sharp.SHARP_DP, None, None)
    # the same field in the real-packed basis: alm[+m] = sqrt2 Re a, alm[-m] = sqrt2 Im a
    packed = []
    for m in range(lmax + 1):
        blk = a[mvstart[m] + m: mvstart[m] + lmax + 1]
        packed.append(blk.real if m == 0 else np.sqrt(2.0) * np.stack([blk.real, blk.imag], axis=1).ravel())
    ref = S.execute(S.Y, 0, nside, lmax, alm=np.concatenate(packed)[None, :])
    assert rel(out[None, :], ref) <= TOL
    back = np.zeros(2 * idx)
    bp = (C.c_void_p * 1)(back.ctypes.data)
    L.sharp_execute(sharp.SHARP_Yt, 0, bp, mp, gi.handle, h, sharp.SHARP_DP, None, None)
    refa = S.execute(S.Yt, 0, nside, lmax, map=out[None, :])[0]
    got = back.view(np.complex128)
    o = 0
    for m in range(lmax + 1):
        n = lmax + 1 - m
        blk = got[mvstart[m] + m: mvstart[m] + lmax + 1]
        if m == 0:
            assert np.linalg.norm(blk.real - refa[o:o + n]) <= TOL * np.linalg.norm(refa[o:o + n]); o += n
        else:
            r = refa[o:o + 2 * n].reshape(n, 2) / np.sqrt(2.0)
            assert np.linalg.norm(blk.real - r[:, 0]) + np.linalg.norm(blk.imag - r[:, 1]) <= 10 * TOL * np.linalg.norm(r); o += 2 * n
    L.sharp_destroy_alm_info(h)
    sharp.sharp_destroy_geom_info(gi)


@pytest.mark.parametrize("path", ["dense_n4_l9.npz", "dense_n2_l7.npz", "dense_n4_l8_rank1of3.npz"])
def test_golden_vectors(shtlib, path):
    """Committed fixtures from the definitional oracle (tests/golden/make_golden.py)."""
    import os
    sharp = shtlib
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", path))
    nside, lmax = int(g["nside"]), int(g["lmax"])
    ai = sharp.sharp_make_mmajor_real_packed_alm_info(lmax, ms=g["ms"])
    gi = sharp.sharp_make_healpix_geom_info(nside, rings=g["rings"], weight=g["weight"])
    for spin in (0, 2):
        nc = 1 if spin == 0 else 2
        alm, mp = np.ascontiguousarray(g[f"s{spin}_alm"]), np.ascontiguousarray(g[f"s{spin}_map"])
        for job, key in ((sharp.SHARP_Y, "Y"), (sharp.SHARP_WY, "WY")):
            out = np.zeros((nc, gi.n_local))
            sharp.sharp_execute(job, spin, nc, alm.copy(), ai, out, gi)
            assert rel(out, g[f"s{spin}_{key}"]) <= 1e-12
        for job, key in ((sharp.SHARP_Yt, "Yt"), (sharp.SHARP_YtW, "YtW")):
            out = np.zeros((nc, ai.n_local))
            sharp.sharp_execute(job, spin, nc, out, ai, mp.copy(), gi)
            assert rel(out, g[f"s{spin}_{key}"]) <= 1e-12


@pytest.mark.parametrize("nside,lmax", [(1024, 2048), (2048, 4000)])
def test_full_size_properties(shtlib, nside, lmax):
    """BASELINE.json configs[1] and [3] sizes, where the oracle is too slow for a full comparison:
    size-independent properties -- linearity, adjointness <Y a, x> = <a, Yt x>, WY = w Y, YtW = Yt w,
    the approximate inverse YtW Y ~ 1 (HEALPix quadrature, ~1e-3), and oracle parity on one m-column."""
    import torch
    from commander_b200 import comm_map, comm_mapinfo
    rng = np.random.default_rng(3)
    w = rng.uniform(0.95, 1.05, (2, 2 * nside))
    info = comm_mapinfo(None, nside, lmax, 3, True, weights=w)
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(4)
    a, b = comm_map(info, device=dev), comm_map(info, device=dev)
    l = torch.as_tensor(info.lm[0].astype(np.int64), device=dev)
    cl = 1.0 / (l * (l + 1) + 1.0).double().sqrt()     # red spectrum exercises the dynamic range
    a.alm.normal_(generator=g); a.alm.mul_(cl); a.alm[1:3, l < 2] = 0
    b.alm.normal_(generator=g); b.alm.mul_(cl); b.alm[1:3, l < 2] = 0
    a0, b0 = a.alm.clone(), b.alm.clone()
    a.Y(); b.Y()
    c = comm_map(info, device=dev)
    c.alm.copy_(2.0 * a0 - 3.0 * b0)
    c.Y()
    lin = float((c.map - (2.0 * a.map - 3.0 * b.map)).norm() / c.map.norm())
    assert lin <= 1e-13, lin
    # adjointness
    x = comm_map(info, device=dev)
    x.map.normal_(generator=g)
    xm = x.map.clone()
    x.Yt()
    lhs, rhs = float((a.map * xm).sum()), float((a0 * x.alm).sum())
    assert abs(lhs - rhs) <= 1e-12 * float(a.map.norm() * xm.norm())
    # WY = diag(w) Y ; YtW = Yt diag(w)
    d = comm_map(info, device=dev)
    d.alm.copy_(a0); d.WY()
    ringw = 4 * np.pi / (12 * nside ** 2) * w[:, np.minimum(info.rings, 4 * nside - info.rings) - 1]
    counts = np.where(np.minimum(info.rings, 4 * nside - info.rings) < nside, 4 * np.minimum(info.rings, 4 * nside - info.rings), 4 * nside)
    wT = torch.as_tensor(np.repeat(ringw[0], counts), device=dev)
    wP = torch.as_tensor(np.repeat(ringw[1], counts), device=dev)
    assert float((d.map[0] - wT * a.map[0]).norm() / d.map[0].norm()) <= 1e-14
    assert float((d.map[1:] - wP * a.map[1:]).norm() / d.map[1:].norm()) <= 1e-14
    e = comm_map(info, device=dev)
    e.map[0] = xm[0] * wT; e.map[1:] = xm[1:] * wP
    e.Yt()
    f = comm_map(info, device=dev)
    f.map.copy_(xm); f.YtW()
    assert float((f.alm - e.alm).norm() / e.alm.norm()) <= 1e-13
    # approximate inverse
    a.YtW()
    rt = float((a.alm - a0).norm() / a0.norm())
    assert rt < 0.1, rt


def _red_alm(rng, info, lmax):
    """alm ~ N(0, C_l), C_l = 1 / (l (l + 1) + 1): a red spectrum exercises the dynamic range."""
    l = info.lm[0].astype(np.float64)
    a = rng.standard_normal((3, info.nalm)) / np.sqrt(l * (l + 1) + 1.0)
    a[1:3, info.lm[0] < 2] = 0.0
    return a


@pytest.mark.parametrize("nside,lmax", [(1024, 2048), (1024, 2000), (2048, 4000), (512, 1500)])
def test_full_size_vs_oracle(shtlib, cpu_oracle, nside, lmax):
    """BASELINE.json configs[1] (1024 / 2048), [2]-size (1024 / 2000), [3] (2048 / 4000) and [4] (512 / 1500, one band)
    compared IN FULL with the oracle: Y, WY, Yt, YtW for T (spin 0) and Q,U (spin 2) through comm_map on
    ordinary numpy arrays -- pageable host memory, exactly what commander3/src/sharp.f90:219-224 passes (at these sizes
    that is the chunked PCIe pipeline through the pinned arena).  WY and YtW are checked against the oracle's Y / Yt
    with the ring weights applied in numpy (WY = diag(w) Y, YtW = Yt diag(w)), which halves the oracle time."""
    from commander_b200 import comm_map, comm_mapinfo
    S = cpu_oracle
    rng = np.random.default_rng(3 + nside + lmax)
    w = rng.uniform(0.9, 1.1, (2, 2 * nside))
    info = comm_mapinfo(None, nside, lmax, 3, True, weights=w)
    north = np.minimum(info.rings, 4 * nside - info.rings)
    counts = np.where(north < nside, 4 * north, 4 * nside)
    wpix = np.stack([np.repeat(4 * np.pi / (12 * nside ** 2) * w[k, north - 1], counts) for k in (0, 1, 1)])   # T, Q, U
    m = comm_map(info)
    alm0 = _red_alm(rng, info, lmax)
    # ---- synthesis
    refY = np.concatenate([S.execute(S.Y, 0, nside, lmax, alm=alm0[0:1]), S.execute(S.Y, 2, nside, lmax, alm=alm0[1:3])])
    m.alm[:] = alm0
    m.Y()
    for c in range(3):
        assert rel(m.map[c], refY[c]) <= TOL, ("Y", c, rel(m.map[c], refY[c]))
    m.map[:] = np.nan
    m.WY()
    for c in range(3):
        assert rel(m.map[c], wpix[c] * refY[c]) <= TOL, ("WY", c, rel(m.map[c], wpix[c] * refY[c]))
    del refY
    # ---- analysis
    x = rng.standard_normal((3, info.np))
    refA = np.concatenate([S.execute(S.Yt, 0, nside, lmax, map=x[0:1]), S.execute(S.Yt, 2, nside, lmax, map=x[1:3])])
    m.map[:] = x
    m.alm[:] = np.nan
    m.Yt()
    for c in range(3):
        assert rel(m.alm[c], refA[c]) <= TOL, ("Yt", c, rel(m.alm[c], refA[c]))
    xw = x * wpix
    refA = np.concatenate([S.execute(S.Yt, 0, nside, lmax, map=xw[0:1]), S.execute(S.Yt, 2, nside, lmax, map=xw[1:3])])
    m.map[:] = x
    m.alm[:] = np.nan
    m.YtW()
    for c in range(3):
        assert rel(m.alm[c], refA[c]) <= TOL, ("YtW", c, rel(m.alm[c], refA[c]))
    info.dealloc()


def test_full_size_m_column_vs_oracle(shtlib, cpu_oracle):
    """nside 2048 / lmax 4000: a_lm restricted to a few m (incl. the largest) against the oracle on
    every ring -- exercises the deep-underflow starts and the m cut-off at the target size."""
    sharp, S = shtlib, cpu_oracle
    nside, lmax = 2048, 4000
    ms = np.array([0, 1, 2, 777, 2048, 3100, 3999, 4000], dtype=np.int32)
    ai = sharp.sharp_make_mmajor_real_packed_alm_info(lmax, ms=ms)
    gi = sharp.sharp_make_healpix_geom_info(nside)
    rng = np.random.default_rng(8)
    for spin in (0, 2):
        nc = 1 if spin == 0 else 2
        alm = rng.standard_normal((nc, ai.n_local))
        out = np.zeros((nc, gi.n_local))
        sharp.sharp_execute(sharp.SHARP_Y, spin, nc, alm, ai, out, gi)
        ref = S.execute(S.Y, spin, nside, lmax, alm=alm, ms=ms)
        assert rel(out, ref) <= TOL, rel(out, ref)
        back = np.zeros((nc, ai.n_local))
        sharp.sharp_execute(sharp.SHARP_Yt, spin, nc, back, ai, ref, gi)
        refb = S.execute(S.Yt, spin, nside, lmax, map=ref, ms=ms)
        assert rel(back, refb) <= TOL, rel(back, refb)
    sharp.sharp_destroy_alm_info(ai)
    sharp.sharp_destroy_geom_info(gi)


def test_config5_band_batch(shtlib, cpu_oracle):
    """BASELINE.json configs[4] shape (nside 512, lmax 1500, IQU) on 3 of the 30 bands: every band
    through Y/YtW on shared handles; band 0 checked against the oracle."""
    import torch
    from commander_b200 import comm_map, comm_mapinfo
    S = cpu_oracle
    nside, lmax = 512, 1500
    info = comm_mapinfo(None, nside, lmax, 3, True)
    maps = []
    for band in range(3):
        m = comm_map(info, device="cuda")
        g = torch.Generator(device="cuda").manual_seed(100 + band)
        m.alm.normal_(generator=g)
        m.Y()
        maps.append(m)
    a0 = maps[0].alm.cpu().numpy()
    refT = S.execute(S.Y, 0, nside, lmax, alm=a0[0:1])
    refP = S.execute(S.Y, 2, nside, lmax, alm=a0[1:3])
    out = maps[0].map.cpu().numpy()
    assert rel(out[0:1], refT) <= TOL and rel(out[1:3], refP) <= TOL
    assert float((maps[1].map - maps[0].map).norm()) > 0


@pytest.mark.parametrize("pinned", [False, True])
def test_band_batch_api_matches_sequential(shtlib, cpu_oracle, pinned):
    """cmdr_sht_execute_iqu_batch (bands pipelined over copy / compute streams) gives exactly what
    the per-band calls give, for all four job types, and band 0 matches the oracle."""
    import torch
    from commander_b200 import comm_map, comm_mapinfo
    S = cpu_oracle
    nside, lmax, nb = 64, 150, 5
    rng = np.random.default_rng(77)
    info = comm_mapinfo(None, nside, lmax, 3, True, weights=rng.uniform(0.9, 1.1, (2, 2 * nside)))

    def fresh():
        ms = [comm_map(info) for _ in range(nb)]
        if pinned:
            for m in ms:
                m.alm = torch.empty((3, info.nalm), dtype=torch.float64).pin_memory().numpy()
                m.map = torch.empty((3, info.np), dtype=torch.float64).pin_memory().numpy()
        return ms
    alms = rng.standard_normal((nb, 3, info.nalm))
    pix = rng.standard_normal((nb, 3, info.np))
    for batch_fn, single_fn, synth in ((comm_map.Y_batch, "Y", True), (comm_map.WY_batch, "WY", True),
                                       (comm_map.Yt_batch, "Yt", False), (comm_map.YtW_batch, "YtW", False)):
        A, B = fresh(), fresh()
        for b in range(nb):
            for m in (A[b], B[b]):
                m.alm[:] = alms[b]
                m.map[:] = pix[b]
        batch_fn(A)
        for m in B:
            getattr(m, single_fn)()
        for b in range(nb):
            got, ref = (A[b].map, B[b].map) if synth else (A[b].alm, B[b].alm)
            assert rel(got, ref) <= 1e-13, (single_fn, b)
    A = fresh()
    for b in range(nb):
        A[b].alm[:] = alms[b]
    comm_map.Y_batch(A)
    refT = S.execute(S.Y, 0, nside, lmax, alm=alms[0, 0:1])
    refP = S.execute(S.Y, 2, nside, lmax, alm=alms[0, 1:3])
    assert rel(A[0].map[0:1], refT) <= TOL and rel(A[0].map[1:3], refP) <= TOL


def _pixel_angles_ring(nside):
    """theta, phi of every pixel in RING order (HEALPix geometry as in commander3/src/sharp.f90:145-166)."""
    th, ph = [], []
    for ring in range(1, 4 * nside):
        north = 4 * nside - ring if ring > 2 * nside else ring
        if north < nside:
            cth = 1.0 - north * north / (3.0 * nside * nside)
            nph = 4 * north
            phi0 = np.pi / nph
        else:
            cth = (2 * nside - north) * 2.0 / (3.0 * nside)
            nph = 4 * nside
            phi0 = 0.0 if ((north - nside) & 1) else np.pi / nph
        if ring != north:
            cth = -cth
        th.append(np.full(nph, np.arccos(cth)))
        ph.append(phi0 + 2 * np.pi * np.arange(nph) / nph)
    return np.concatenate(th), np.concatenate(ph)


def test_closed_form_known_answers_large(shtlib):
    """Closed forms whose signs the reference fixes, at nside 1024 / lmax 2048 (no oracle involved):
    monopole, the dipole of commander3/src/comm_cmb_comp_mod.f90:145-156, and the spin-2 KATs of
    SURVEY 8c-8 (a^E_20, a^B_20, complex a^E_22 = 1) in the HEALPix COSMO convention
    (commander3/src/comm_map_mod.f90:1002)."""
    from commander_b200 import comm_map, comm_mapinfo
    nside, lmax = 1024, 2048
    info = comm_mapinfo(None, nside, lmax, 3, True)
    th, ph = _pixel_angles_ring(nside)
    m = comm_map(info)
    lm = {(int(l), int(mm)): i for i, (l, mm) in enumerate(zip(info.lm[0], info.lm[1]))}
    # T: monopole + dipole d = (dx, dy, dz)
    d = np.array([0.3, -1.1, 0.7])
    m.alm[:] = 0.0
    m.alm[0, lm[(0, 0)]] = 2.5
    m.alm[0, lm[(1, 1)]] = -np.sqrt(4 * np.pi / 3) * d[0]
    m.alm[0, lm[(1, 0)]] = np.sqrt(4 * np.pi / 3) * d[2]
    m.alm[0, lm[(1, -1)]] = np.sqrt(4 * np.pi / 3) * d[1]
    # E_20 = 1.5, B_20 = -0.5, complex E_22 = 1  (real-packed: alm[+2] = sqrt2)
    m.alm[1, lm[(2, 0)]] = 1.5
    m.alm[2, lm[(2, 0)]] = -0.5
    m.alm[1, lm[(2, 2)]] = np.sqrt(2.0)
    m.Y()
    T = 2.5 / np.sqrt(4 * np.pi) + d[0] * np.sin(th) * np.cos(ph) + d[1] * np.sin(th) * np.sin(ph) + d[2] * np.cos(th)
    k20 = -np.sqrt(15.0 / (32 * np.pi)) * np.sin(th) ** 2
    Q = 1.5 * k20 - 0.25 * np.sqrt(5 / np.pi) * (1 + np.cos(th) ** 2) * np.cos(2 * ph)
    U = -0.5 * k20 + 0.5 * np.sqrt(5 / np.pi) * np.cos(th) * np.sin(2 * ph)
    assert rel(m.map[0], T) <= 1e-13
    assert rel(m.map[1], Q) <= 1e-13
    assert rel(m.map[2], U) <= 1e-13


@pytest.mark.parametrize("nside,lmax", [(2, 9), (8, 16), (32, 64), (64, 200)])
@pytest.mark.parametrize("spin", [1, 3, 5, 8])
def test_arbitrary_spin_vs_oracle(shtlib, cpu_oracle, nside, lmax, spin):
    """sharp_execute(SHARP_Y, j, 2, ...) as commander3/src/comm_conviqt_mod.f90:234-239 calls it for j up to bmax
    (and the other three job types): (+-s)lambda of Goldberg et al., (+s)a = sgn (E+iB), (-s)a = -(E-iB),
    sgn = -1 / +1 for even / odd s.  Parity unpinned for s != 2."""
    sharp, S = shtlib, cpu_oracle
    rng = np.random.default_rng(7000 * nside + 10 * lmax + spin)
    w = rng.uniform(0.9, 1.1, 2 * nside)
    ai, gi = _handles(sharp, nside, lmax, weight=w)
    alm = rng.standard_normal((2, ai.n_local))
    mp = rng.standard_normal((2, gi.n_local))
    for job, name in ((sharp.SHARP_Y, "Y"), (sharp.SHARP_WY, "WY")):
        out = np.full((2, gi.n_local), np.nan)
        sharp.sharp_execute(job, spin, 2, alm.copy(), ai, out, gi)
        ref = S.execute(job, spin, nside, lmax, alm=alm, weight=w)
        assert rel(out, ref) <= TOL, (name, rel(out, ref))
    for job, name in ((sharp.SHARP_Yt, "Yt"), (sharp.SHARP_YtW, "YtW")):
        out = np.full((2, ai.n_local), np.nan)
        sharp.sharp_execute(job, spin, 2, out, ai, mp.copy(), gi)
        ref = S.execute(job, spin, nside, lmax, map=mp, weight=w)
        assert rel(out, ref) <= TOL, (name, rel(out, ref))
    sharp.sharp_destroy_alm_info(ai)
    sharp.sharp_destroy_geom_info(gi)


@pytest.mark.parametrize("nside,lmax", [(256, 300), (512, 200), (3, 8), (12, 30), (48, 100)])
@pytest.mark.parametrize("spin", [0, 2])
def test_midsize_and_odd_nside_vs_oracle(shtlib, cpu_oracle, nside, lmax, spin):
    """Sizes whose polar caps use the longer chirp-z classes (work lengths 2048 and 4096 of the fused ring kernel), and
    values of nside that are not powers of two (the ring scheme allows any nside): Y and YtW against the oracle."""
    sharp, S = shtlib, cpu_oracle
    rng = np.random.default_rng(17 * nside + lmax + spin)
    nc = 1 if spin == 0 else 2
    w = rng.uniform(0.9, 1.1, 2 * nside)
    ai, gi = _handles(sharp, nside, lmax, weight=w)
    alm = rng.standard_normal((nc, ai.n_local))
    if spin == 2:
        _zero_low_l(alm, lmax, None)
    out = np.full((nc, gi.n_local), np.nan)
    sharp.sharp_execute(sharp.SHARP_Y, spin, nc, alm.copy(), ai, out, gi)
    ref = S.execute(S.Y, spin, nside, lmax, alm=alm, weight=w)
    assert rel(out, ref) <= TOL, rel(out, ref)
    mp = rng.standard_normal((nc, gi.n_local))
    outa = np.full((nc, ai.n_local), np.nan)
    sharp.sharp_execute(sharp.SHARP_YtW, spin, nc, outa, ai, mp.copy(), gi)
    refa = S.execute(S.YtW, spin, nside, lmax, map=mp, weight=w)
    assert rel(outa, refa) <= TOL, rel(outa, refa)
    sharp.sharp_destroy_alm_info(ai)
    sharp.sharp_destroy_geom_info(gi)


def test_ring_fft_paths_agree(shtlib, cpu_oracle):
    """The ring-FFT stage has several execution paths per ring class (commander_b200/csrc/ringfft.cu: region_kind): batched
    cuFFT with chirp-z through HBM (the reference path of this test: every switch off), the whole-ring fused chirp-z
    kernel (plain and register-blocked), the radix-4 split chirp-z kernel (default only for the classes of work length
    16384; CMDR_SHT_SPLIT_MIN=1024 forces it for every cap ring here) and the two power-of-two kernels for the belt (one ring
    pair per CTA, the default, and one ring per CTA through a half-length transform).
    The switches are read once per process, hence the subprocesses.  Same maps and a_lm to rounding, and the all-cuFFT
    result against the oracle."""
    import os
    import subprocess
    import sys
    import tempfile
    S = cpu_oracle
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from commander_b200 import comm_map, comm_mapinfo
info = comm_mapinfo(None, 512, 700, 3, True)
m = comm_map(info)
rng = np.random.default_rng(8)
m.alm[:] = rng.standard_normal(m.alm.shape)
m.alm[1:3, info.lm[0] < 2] = 0
a0 = m.alm.copy()
m.Y(); mp = m.map.copy()
m.map[:] = rng.standard_normal(m.map.shape)
x0 = m.map.copy()
m.YtW()
np.savez(sys.argv[1], map=mp, alm=m.alm, a0=a0, x0=x0)
""" % root
    off = dict(CMDR_SHT_FUSED_BLUE="0", CMDR_SHT_RING_SPLIT="0", CMDR_SHT_BELT_FUSED="0")
    variants = {
        "cufft": off,
        "fused": dict(off, CMDR_SHT_FUSED_BLUE="1", CMDR_SHT_FFT_BLOCKED="0"),
        "blocked": dict(off, CMDR_SHT_FUSED_BLUE="1", CMDR_SHT_FFT_BLOCKED="1"),
        "split": dict(off, CMDR_SHT_RING_SPLIT="1", CMDR_SHT_SPLIT_MIN="1024"),
        "belt": dict(off, CMDR_SHT_BELT_FUSED="1"),                                  # whole ring pair per CTA
        "belt_half": dict(off, CMDR_SHT_BELT_FUSED="1", CMDR_SHT_BELT_HALF="1"),     # one ring per CTA, half-length transform
        "default": {},
    }
    res = {}
    with tempfile.TemporaryDirectory() as td:
        for name, envv in variants.items():
            out = os.path.join(td, name + ".npz")
            r = subprocess.run([sys.executable, "-c", code, out], env=dict(os.environ, **envv), capture_output=True, text=True,
                               timeout=300)
            assert r.returncode == 0, (name, r.stderr[-2000:])
            res[name] = dict(np.load(out))
    ref = res["cufft"]
    for name in ("fused", "blocked", "split", "belt", "default"):
        assert rel(res[name]["map"], ref["map"]) <= 1e-12, (name, rel(res[name]["map"], ref["map"]))
        assert rel(res[name]["alm"], ref["alm"]) <= 1e-12, (name, rel(res[name]["alm"], ref["alm"]))
        assert not np.array_equal(res[name]["map"], ref["map"]), name      # a different code path did run
    oY = np.concatenate([S.execute(S.Y, 0, 512, 700, alm=ref["a0"][0:1]), S.execute(S.Y, 2, 512, 700, alm=ref["a0"][1:3])])
    oA = np.concatenate([S.execute(S.YtW, 0, 512, 700, map=ref["x0"][0:1]), S.execute(S.YtW, 2, 512, 700, map=ref["x0"][1:3])])
    for name in variants:
        assert rel(res[name]["map"], oY) <= TOL and rel(res[name]["alm"], oA) <= TOL, name


def test_ring_kernels_with_aliasing(shtlib, cpu_oracle):
    """lmax > ring length: several m fold into one bin of the belt (n = 4 nside = 1024 <= mmax = 1100) and of the cap rings.
    The whole-ring power-of-two kernel and the radix-4 split kernel (forced for every cap class) then take their
    bin-major fold; compared with the oracle and with the all-cuFFT path."""
    import os
    import subprocess
    import sys
    import tempfile
    S = cpu_oracle
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    nside, lmax = 256, 1100
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r)
from commander_b200 import comm_map, comm_mapinfo
info = comm_mapinfo(None, %d, %d, 3, True)
m = comm_map(info)
rng = np.random.default_rng(18)
m.alm[:] = rng.standard_normal(m.alm.shape)
m.alm[1:3, info.lm[0] < 2] = 0
a0 = m.alm.copy()
m.Y(); mp = m.map.copy()
m.map[:] = rng.standard_normal(m.map.shape)
x0 = m.map.copy()
m.Yt()
np.savez(sys.argv[1], map=mp, alm=m.alm, a0=a0, x0=x0)
""" % (root, nside, lmax)
    res = {}
    with tempfile.TemporaryDirectory() as td:
        for name, envv in (("kernels", dict(CMDR_SHT_SPLIT_MIN="1024")),
                           ("cufft", dict(CMDR_SHT_FUSED_BLUE="0", CMDR_SHT_RING_SPLIT="0", CMDR_SHT_BELT_FUSED="0"))):
            out = os.path.join(td, name + ".npz")
            r = subprocess.run([sys.executable, "-c", code, out], env=dict(os.environ, **envv), capture_output=True, text=True,
                               timeout=300)
            assert r.returncode == 0, (name, r.stderr[-2000:])
            res[name] = dict(np.load(out))
    a0, x0 = res["cufft"]["a0"], res["cufft"]["x0"]
    oY = np.concatenate([S.execute(S.Y, 0, nside, lmax, alm=a0[0:1]), S.execute(S.Y, 2, nside, lmax, alm=a0[1:3])])
    oA = np.concatenate([S.execute(S.Yt, 0, nside, lmax, map=x0[0:1]), S.execute(S.Yt, 2, nside, lmax, map=x0[1:3])])
    for name in res:
        assert rel(res[name]["map"], oY) <= TOL, (name, rel(res[name]["map"], oY))
        assert rel(res[name]["alm"], oA) <= TOL, (name, rel(res[name]["alm"], oA))
    assert not np.array_equal(res["kernels"]["map"], res["cufft"]["map"])


@pytest.mark.parametrize("spin", [0, 2])
def test_empty_and_ragged_inputs(shtlib, cpu_oracle, spin):
    """What a rank with nothing to do passes (commander3/src/sharp.f90:219-224: null pointers when n_local == 0): no m's ->
    synthesis writes a zero map and analysis returns; no rings -> analysis writes zero a_lm and synthesis returns.  Plus
    ragged geometries against the oracle: the equator alone (an unpaired ring), the two polar rings, a single m."""
    sharp, S = shtlib, cpu_oracle
    nside, lmax = 8, 20
    nc = 1 if spin == 0 else 2
    rng = np.random.default_rng(spin)
    full_a = sharp.sharp_make_mmajor_real_packed_alm_info(lmax)
    full_g = sharp.sharp_make_healpix_geom_info(nside)
    no_a = sharp.sharp_make_mmajor_real_packed_alm_info(lmax, ms=np.zeros(0, dtype=np.int32))
    no_g = sharp.sharp_make_healpix_geom_info(nside, rings=np.zeros(0, dtype=np.int32))
    empty = np.zeros((nc, 0))
    mp = np.full((nc, full_g.n_local), np.nan)
    sharp.sharp_execute(sharp.SHARP_Y, spin, nc, empty, no_a, mp, full_g)
    assert np.array_equal(mp, np.zeros_like(mp))
    sharp.sharp_execute(sharp.SHARP_YtW, spin, nc, empty, no_a, rng.standard_normal(mp.shape), full_g)
    alm = np.full((nc, full_a.n_local), np.nan)
    sharp.sharp_execute(sharp.SHARP_Yt, spin, nc, alm, full_a, empty, no_g)
    assert np.array_equal(alm, np.zeros_like(alm))
    sharp.sharp_execute(sharp.SHARP_Y, spin, nc, rng.standard_normal(alm.shape), full_a, empty, no_g)
    a_in = rng.standard_normal((nc, full_a.n_local))
    if spin == 2:
        _zero_low_l(a_in, lmax, None)
    for rings in ([2 * nside], [1, 4 * nside - 1], [3, nside, 2 * nside, 3 * nside + 1]):
        g = sharp.sharp_make_healpix_geom_info(nside, rings=np.array(rings, dtype=np.int32))
        out = np.full((nc, g.n_local), np.nan)
        sharp.sharp_execute(sharp.SHARP_Y, spin, nc, a_in.copy(), full_a, out, g)
        ref = S.execute(S.Y, spin, nside, lmax, alm=a_in, rings=rings)
        assert rel(out, ref) <= TOL, (rings, rel(out, ref))
        back = np.full((nc, full_a.n_local), np.nan)
        sharp.sharp_execute(sharp.SHARP_Yt, spin, nc, back, full_a, out.copy(), g)
        refb = S.execute(S.Yt, spin, nside, lmax, map=ref, rings=rings)
        assert rel(back, refb) <= TOL, (rings, rel(back, refb))
        sharp.sharp_destroy_geom_info(g)
    one = sharp.sharp_make_mmajor_real_packed_alm_info(lmax, ms=np.array([lmax], dtype=np.int32))
    a1 = rng.standard_normal((nc, one.n_local))
    out = np.full((nc, full_g.n_local), np.nan)
    sharp.sharp_execute(sharp.SHARP_Y, spin, nc, a1.copy(), one, out, full_g)
    assert rel(out, S.execute(S.Y, spin, nside, lmax, alm=a1, ms=[lmax])) <= TOL
    for h in (full_a, no_a, one):
        sharp.sharp_destroy_alm_info(h)
    for h in (full_g, no_g):
        sharp.sharp_destroy_geom_info(h)
