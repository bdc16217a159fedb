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
