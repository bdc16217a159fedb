"""comm_map operators that the round-1 suite did not call directly: Y_scalar, Yt_scalar,
YtW_scalar, Y_EB (commander3/src/comm_map_mod.f90:477-509, 532-544, 567-579) against the oracle, host (pageable
numpy arrays, what the Fortran caller passes) and device resident."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-10


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("device", [None, "cuda"])
@pytest.mark.parametrize("nmaps,pol", [(3, True), (3, False), (1, False), (2, False)])
def test_scalar_family_vs_oracle(shtlib, cpu_oracle, device, nmaps, pol):
    import torch
    from commander_b200 import comm_map, comm_mapinfo
    S = cpu_oracle
    nside, lmax = 32, 70
    rng = np.random.default_rng(5 + nmaps)
    w = rng.uniform(0.9, 1.1, (1 if nmaps == 1 else 2, 2 * nside))
    info = comm_mapinfo(None, nside, lmax, nmaps, pol, weights=w)
    alm = rng.standard_normal((nmaps, info.nalm))
    mp = rng.standard_normal((nmaps, info.np))

    def put(dst, src):
        if device is None:
            dst[...] = src
        else:
            dst.copy_(torch.as_tensor(src))

    def get(x):
        return x if device is None else x.cpu().numpy()

    refY = np.concatenate([S.execute(S.Y, 0, nside, lmax, alm=alm[i:i + 1]) for i in range(nmaps)])
    refYt = np.concatenate([S.execute(S.Yt, 0, nside, lmax, map=mp[i:i + 1]) for i in range(nmaps)])
    refYtW = np.concatenate([S.execute(S.YtW, 0, nside, lmax, map=mp[i:i + 1], weight=w[0]) for i in range(nmaps)])
    m = comm_map(info, device=device)
    put(m.alm, alm); m.Y_scalar()
    assert rel(get(m.map), refY) <= TOL
    put(m.alm, alm); put(m.map, 0 * mp); m.Y_EB()                 # every column as a spin-0 field, :491-509
    assert rel(get(m.map), refY) <= TOL
    put(m.map, mp); m.Yt_scalar()
    assert rel(get(m.alm), refYt) <= TOL
    put(m.map, mp); m.YtW_scalar()
    assert rel(get(m.alm), refYtW) <= TOL
    if not pol:
        # a non-polarised object's Y / Yt / YtW are the scalar loops (the else branches of :450-453, 525-527, 559-561)
        put(m.alm, alm); m.Y()
        assert rel(get(m.map), refY) <= TOL
        put(m.map, mp); m.YtW()
        assert rel(get(m.alm), refYtW) <= TOL
    info.dealloc()
