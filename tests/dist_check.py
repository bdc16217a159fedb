"""torchrun worker: distributed transforms on N GPUs vs the single-GPU result and the CPU oracle.
Launched by tests/test_gpu_dist.py (or by hand:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from commander_b200 import comm_map, comm_mapinfo, sharp
    from commander_b200 import dist as cdist
    from oracle import sht_cpu as S
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    comm = cdist.init_from_torch(dev)
    worst = 0.0
    for nside, lmax in ((16, 40), (64, 150), (128, 383)):
        info = comm_mapinfo(comm, nside, lmax, 3, True)
        # global inputs from a fixed seed, identical on all ranks
        rng = np.random.default_rng(1234 + nside)
        nalm_g, npix_g = (lmax + 1) ** 2, 12 * nside ** 2
        alm_g = rng.standard_normal((3, nalm_g))
        map_g = rng.standard_normal((3, npix_g))
        w = rng.uniform(0.9, 1.1, (2, 2 * nside))
        infow = comm_mapinfo(comm, nside, lmax, 3, True, weights=w)
        # global real-packed index of my local alm slots (m-major over all m)
        mstart = np.zeros(lmax + 2, dtype=np.int64)
        for m in range(lmax + 1):
            mstart[m + 1] = mstart[m] + (lmax + 1 - m) * (1 if m == 0 else 2)
        l, mm = infow.lm[0].astype(np.int64), infow.lm[1].astype(np.int64)
        am = np.abs(mm)
        gidx = mstart[am] + np.where(am == 0, l, 2 * (l - am) + (mm < 0))
        # oracle on the full sphere
        refY = np.concatenate([S.execute(S.Y, 0, nside, lmax, alm=alm_g[0:1]), S.execute(S.Y, 2, nside, lmax, alm=alm_g[1:3])])
        refA = np.concatenate([S.execute(S.YtW, 0, nside, lmax, map=map_g[0:1], weight=w[0]),
                               S.execute(S.YtW, 2, nside, lmax, map=map_g[1:3], weight=w[1])])
        for fused in (False, True):
            for device in (None, dev):
                m = comm_map(infow, device=device)
                if device is None:
                    m.alm[:] = alm_g[:, gidx]
                else:
                    m.alm.copy_(torch.as_tensor(alm_g[:, gidx], device=dev))
                (m.Y_iqu if fused else m.Y)()
                out = m.map if device is None else m.map.cpu().numpy()
                e1 = np.linalg.norm(out - refY[:, infow.pix]) / np.linalg.norm(refY[:, infow.pix])
                if device is None:
                    m.map[:] = map_g[:, infow.pix]
                else:
                    m.map.copy_(torch.as_tensor(map_g[:, infow.pix], device=dev))
                (m.YtW_iqu if fused else m.YtW)()
                out = m.alm if device is None else m.alm.cpu().numpy()
                e2 = np.linalg.norm(out - refA[:, gidx]) / np.linalg.norm(refA[:, gidx])
                worst = max(worst, e1, e2)
                assert e1 <= 1e-10 and e2 <= 1e-10, (nside, lmax, fused, device, e1, e2)
        # arbitrary spin through the distributed entry point (conviqt calls sharp_execute with comm=, spin=j)
        if nside == 64:
            for spin in (1, 3):
                ref = S.execute(S.Y, spin, nside, lmax, alm=alm_g[1:3])
                out = np.zeros((2, infow.np))
                sharp.sharp_execute(sharp.SHARP_Y, spin, 2, np.ascontiguousarray(alm_g[1:3][:, gidx]), infow.alm_info, out,
                                    infow.geom_info_T, comm=comm.handle)
                e3 = np.linalg.norm(out - ref[:, infow.pix]) / np.linalg.norm(ref[:, infow.pix])
                refa = S.execute(S.Yt, spin, nside, lmax, map=map_g[1:3])
                outa = np.zeros((2, infow.nalm))
                sharp.sharp_execute(sharp.SHARP_Yt, spin, 2, outa, infow.alm_info, np.ascontiguousarray(map_g[1:3][:, infow.pix]),
                                    infow.geom_info_T, comm=comm.handle)
                e4 = np.linalg.norm(outa - refa[:, gidx]) / np.linalg.norm(refa[:, gidx])
                worst = max(worst, e3, e4)
                assert e3 <= 1e-10 and e4 <= 1e-10, (spin, e3, e4)
        # conviqt cube through the distributed entry point (beam table synchronised over ranks, spin-j syntheses
        # with the fused exchange, psi transform on the local pixels)
        if nside == 64:
            from commander_b200.comm_conviqt import comm_conviqt
            from oracle import conviqt as O
            bmax = 3
            beam_g = rng.standard_normal((3, nalm_g))
            sky, beam = comm_map(infow), comm_map(infow)
            sky.alm[:] = alm_g[:, gidx]
            beam.alm[:] = beam_g[:, gidx]
            cv = comm_conviqt(nside, lmax, 3, bmax, beam, sky, precompute=False)
            assert cv.info.np == infow.np and cv.info.nalm == infow.nalm
            c64 = np.zeros((2 * bmax, infow.np))
            cv.precompute_sky(sky, cube=c64)
            lm = O.lm_table(lmax)
            tab = O.beam_table(lmax, 3, lm, beam_g)
            assert np.array_equal(tab.view(np.float32), cv.alm_beam.view(np.float32))
            refc = O.precompute_sky(S, nside, lmax, bmax, alm_g, tab)[:, infow.pix]
            e5 = np.linalg.norm(c64 - refc) / np.linalg.norm(refc)
            worst = max(worst, e5)
            assert e5 <= 1e-10, ("conviqt", e5)
        infow.dealloc()
    # pinned host buffers at a size where the chunked pipeline (per-chunk exchange barriers, PCIe copies beside
    # the Legendre kernels) is taken
    nside, lmax = 512, 767
    info = comm_mapinfo(comm, nside, lmax, 3, True)
    rng = np.random.default_rng(4321)
    alm_g = rng.standard_normal((3, (lmax + 1) ** 2))
    map_g = rng.standard_normal((3, 12 * nside ** 2))
    mstart = np.zeros(lmax + 2, dtype=np.int64)
    for m in range(lmax + 1):
        mstart[m + 1] = mstart[m] + (lmax + 1 - m) * (1 if m == 0 else 2)
    l, mm = info.lm[0].astype(np.int64), info.lm[1].astype(np.int64)
    am = np.abs(mm)
    gidx = mstart[am] + np.where(am == 0, l, 2 * (l - am) + (mm < 0))
    m = comm_map(info)
    m.alm = torch.empty((3, info.nalm), dtype=torch.float64).pin_memory().numpy()
    m.map = torch.empty((3, info.np), dtype=torch.float64).pin_memory().numpy()
    m.alm[:] = alm_g[:, gidx]
    m.Y()
    refY = np.concatenate([S.execute(S.Y, 0, nside, lmax, alm=alm_g[0:1]), S.execute(S.Y, 2, nside, lmax, alm=alm_g[1:3])])
    e1 = np.linalg.norm(m.map - refY[:, info.pix]) / np.linalg.norm(refY[:, info.pix])
    m.map[:] = map_g[:, info.pix]
    m.Yt()
    refA = np.concatenate([S.execute(S.Yt, 0, nside, lmax, map=map_g[0:1]), S.execute(S.Yt, 2, nside, lmax, map=map_g[1:3])])
    e2 = np.linalg.norm(m.alm - refA[:, gidx]) / np.linalg.norm(refA[:, gidx])
    worst = max(worst, e1, e2)
    assert e1 <= 1e-10 and e2 <= 1e-10, ("pinned pipelined", e1, e2)
    info.dealloc()
    # CG dot-product all-reduce
    t = torch.full((3,), float(rank + 1), dtype=torch.float64, device=dev)
    comm.allreduce_sum_(t)
    torch.cuda.synchronize()
    assert float(t[0]) == world * (world + 1) / 2
    tw = torch.tensor([worst], dtype=torch.float64, device=dev)
    dist.all_reduce(tw, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"DIST_CHECK_OK world={world} worst_rel_l2={float(tw):.3e}")
    cdist.destroy(comm)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
