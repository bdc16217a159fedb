"""World-size-2 gloo test (CPU) of the multi-rank host logic: the comm_mapinfo partition
(commander3/src/comm_map_mod.f90:197-261) as seen from two real processes, and the exchange
semantics the GPU path implements (Legendre on own m's for every rank's rings, all-to-all, sum),
emulated with the CPU oracle as the per-rank engine.  No CUDA involved."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nside, lmax, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from commander_b200.comm_map import comm_mapinfo
    from commander_b200.dist import Comm
    from oracle import sht_cpu as S

    info = comm_mapinfo(Comm(rank, world, 7), nside, lmax, 3, True)
    # 1. every pixel and every (l,m) is owned exactly once across the group
    counts = torch.tensor([info.np, info.nalm, info.nm, info.nring], dtype=torch.int64)
    allc = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(allc, counts)
    tot = torch.stack(allc).sum(0)
    assert int(tot[0]) == 12 * nside ** 2 and int(tot[1]) == (lmax + 1) ** 2
    assert int(tot[2]) == lmax + 1 and int(tot[3]) == 4 * nside - 1
    own = torch.zeros(12 * nside ** 2, dtype=torch.int32)
    own[torch.as_tensor(info.pix)] = 1
    dist.all_reduce(own)
    assert int(own.min()) == 1 and int(own.max()) == 1
    # 2. peers' ring and m lists, as the library discovers them with all-gathers
    maxr, maxm = int(torch.stack(allc)[:, 3].max()), int(torch.stack(allc)[:, 2].max())
    pad_r = torch.full((maxr,), -1, dtype=torch.int32); pad_r[:info.nring] = torch.as_tensor(info.rings)
    pad_m = torch.full((maxm,), -1, dtype=torch.int32); pad_m[:info.nm] = torch.as_tensor(info.ms)
    all_r = [torch.zeros_like(pad_r) for _ in range(world)]
    all_m = [torch.zeros_like(pad_m) for _ in range(world)]
    dist.all_gather(all_r, pad_r); dist.all_gather(all_m, pad_m)
    rings_of = [t[t >= 0].numpy() for t in all_r]
    ms_of = [t[t >= 0].numpy() for t in all_m]
    # 3. distributed synthesis == single-rank synthesis restricted to my rings (spin 0 and 2)
    rng = np.random.default_rng(100 + rank)
    for spin, cols in ((0, slice(0, 1)), (2, slice(1, 3))):
        nc = 1 if spin == 0 else 2
        alm_local = rng.standard_normal((nc, info.nalm))
        # "Legendre for my m's on every rank's rings", then exchange and sum
        parts = [torch.as_tensor(S.execute(S.Y, spin, nside, lmax, alm=alm_local, ms=info.ms, rings=rings_of[r]))
                 for r in range(world)]
        recv = [torch.zeros((nc, info.np), dtype=torch.float64) for _ in range(world)]
        reqs = []
        for r in range(world):
            if r == rank:
                recv[r] = parts[r]
            else:
                reqs.append(dist.isend(parts[r].contiguous(), dst=r))
                reqs.append(dist.irecv(recv[r], src=r))
        for rq in reqs:
            rq.wait()
        mine = sum(recv).numpy()
        # reference: gather every rank's alm into the global real-packed vector, one synthesis
        nal = [S.alm_count(lmax, ms_of[r]) for r in range(world)]
        bufs = [torch.zeros((nc, nal[r]), dtype=torch.float64) for r in range(world)]
        dist.all_gather(bufs, torch.as_tensor(alm_local)) if len(set(nal)) == 1 else None
        if len(set(nal)) != 1:   # unequal sizes: broadcast one by one
            for r in range(world):
                if r == rank:
                    bufs[r] = torch.as_tensor(alm_local).clone()
                dist.broadcast(bufs[r], src=r)
        ref = sum(S.execute(S.Y, spin, nside, lmax, alm=bufs[r].numpy(), ms=ms_of[r], rings=info.rings)
                  for r in range(world))
        err = np.linalg.norm(mine - ref) / np.linalg.norm(ref)
        assert err < 1e-13, err
        # analysis: my rings contribute to every rank's m's; contributions sum over ring owners
        mp = rng.standard_normal((nc, info.np))
        parts = [torch.as_tensor(S.execute(S.YtW, spin, nside, lmax, map=mp, ms=ms_of[r], rings=info.rings))
                 for r in range(world)]
        recv = [torch.zeros((nc, info.nalm), dtype=torch.float64) for _ in range(world)]
        reqs = []
        for r in range(world):
            if r == rank:
                recv[r] = parts[r]
            else:
                reqs.append(dist.isend(parts[r].contiguous(), dst=r))
                reqs.append(dist.irecv(recv[r], src=r))
        for rq in reqs:
            rq.wait()
        mine = sum(recv).numpy()
        # reference: full-sky map assembled from all ranks, one analysis for my m's
        full = torch.zeros((nc, 12 * nside ** 2), dtype=torch.float64)
        full[:, torch.as_tensor(info.pix)] = torch.as_tensor(mp)
        dist.all_reduce(full)
        ref = S.execute(S.YtW, spin, nside, lmax, map=full.numpy(), ms=info.ms)
        err = np.linalg.norm(mine - ref) / np.linalg.norm(ref)
        assert err < 1e-13, err
    # 4. conviqt beam table: every rank contributes its own m's, all ranks end with the full single-precision table
    #    (sync_shared_2d_spc_alm, commander3/src/comm_conviqt_mod.f90:124-125)
    from commander_b200.comm_map import comm_map
    from commander_b200.comm_conviqt import comm_conviqt
    from oracle import conviqt as O
    grng = np.random.default_rng(55)
    lm_g = O.lm_table(lmax)
    beam_g = grng.standard_normal((3, len(lm_g)))
    pos = {t: i for i, t in enumerate(lm_g)}
    gidx = np.array([pos[(int(info.lm[0, i]), int(info.lm[1, i]))] for i in range(info.nalm)])
    sky, beam = comm_map(info), comm_map(info)
    beam.alm[...] = beam_g[:, gidx]
    cv = comm_conviqt(nside, lmax, 3, 2, beam, sky, precompute=False)
    tab = O.beam_table(lmax, 3, lm_g, beam_g)
    assert np.array_equal(cv.alm_beam.view(np.float32), tab.view(np.float32))
    ref = O.get_alms(1, lmax, lm_g, beam_g, tab)[:, gidx]        # any a_lm serve as the sky
    sky.alm[...] = beam_g[:, gidx]
    assert np.allclose(cv.get_alms(1, sky), ref, rtol=0, atol=1e-14 * np.abs(ref).max())
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, "ok"))


def test_two_rank_partition_and_exchange(cpu_oracle, shtlib):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world, nside, lmax = 2, 8, 19
    procs = [ctx.Process(target=_worker, args=(r, world, port, nside, lmax, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    codes = [p.exitcode for p in procs]
    assert codes == [0, 0], codes
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got == [(0, "ok"), (1, "ok")]
