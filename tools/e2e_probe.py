#!/usr/bin/env python
"""e2e_probe.py -- where does the host-buffer (pinned) path spend its time?  (tuning aid, GPU only)
Times each of the four sharp_execute calls of one Y / YtW pair with host buffers against the same call with
device-resident buffers, and the plain pinned H2D / D2H copy bandwidth of the box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from commander_b200 import comm_map, comm_mapinfo, sharp

nside, lmax = int(os.environ.get("NSIDE", 2048)), int(os.environ.get("LMAX", 4000))
dev = torch.device("cuda", 0)
info = comm_mapinfo(None, nside, lmax, 3, True)
md = comm_map(info, device=dev)
md.alm.normal_()
md.alm[1:3, torch.as_tensor(info.lm[0] < 2, device=dev)] = 0
h = comm_map(info)
pa = torch.empty((3, info.nalm), dtype=torch.float64).pin_memory(); pm = torch.empty((3, info.np), dtype=torch.float64).pin_memory()
pa.copy_(md.alm.cpu()); h.alm, h.map = pa.numpy(), pm.numpy()

def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

def call(mm, job, spin):
    if spin == 0:
        return lambda: sharp.sharp_execute(job, 0, 1, mm.alm[0:1], info.alm_info, mm.map[0:1], info.geom_info_T)
    return lambda: sharp.sharp_execute(job, 2, 2, mm.alm[1:3], info.alm_info, mm.map[1:3], info.geom_info_P)

buf = torch.empty(info.np, dtype=torch.float64, device=dev)
for name, fn in (("H2D", lambda: buf.copy_(pm[0], non_blocking=True)), ("D2H", lambda: pm[1].copy_(buf, non_blocking=True))):
    ms = t(fn, 5); print(f"pinned {name}: {info.np * 8 / ms / 1e6:.1f} GB/s ({ms:.2f} ms for {info.np * 8 / 1e6:.0f} MB)")
def both():
    s2 = both.s2
    with torch.cuda.stream(s2): pm[1].copy_(buf, non_blocking=True)
    both.b2.copy_(pm[0], non_blocking=True)
both.s2 = torch.cuda.Stream(); both.b2 = torch.empty_like(buf)
ms = t(both, 5); print(f"pinned H2D + D2H concurrently: {info.np * 8 / ms / 1e6:.1f} GB/s each way")
JOBS = (("Y", sharp.SHARP_Y), ("YtW", sharp.SHARP_YtW))
tot_d = tot_h = 0
for jn, job in JOBS:
    for spin in (0, 2):
        d, hh = t(call(md, job, spin)), t(call(h, job, spin))
        tot_d += d; tot_h += hh
        nin = (info.nalm if jn == "Y" else info.np) * (1 if spin == 0 else 2) * 8
        nout = (info.np if jn == "Y" else info.nalm) * (1 if spin == 0 else 2) * 8
        print(f"{jn:4s} spin {spin}: device {d:7.2f} ms   host {hh:7.2f} ms   (+{hh - d:5.2f})   in {nin / 1e6:6.0f} MB out {nout / 1e6:6.0f} MB")
print(f"pair: device {tot_d:.2f} ms, host {tot_h:.2f} ms")
