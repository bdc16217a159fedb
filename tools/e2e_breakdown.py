"""Times the four sharp_execute calls of one IQU pair with pinned host buffers (tuning aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from commander_b200 import comm_map, comm_mapinfo, sharp
nside, lmax = 2048, 4000
info = comm_mapinfo(None, nside, lmax, 3, True)
h = comm_map(info)
pa = torch.empty((3, info.nalm), dtype=torch.float64).pin_memory(); pm = torch.empty((3, info.np), dtype=torch.float64).pin_memory()
pa.normal_(); h.alm, h.map = pa.numpy(), pm.numpy()
def t(f):
    torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
ai, gT, gP = info.alm_info, info.geom_info_T, info.geom_info_P
calls = {"Y T": lambda: sharp.sharp_execute(sharp.SHARP_Y, 0, 1, h.alm[0:1], ai, h.map[0:1], gT),
         "Y QU": lambda: sharp.sharp_execute(sharp.SHARP_Y, 2, 2, h.alm[1:3], ai, h.map[1:3], gP),
         "YtW T": lambda: sharp.sharp_execute(sharp.SHARP_YtW, 0, 1, h.alm[0:1], ai, h.map[0:1], gT),
         "YtW QU": lambda: sharp.sharp_execute(sharp.SHARP_YtW, 2, 2, h.alm[1:3], ai, h.map[1:3], gP)}
for k, f in calls.items(): f()
res = {k: min(t(f) for _ in range(3)) for k, f in calls.items()}
# raw copy bandwidth
d = torch.empty_like(pm, device="cuda")
bw_h2d = pm.numel() * 8 / 1e6 / min(t(lambda: d.copy_(pm, non_blocking=True)) for _ in range(3))
bw_d2h = pm.numel() * 8 / 1e6 / min(t(lambda: pm.copy_(d, non_blocking=True)) for _ in range(3))
print("chunks", os.environ.get("CMDR_SHT_CHUNKS", "8"), "nopipe" if os.environ.get("CMDR_SHT_NO_PIPELINE") else "pipe",
      {k: round(v, 1) for k, v in res.items()}, "sum %.1f ms" % sum(res.values()), "H2D %.1f GB/s D2H %.1f GB/s" % (bw_h2d, bw_d2h))
