"""One device-resident Y + Yt pair (IQU) at a chosen size, after a warm-up pair: the target of ncu launch lists for the
CG-sized (nside 1024 / lmax 2000) and band-sized (nside 512 / lmax 1500) transforms."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from commander_b200 import comm_map, comm_mapinfo, sharp

nside, lmax = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda", 0)
info = comm_mapinfo(None, nside, lmax, 3, True)
m = comm_map(info, device=dev)
m.alm.normal_()
m.alm[1:3, torch.as_tensor(info.lm[0] < 2, device=dev)] = 0
for it in range(2):
    n0 = sharp.launch_count()
    torch.cuda.synchronize()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    e[0].record(); m.Y(); m.Yt(); e[1].record()
    torch.cuda.synchronize()
    print(f"pair {it}: {e[0].elapsed_time(e[1]):.3f} ms, {sharp.launch_count() - n0} launches")
