"""One conviqt cube (comm_conviqt%precompute_sky) at a chosen size after a warm-up: target for ncu on the two
conviqt kernels (conviqt_alms_kernel, conviqt_psi_kernel).  usage: conviqt_probe.py nside lmax bmax"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from commander_b200 import comm_map, comm_mapinfo
from commander_b200.comm_conviqt import comm_conviqt

nside, lmax, bmax = (int(a) for a in sys.argv[1:4])
dev = torch.device("cuda", 0)
info = comm_mapinfo(None, nside, lmax, 3, True)
rng = np.random.default_rng(1)
sky = comm_map(info, device=dev)
sky.alm.normal_()
beam = comm_map(info)
beam.alm[...] = rng.standard_normal(beam.alm.shape) / (1.0 + info.lm[0])
cv = comm_conviqt(nside, lmax, 3, bmax, beam, sky, device=dev)
torch.cuda.synchronize()
e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
e[0].record(); cv.precompute_sky(sky); e[1].record()
torch.cuda.synchronize()
npix = info.np
print(f"cube {e[0].elapsed_time(e[1]):.3f} ms; psi kernel algorithmic bytes {npix * (8 * (2 * bmax + 1) + 4 * 2 * bmax)}")
