"""Band evaluation / projection of a diffuse component: the mirror of the part of
`comm_diffuse_comp_mod` that sits on the SHT hot path (commander3/src/comm_diffuse_comp_mod.f90).

    evalDiffuseBand    :2027-2108   amplitude alm -> [mixing] -> beam -> band alm or band map
    projectDiffuseBand :2110-2167   band map or alm -> beam^T -> [mixing] -> amplitude alm (its transpose)

The mixing step has two branches in the reference (:2072-2082, :2142-2151):
  * spectral parameters constant on the sky: alm(:,i) *= F_mean(i)          (a scalar per component)
  * spatially varying:  call m%Y(); m%map = m%map * F%map; call m%YtW()      (one SHT pair per band)
The second branch is the expensive one; here it is ONE library call (`cmdr_sht_mix`) that keeps the
map on the device between the two transforms, instead of four sharp_execute calls with a host-side
multiply in between.

Only what touches the SHT is mirrored: no spectral-model evaluation (F is an input), no detector
index, no `F_null` bookkeeping.
"""
from __future__ import annotations

import numpy as np

from . import sharp
from .comm_map import comm_map, comm_mapinfo


_conv_cache: dict = {}   # per (layout, beam, device): b_l expanded over the alm slots


def mix(m: comm_map, F) -> None:
    """m%alm <- YtW(F .* Y(m%alm)); F has shape (nmaps, np), on the host (numpy) or the device (torch).
    commander3/src/comm_diffuse_comp_mod.f90:2078-2080 / :2148-2150."""
    info = m.info
    if info.nmaps not in (1, 3) or (info.nmaps == 3 and not info.pol):
        # other column counts: independent scalar transforms, as comm_map%Y does for pol = .false.
        m.Y()
        m.map *= F
        m.YtW()
        return
    sharp.mix(m.alm, F, info.geom_info_T, info.geom_info_P if info.nmaps == 3 else None, info.alm_info,
              nmaps=info.nmaps, comm=m._comm())


def _conv(m: comm_map, b_l) -> None:
    """B%conv, commander3/src/comm_B_bl_mod.f90:108-127: alm(l,m,j) *= b_l(l,j) (its own transpose)."""
    l = m.info.lm[0]
    nm = m.info.nmaps
    if m.device is None:
        m.alm *= np.stack([np.asarray(b_l)[l, j] for j in range(nm)])
    else:
        import torch
        key = (id(m.info), id(b_l), str(m.alm.device))
        fac = _conv_cache.get(key)
        if fac is None:
            fac = (torch.as_tensor(np.stack([np.asarray(b_l)[l, j] for j in range(nm)]), device=m.alm.device), m.info, b_l)
            _conv_cache[key] = fac
        m.alm *= fac[0]


class diffuse_band:
    """One (component, band) pair: amplitude layout `x_info`, band layout `band_info`, beam b_l (lmax+1, nmaps),
    and either F_mean (nmaps scalars; spectral parameters constant on the sky) or F (nmaps, np) per pixel."""

    def __init__(self, x_info: comm_mapinfo, band_info: comm_mapinfo, b_l, F=None, F_mean=None, device=None):
        if (F is None) == (F_mean is None):
            raise ValueError("give exactly one of F (per pixel) and F_mean (per component)")
        self.x_info, self.band_info, self.b_l, self.F, self.F_mean, self.device = x_info, band_info, b_l, F, F_mean, device

    def _apply_mixmat(self, m: comm_map) -> None:
        if self.F_mean is not None:
            for i in range(m.info.nmaps):
                m.alm[i] *= float(self.F_mean[i])
        else:
            mix(m, self.F)

    def evalDiffuseBand(self, x: comm_map, alm_out: bool = False):
        """commander3/src/comm_diffuse_comp_mod.f90:2027-2108."""
        m = comm_map(self.band_info, device=self.device)
        x.alm_equal(m)                       # :2064
        self._apply_mixmat(m)                # :2070-2083
        _conv(m, self.b_l)                   # :2086
        if alm_out:
            return m.alm
        m.Y()                                # :2087
        return m.map

    def projectDiffuseBand(self, band_map: comm_map, alm_in: bool = False):
        """commander3/src/comm_diffuse_comp_mod.f90:2110-2167 (the transpose of evalDiffuseBand)."""
        m = comm_map(self.band_info, device=self.device)
        if alm_in:
            m.alm[...] = band_map.alm        # :2137
        else:
            m.map[...] = band_map.map        # :2139-2140
            m.Yt()
        _conv(m, self.b_l)                   # :2142
        self._apply_mixmat(m)                # :2144-2151
        out = comm_map(self.x_info, device=self.device)
        m.alm_equal(out)                     # :2152
        return out.alm
