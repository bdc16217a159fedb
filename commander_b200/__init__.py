"""commander_b200 -- B200-native SHT engine behind Commander3's comm_map Y/Yt/YtW/WY.

Host-side mirror of the reference's Fortran interface for this one path
(commander3/src/sharp.f90, commander3/src/comm_map_mod.f90, commander3/src/comm_cr_mod.f90)
on top of the C-ABI library commander_b200/lib/libcmdr_sht.so (include/cmdr_sht.h).
There is no CPU fallback: every transform runs CUDA kernels.
"""
from . import sharp  # noqa: F401
from .comm_map import comm_map, comm_mapinfo  # noqa: F401

__all__ = ["sharp", "comm_map", "comm_mapinfo"]
