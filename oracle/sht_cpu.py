"""ctypes front-end of oracle/sht_cpu.c -- TEST INFRASTRUCTURE ONLY (see that file's header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _has_avx512() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        return all(k in txt for k in ("avx512f", "avx512dq", "avx512bw", "avx512vl", "avx512cd"))
    except OSError:
        return False


def build(force: bool = False) -> None:
    out = os.path.join(_HERE, "_build")
    want = [os.path.join(out, f"libsht_cpu_{v}.so") for v in ("v3", "v4")]
    src = os.path.join(_HERE, "sht_cpu.c")
    if not force and all(os.path.exists(w) and os.path.getmtime(w) >= os.path.getmtime(src) for w in want):
        return
    subprocess.check_call(["make", "-C", _HERE, "-s", "all"])


def lib():
    global _lib
    if _lib is None:
        out = os.path.join(_HERE, "_build")
        v = "v4" if _has_avx512() else "v3"
        path = os.path.join(out, f"libsht_cpu_{v}.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.osht_execute.restype = C.c_int
        L.osht_execute.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.osht_map_size.restype = C.c_int64
        L.osht_map_size.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.osht_alm_count.restype = C.c_int64
        L.osht_alm_count.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.osht_set_mlim_skip.argtypes = [C.c_int]
        L.osht_max_threads.restype = C.c_int
        L.osht_last_times.argtypes = [C.c_void_p]
        L.variant = v
        _lib = L
    return _lib


YtW, Y, Yt, WY = 0, 1, 2, 3


def _iarr(x):
    return None if x is None else np.ascontiguousarray(x, dtype=np.int32)


def map_size(nside, rings=None):
    r = _iarr(rings)
    n = 4 * nside - 1 if r is None else len(r)
    return int(lib().osht_map_size(nside, n, None if r is None else r.ctypes.data))


def alm_count(lmax, ms=None):
    m = _iarr(ms)
    n = lmax + 1 if m is None else len(m)
    return int(lib().osht_alm_count(lmax, n, None if m is None else m.ctypes.data))


def execute(job, spin, nside, lmax, alm=None, map=None, rings=None, ms=None, weight=None,
            add=False, nthreads=0, mlim_skip=False):
    """alm: (ncomp, nalm) float64, map: (ncomp, npix) float64.  Output array is created
    when None.  Returns the output array (map for Y/WY, alm for Yt/YtW)."""
    L = lib()
    ncomp = 1 if spin == 0 else 2
    r = _iarr(rings)
    m = _iarr(ms)
    nr = 4 * nside - 1 if r is None else len(r)
    nm = lmax + 1 if m is None else len(m)
    npix = map_size(nside, rings)
    nalm = alm_count(lmax, ms)
    synth = job in (Y, WY)
    if synth:
        alm = np.ascontiguousarray(np.asarray(alm, dtype=np.float64).reshape(ncomp, nalm))
        if map is None:
            map = np.zeros((ncomp, npix))
    else:
        map = np.ascontiguousarray(np.asarray(map, dtype=np.float64).reshape(ncomp, npix))
        if alm is None:
            alm = np.zeros((ncomp, nalm))
    assert alm.flags.c_contiguous and map.flags.c_contiguous
    w = None if weight is None else np.ascontiguousarray(weight, dtype=np.float64)
    ap = (C.c_void_p * ncomp)(*[alm[c].ctypes.data for c in range(ncomp)])
    mp = (C.c_void_p * ncomp)(*[map[c].ctypes.data for c in range(ncomp)])
    L.osht_set_mlim_skip(1 if mlim_skip else 0)
    rc = L.osht_execute(job, spin, nside, lmax, nr, None if r is None else r.ctypes.data,
                        None if w is None else w.ctypes.data, nm, None if m is None else m.ctypes.data,
                        ap, mp, 1 if add else 0, nthreads)
    if rc != 0:
        raise RuntimeError(f"osht_execute failed rc={rc}")
    return map if synth else alm


def last_times():
    """(legendre seconds, fft seconds) of the most recent execute()."""
    t = (C.c_double * 2)()
    lib().osht_last_times(t)
    return float(t[0]), float(t[1])
