"""DAB Mode I tables (include/dabgpu_tables.h is the source of truth; tests/test_tables.py pins it
against the compiled reference).  They are read through libdabtables.so, a host-only build of the
same dabgpu_tab_* accessors libdabgpu.so exports (csrc/tables_host.c): manufacturing test or
benchmark input never maps the CUDA library."""
from __future__ import annotations

import ctypes as C
import functools
import os

import numpy as np


class _TablesLib:
    _lib = None

    @classmethod
    def load(cls):
        if cls._lib is None:
            from . import build as _b
            so = _b.TABLES_LIB
            if not os.path.exists(so):
                _b.build_tables()
            lib = C.CDLL(so)
            lib.dabgpu_tab_shape.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]
            lib.dabgpu_tab_uep.argtypes = [C.POINTER(C.c_int32)]
            lib.dabgpu_tab_puncture_mask.restype = C.c_uint32
            lib.dabgpu_tab_freq_deint.argtypes = [C.POINTER(C.c_uint16)]
            lib.dabgpu_tab_prs.argtypes = [C.POINTER(C.c_uint8)]
            lib.dabgpu_tab_prbs.argtypes = [C.POINTER(C.c_uint8), C.c_int]
            lib.dabgpu_tab_crc16.argtypes = [C.POINTER(C.c_uint8), C.c_int, C.c_uint16]
            lib.dabgpu_tab_crc16.restype = C.c_uint16
            cls._lib = lib
        return cls._lib


_lib = _TablesLib

POLYS = (0x6D, 0x4F, 0x53, 0x6D)
TDI_DELAY = (0, 8, 4, 12, 2, 10, 6, 14, 1, 9, 5, 13, 3, 11, 7, 15)


def _shape(kind, a=0, b=0):
    out = (C.c_int32 * 23)()
    rc = _lib.load().dabgpu_tab_shape(kind, a, b, out)
    if rc:
        raise ValueError(f"no such protection profile kind={kind} a={a} b={b}")
    v = np.frombuffer(out, dtype=np.int32).copy()
    regions = [dict(steps=int(r[0]), pi=int(r[1]), step0=int(r[2]), in0=int(r[3]))
               for r in v[3:].reshape(5, 4)[: int(v[2])]]
    return dict(nbits=int(v[0]), in_bits=int(v[1]), n_regions=int(v[2]), regions=regions)


@functools.lru_cache(None)
def shape_fic():
    return _shape(0)


@functools.lru_cache(None)
def shape_uep(index: int):
    return _shape(1, index)


@functools.lru_cache(None)
def shape_eep(level: int, size_cu: int):
    return _shape(2, level, size_cu)


def _uep():
    out = (C.c_int32 * (64 * 12))()
    _lib.load().dabgpu_tab_uep(out)
    v = np.frombuffer(out, dtype=np.int32).reshape(64, 12)
    return [(int(r[0]), int(r[1]), int(r[2]), tuple(int(x) for x in r[3:7]), tuple(int(x) for x in r[7:11]),
             int(r[11])) for r in v]


class _Lazy(list):
    def __init__(self, fn):
        super().__init__()
        self._fn = fn

    def _fill(self):
        if not len(self):
            self.extend(self._fn())

    def __getitem__(self, i):
        self._fill()
        return super().__getitem__(i)

    def __iter__(self):
        self._fill()
        return super().__iter__()


UEP = _Lazy(_uep)  # rows: (bitrate, size_cu, prot_level, (L1..L4), (PI1..PI4), pad_bits)


def puncture_mask(pi: int) -> int:
    return int(_lib.load().dabgpu_tab_puncture_mask(pi))


def kept_positions(shape) -> np.ndarray:
    """indices into the 4*(nbits+6) mother code word that are transmitted, in order"""
    keep = []
    for r in shape["regions"]:
        m = puncture_mask(r["pi"])
        pos = np.arange(4 * r["steps"])
        sel = ((m >> (pos & 31)) & 1).astype(bool)
        keep.append(4 * r["step0"] + pos[sel])
    k = np.concatenate(keep).astype(np.int64)
    assert k.size == shape["in_bits"]
    return k


@functools.lru_cache(None)
def _freq_deint():
    t = np.zeros(1536, dtype=np.uint16)
    _lib.load().dabgpu_tab_freq_deint(t.ctypes.data_as(C.POINTER(C.c_uint16)))
    return t


def freq_deint() -> np.ndarray:
    return _freq_deint().copy()


def prs() -> np.ndarray:
    q = np.zeros(1536, dtype=np.uint8)
    _lib.load().dabgpu_tab_prs(q.ctypes.data_as(C.POINTER(C.c_uint8)))
    return q


def prbs(nbytes: int) -> np.ndarray:
    out = np.zeros(nbytes, dtype=np.uint8)
    _lib.load().dabgpu_tab_prbs(out.ctypes.data_as(C.POINTER(C.c_uint8)), nbytes)
    return out


def crc16(data: np.ndarray, init: int = 0xFFFF) -> int:
    """CRC-16-CCITT as misc.c:96-150 computes it (no final inversion)"""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    return int(_lib.load().dabgpu_tab_crc16(data.ctypes.data_as(C.POINTER(C.c_uint8)), data.size, init))
