"""ctypes loaders for the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Two libraries, same call shapes:

* ``port()``  -> oracle/liboracle.so, our C restatement (oracle/dab_oracle.c); built on demand with
  gcc, available everywhere (build container and GPU box).
* ``ref()``   -> oracle/_ref/libdabref.so, the UNMODIFIED reference sources compiled in place from
  /root/reference/src plus oracle/ref_harness.c.  Only buildable in the build container; the built
  .so travels to the GPU box.  Returns None when neither the .so nor the reference tree exists.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (dabtools_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"

u8p = C.POINTER(C.c_uint8)


def _p(a: np.ndarray, ty=C.c_uint8):
    return a.ctypes.data_as(C.POINTER(ty))


class CallTrace(C.Structure):
    _fields_ = [
        ("ok", C.c_int32),
        ("coarse_timeshift", C.c_int32),
        ("fine_timeshift", C.c_int32),
        ("coarse_freq_shift", C.c_int32),
        ("fine_freq_shift", C.c_double),
        ("frequency", C.c_uint32),
        ("locked", C.c_int32),
        ("eti_frames", C.c_int32),
    ]


TRACE_DTYPE = np.dtype(
    [
        ("ok", "<i4"),
        ("coarse_timeshift", "<i4"),
        ("fine_timeshift", "<i4"),
        ("coarse_freq_shift", "<i4"),
        ("fine_freq_shift", "<f8"),
        ("frequency", "<u4"),
        ("locked", "<i4"),
        ("eti_frames", "<i4"),
    ],
    align=True,
)
assert TRACE_DTYPE.itemsize == C.sizeof(CallTrace)


def build_port(force: bool = False) -> str:
    so = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, f) for f in ("dab_oracle.c", "dab_oracle.h", "ref_shim/fftw_shim.c")]
    srcs.append(os.path.join(HERE, "..", "include", "dabgpu_tables.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    return so


def build_ref(force: bool = False):
    so = os.path.join(HERE, "_ref", "libdabref.so")
    if os.path.isdir(REF_SRC):
        # make tracks the dependencies (harness, shims, libdabgpu.so for the dab2eti drop-in build)
        subprocess.check_call(["make", "-C", HERE, "ref"] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return so if os.path.exists(so) else None


class _Common:
    """Call shapes shared by the port and the reference build."""

    kind = "?"

    # ---- Viterbi -----------------------------------------------------------------
    def encode(self, data: np.ndarray) -> np.ndarray:
        data = np.ascontiguousarray(data, dtype=np.uint8)
        out = np.empty(4 * (8 * data.size + 6), dtype=np.uint8)
        self._encode(out, data)
        return out

    def viterbi(self, symbols: np.ndarray, nbits: int) -> np.ndarray:
        symbols = np.ascontiguousarray(symbols, dtype=np.uint8)
        assert symbols.size >= 4 * (nbits + 6)
        out = np.zeros((nbits + 7) // 8, dtype=np.uint8)
        self._viterbi(symbols, out, nbits)
        return out

    def descramble(self, buf: np.ndarray) -> np.ndarray:
        b = np.array(buf, dtype=np.uint8, copy=True)
        self._descramble(_p(b), b.size)
        return b

    def check_fib_crc(self, fib: np.ndarray) -> int:
        fib = np.ascontiguousarray(fib, dtype=np.uint8)
        return int(self._check_fib_crc(_p(fib)))

    def time_deinterleave(self, cifs) -> np.ndarray:
        cifs = [np.ascontiguousarray(c, dtype=np.uint8) for c in cifs]
        assert len(cifs) == 16 and all(c.size == 55296 for c in cifs)
        arr = (u8p * 16)(*[_p(c) for c in cifs])
        out = np.empty(55296, dtype=np.uint8)
        self._time_deinterleave(_p(out), arr)
        return out

    # ---- whole-path runs ------------------------------------------------------------
    def run_backend(self, tfs: np.ndarray):
        """tfs: [n][230400] demapped hard bits -> (eti [m][6144], fibs [n][384], crc [n][12])"""
        tfs = np.ascontiguousarray(tfs, dtype=np.uint8).reshape(-1, 230400)
        n = tfs.shape[0]
        eti = np.zeros((4 * n + 4, 6144), dtype=np.uint8)
        fibs = np.zeros((n, 384), dtype=np.uint8)
        crc = np.zeros((n, 12), dtype=np.uint8)
        m = self._run_backend(_p(tfs), n, _p(eti), eti.size, _p(fibs), _p(crc))
        return eti[:m].copy(), fibs, crc

    def run_iq(self, iq: np.ndarray, chunk: int = 262144, f0: int = 200_000_000, seed: int = 1,
               want_tfs: int = 0):
        """iq: uint8 interleaved capture -> dict(eti, trace, tfs)"""
        iq = np.ascontiguousarray(iq, dtype=np.uint8).ravel()
        ncalls = iq.size // chunk
        eti = np.zeros((max(4, 4 * (iq.size // 393216 + 1)), 6144), dtype=np.uint8)
        trace = np.zeros(ncalls, dtype=TRACE_DTYPE)
        tfs = np.zeros((max(want_tfs, 1), 230400), dtype=np.uint8)
        n_calls = C.c_long(0)
        n_tfs = C.c_long(0)
        m = self._run_iq(_p(iq), iq.size, chunk, f0, seed, _p(eti), eti.size,
                         trace.ctypes.data_as(C.POINTER(CallTrace)), ncalls, C.byref(n_calls),
                         _p(tfs) if want_tfs else None, want_tfs, C.byref(n_tfs))
        return dict(eti=eti[:m].copy(), trace=trace[: n_calls.value], n_tfs=n_tfs.value,
                    tfs=tfs[: min(want_tfs, n_tfs.value)].copy())

    def run_wf(self, packets: np.ndarray) -> np.ndarray:
        """packets: uint8 [n][524] Wavefinder USB packets -> ETI [m][6144] (do_wf_decode, dab2eti.c:251-272)"""
        packets = np.ascontiguousarray(packets, dtype=np.uint8).reshape(-1, 524)
        eti = np.zeros((4 * (packets.shape[0] // 76 + 2), 6144), dtype=np.uint8)
        m = self._run_wf(_p(packets), packets.shape[0], _p(eti), eti.size)
        return eti[:m].copy()

    # ---- streaming receive loop (state kept across calls; bench.py --impl reference) -------------
    def stream_open(self, f0: int = 200_000_000, seed: int = 1):
        return self._stream_open(f0, seed)

    def stream_feed(self, h, iq: np.ndarray, chunk: int = 262144, want_eti: bool = True):
        """feed a multiple of `chunk` bytes; returns (n_frames, eti [n][6144] or None)"""
        iq = np.ascontiguousarray(iq, dtype=np.uint8).ravel()
        cap = 4 * (iq.size // 393216 + 2)
        eti = np.zeros((cap, 6144), dtype=np.uint8) if want_eti else None
        n = self._stream_feed(h, _p(iq), iq.size, chunk, _p(eti) if want_eti else None, eti.size if want_eti else 0)
        return int(n), (eti[: min(n, cap)].copy() if want_eti else None)

    def stream_locked(self, h) -> bool:
        return bool(self._stream_locked(h))

    def stream_close(self, h):
        self._stream_close(h)

    def demod_frame(self, frame: np.ndarray, force_timesync: int = 0, want_spectra: bool = True):
        frame = np.ascontiguousarray(frame, dtype=np.uint8).ravel()
        assert frame.size == 393216
        cts, fts, cfs = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        ffs = C.c_double(0)
        sym = np.zeros((76, 2048, 2), dtype=np.float64) if want_spectra else None
        symd = np.zeros((76, 2048, 2), dtype=np.float64) if want_spectra else None
        bits = np.zeros(230400, dtype=np.uint8)
        ok = self._demod_frame(_p(frame), force_timesync, C.byref(cts), C.byref(fts), C.byref(cfs),
                               C.byref(ffs), _p(sym, C.c_double) if want_spectra else None,
                               _p(symd, C.c_double) if want_spectra else None, _p(bits))
        return dict(ok=int(ok), coarse_timeshift=cts.value, fine_timeshift=fts.value,
                    coarse_freq_shift=cfs.value, fine_freq_shift=ffs.value,
                    symbols=None if sym is None else sym[..., 0] + 1j * sym[..., 1],
                    symbols_d=None if symd is None else symd[..., 0] + 1j * symd[..., 1],
                    bits=bits)


_run_iq_args = [u8p, C.c_long, C.c_int, C.c_uint32, C.c_uint, u8p, C.c_long, C.POINTER(CallTrace),
                C.c_long, C.POINTER(C.c_long), u8p, C.c_long, C.POINTER(C.c_long)]
_demod_args = [u8p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32),
               C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), u8p]


def _bind_stream(obj, lib, prefix):
    o, f, l, c = (getattr(lib, prefix + n) for n in ("open", "feed", "locked", "close"))
    o.argtypes = [C.c_uint32, C.c_uint]
    o.restype = C.c_void_p
    f.argtypes = [C.c_void_p, u8p, C.c_long, C.c_int, u8p, C.c_long]
    f.restype = C.c_long
    l.argtypes = [C.c_void_p]
    c.argtypes = [C.c_void_p]
    c.restype = None
    obj._stream_open, obj._stream_feed, obj._stream_locked, obj._stream_close = o, f, l, c


class Port(_Common):
    """oracle/liboracle.so (dab_oracle.c)"""

    kind = "port"

    def __init__(self):
        self.lib = lib = C.CDLL(build_port())
        lib.orc_encode.argtypes = [u8p, u8p, C.c_uint]
        lib.orc_viterbi.argtypes = [u8p, u8p, C.c_uint]
        lib.orc_fic_depuncture.argtypes = [u8p, u8p]
        lib.orc_uep_depuncture.argtypes = [u8p, u8p, C.c_int]
        lib.orc_eep_depuncture.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_int]
        lib.orc_descramble.argtypes = [u8p, C.c_int]
        lib.orc_check_fib_crc.argtypes = [u8p]
        lib.orc_crc16.argtypes = [u8p, C.c_int, C.c_uint16]
        lib.orc_crc16.restype = C.c_uint16
        lib.orc_time_deinterleave.argtypes = [u8p, C.POINTER(u8p)]
        lib.orc_fic_decode.argtypes = [u8p, u8p, u8p]
        lib.orc_run_backend.argtypes = [u8p, C.c_long, u8p, C.c_long, u8p, u8p]
        lib.orc_run_backend.restype = C.c_long
        lib.orc_run_iq.argtypes = _run_iq_args
        lib.orc_run_iq.restype = C.c_long
        lib.orc_demod_frame.argtypes = _demod_args
        lib.orc_gen_metrics.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_double, C.c_double, C.c_int]
        lib.orc_tab_puncture_mask.restype = C.c_uint32
        lib.orc_coarse_time_sync.argtypes = [C.POINTER(C.c_int8), C.c_int]
        lib.orc_coarse_time_sync.restype = C.c_uint32
        self._encode = lambda out, data: lib.orc_encode(_p(out), _p(data), data.size)
        self._viterbi = lambda sym, out, nbits: lib.orc_viterbi(_p(sym), _p(out), nbits)
        self._descramble = lib.orc_descramble
        self._check_fib_crc = lib.orc_check_fib_crc
        self._time_deinterleave = lib.orc_time_deinterleave
        self._run_backend = lib.orc_run_backend
        self._run_iq = lib.orc_run_iq
        self._demod_frame = lib.orc_demod_frame
        _bind_stream(self, lib, "orc_stream_")
        lib.orc_run_wf.argtypes = [u8p, C.c_long, u8p, C.c_long]
        lib.orc_run_wf.restype = C.c_long
        self._run_wf = lib.orc_run_wf

    def gen_metrics(self, amp=1, noise=1.0, bias=0.0, scale=4):
        t = np.zeros((2, 256), dtype=np.int32)
        self.lib.orc_gen_metrics(_p(t, C.c_int), amp, noise, bias, scale)
        return t

    def set_soft(self, on: bool):
        """soft-decision extension: depuncturers pass symbol values (saturated to 121..135) through"""
        self.lib.orc_set_soft(int(on))

    def fic_depuncture(self, bits2304):
        bits = np.ascontiguousarray(bits2304, dtype=np.uint8)
        out = np.empty(3096, dtype=np.uint8)
        self.lib.orc_fic_depuncture(_p(out), _p(bits))
        return out

    def uep_depuncture(self, bits, uep_index):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        out = np.empty(4 * (9216 + 6), dtype=np.uint8)
        n = self.lib.orc_uep_depuncture(_p(out), _p(bits), uep_index)
        return out[:n].copy()

    def eep_depuncture(self, bits, protlev, size, bitrate):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        out = np.empty(4 * (9216 + 6), dtype=np.uint8)
        n = self.lib.orc_eep_depuncture(_p(out), _p(bits), protlev, size, bitrate)
        return out[:n].copy()

    def fic_decode(self, fic_bits9216):
        bits = np.ascontiguousarray(fic_bits9216, dtype=np.uint8)
        fibs = np.zeros(384, dtype=np.uint8)
        crc = np.zeros(12, dtype=np.uint8)
        ok = self.lib.orc_fic_decode(_p(bits), _p(fibs), _p(crc))
        return fibs, crc, ok

    def crc16(self, data, init=0xFFFF):
        data = np.ascontiguousarray(data, dtype=np.uint8)
        return int(self.lib.orc_crc16(_p(data), data.size, init))

    def freq_deint(self):
        t = np.zeros(1536, dtype=np.uint16)
        self.lib.orc_tab_freq_deint(_p(t, C.c_uint16))
        return t

    def prs(self):
        q = np.zeros(1536, dtype=np.uint8)
        self.lib.orc_tab_prs(_p(q))
        return q

    def puncture_mask(self, pi):
        return int(self.lib.orc_tab_puncture_mask(pi))

    def shape(self, kind, a=0, b=0):
        out = np.zeros(23, dtype=np.int32)
        rc = self.lib.orc_tab_shape(kind, a, b, _p(out, C.c_int32))
        if rc:
            return None
        return dict(nbits=int(out[0]), in_bits=int(out[1]), n_regions=int(out[2]),
                    regions=out[3:].reshape(5, 4)[: int(out[2])].copy())

    def coarse_time_sync(self, real, force=0):
        real = np.ascontiguousarray(real, dtype=np.int8)
        return int(self.lib.orc_coarse_time_sync(_p(real, C.c_int8), force))

    def rand_sequence(self, seed, n):
        st = (C.c_int32 * 35)()
        self.lib.orc_srand(st, C.c_uint(seed))
        self.lib.orc_rand.restype = C.c_int
        return [self.lib.orc_rand(st) for _ in range(n)]


class _SubCh(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("id", "eepprot", "slForm", "uep_index", "eep_option", "start_cu",
                                       "size", "bitrate", "eep_protlev", "protlev", "ASCTy")]


class Ref(_Common):
    """oracle/_ref/libdabref.so: the unmodified reference (plain viterbi.c)"""

    kind = "reference"

    def __init__(self, so: str):
        self.lib = lib = C.CDLL(so)
        lib.init_viterbi()
        lib.encode.argtypes = [u8p, u8p, C.c_uint, C.c_uint, C.c_uint]
        lib.viterbi.argtypes = [C.c_void_p, u8p, u8p, C.c_uint]
        lib.fic_depuncture.argtypes = [u8p, u8p]
        lib.uep_depuncture.argtypes = [u8p, u8p, C.POINTER(_SubCh), C.POINTER(C.c_int)]
        lib.eep_depuncture.argtypes = [u8p, u8p, C.POINTER(_SubCh), C.POINTER(C.c_int)]
        lib.dab_descramble_bytes.argtypes = [u8p, C.c_int32]
        lib.check_fib_crc.argtypes = [u8p]
        lib.time_deinterleave.argtypes = [u8p, C.POINTER(u8p)]
        lib.ref_run_backend.argtypes = [u8p, C.c_long, u8p, C.c_long, u8p, u8p]
        lib.ref_run_backend.restype = C.c_long
        lib.ref_run_iq.argtypes = _run_iq_args
        lib.ref_run_iq.restype = C.c_long
        lib.ref_demod_frame.argtypes = _demod_args
        lib.gen_met.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_double, C.c_double, C.c_int]
        lib.dab_coarse_time_sync.argtypes = [C.POINTER(C.c_int8), C.POINTER(C.c_float), C.c_uint8]
        lib.dab_coarse_time_sync.restype = C.c_uint32
        self._encode = lambda out, data: lib.encode(_p(out), _p(data), data.size, 0, 0)
        self._viterbi = lambda sym, out, nbits: lib.viterbi(None, _p(sym), _p(out), nbits)
        self._descramble = lib.dab_descramble_bytes
        self._check_fib_crc = lib.check_fib_crc
        self._time_deinterleave = lib.time_deinterleave
        self._run_backend = lib.ref_run_backend
        self._run_iq = lib.ref_run_iq
        self._demod_frame = lib.ref_demod_frame
        _bind_stream(self, lib, "ref_stream_")
        lib.ref_run_wf.argtypes = [u8p, C.c_long, u8p, C.c_long]
        lib.ref_run_wf.restype = C.c_long
        self._run_wf = lib.ref_run_wf

    def gen_metrics(self, amp=1, noise=1.0, bias=0.0, scale=4):
        t = np.zeros((2, 256), dtype=np.int32)
        self.lib.gen_met(_p(t, C.c_int), amp, noise, bias, scale)
        return t

    def fic_depuncture(self, bits2304):
        bits = np.ascontiguousarray(bits2304, dtype=np.uint8)
        out = np.empty(3096, dtype=np.uint8)
        self.lib.fic_depuncture(_p(out), _p(bits))
        return out

    def uep_depuncture(self, bits, uep_index):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        out = np.empty(4 * (9216 + 6) + 64, dtype=np.uint8)
        sc = _SubCh(uep_index=uep_index)
        n = C.c_int(0)
        self.lib.uep_depuncture(_p(out), _p(bits), C.byref(sc), C.byref(n))
        return out[: n.value].copy()

    def eep_depuncture(self, bits, protlev, size, bitrate):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        out = np.empty(4 * (9216 + 6) + 64, dtype=np.uint8)
        sc = _SubCh(protlev=protlev, size=size, bitrate=bitrate)
        n = C.c_int(0)
        self.lib.eep_depuncture(_p(out), _p(bits), C.byref(sc), C.byref(n))
        return out[: n.value].copy()

    def uep_table(self):
        class Uep(C.Structure):
            _fields_ = [("bitrate", C.c_uint), ("subchsz", C.c_uint), ("protlvl", C.c_uint),
                        ("l", C.c_int * 4), ("pi", C.c_int * 4), ("padbits", C.c_int)]
        tab = (Uep * 64).in_dll(self.lib, "ueptable")
        return [(u.bitrate, u.subchsz, u.protlvl, list(u.l), list(u.pi), u.padbits) for u in tab]

    def pvec(self):
        raw = (C.c_char * (24 * 32)).in_dll(self.lib, "pvec")
        return np.frombuffer(raw, dtype=np.uint8).reshape(24, 32).copy()

    def freq_deint(self):
        raw = (C.c_uint16 * 1536).in_dll(self.lib, "rev_freq_deint_tab")
        return np.frombuffer(raw, dtype=np.uint16).copy()

    def prs(self):
        raw = (C.c_double * (1536 * 2)).in_dll(self.lib, "prs_static")
        a = np.frombuffer(raw, dtype=np.float64).reshape(1536, 2)
        return a[:, 0] + 1j * a[:, 1]

    def syms(self):
        raw = (C.c_int * 128).in_dll(self.lib, "Syms")
        return np.frombuffer(raw, dtype=np.int32).copy()

    def coarse_time_sync(self, real, force=0):
        real = np.ascontiguousarray(real, dtype=np.int8)
        filt = np.zeros(196608 - 2662, dtype=np.float32)
        return int(self.lib.dab_coarse_time_sync(_p(real, C.c_int8), _p(filt, C.c_float), force))


class RefSpiral(_Common):
    """oracle/_ref/libdabref_spiral.so: the reference built with -DENABLE_SPIRAL_VITERBI
    (viterbi_spiral.c + viterbi_spiral_sse16.c, src/Makefile:8-16).  Whole-path runs and the raw
    decoder only; it is a CPU baseline and a statistical cross-check, not an oracle: its 8-bit
    saturating metrics and tie-break differ from viterbi.c."""

    kind = "reference (Spiral SSE2 Viterbi)"

    def __init__(self, so: str):
        self.lib = lib = C.CDLL(so)
        lib.ref_run_iq.argtypes = _run_iq_args
        lib.ref_run_iq.restype = C.c_long
        lib.ref_run_backend.argtypes = [u8p, C.c_long, u8p, C.c_long, u8p, u8p]
        lib.ref_run_backend.restype = C.c_long
        lib.create_viterbi.argtypes = [C.c_int]
        lib.create_viterbi.restype = C.c_void_p
        lib.viterbi.argtypes = [C.c_void_p, u8p, u8p, C.c_int]
        self._run_iq = lib.ref_run_iq
        self._run_backend = lib.ref_run_backend
        self._vp = {}
        _bind_stream(self, lib, "ref_stream_")

    def viterbi_spiral(self, symbols: np.ndarray, nbits: int) -> np.ndarray:
        """symbols: 4*(nbits+6) bytes in the Spiral alphabet {0, 128 = erasure, 255} (depuncture.c:36-43)"""
        symbols = np.ascontiguousarray(symbols, dtype=np.uint8)
        assert symbols.size >= 4 * (nbits + 6)
        if nbits not in self._vp:
            self._vp[nbits] = self.lib.create_viterbi(nbits)
        out = np.zeros((nbits + 7) // 8 + 8, dtype=np.uint8)
        self.lib.viterbi(self._vp[nbits], _p(symbols), _p(out), nbits)
        return out[: (nbits + 7) // 8].copy()


_port = None
_ref = None
_ref_spiral = None


def ref_spiral():
    """The reference built with -DENABLE_SPIRAL_VITERBI (viterbi_spiral*.c): a secondary CPU baseline
    for whole-path runs only (run_iq / stream_*); NOT an oracle -- its 8-bit metrics and tie-break
    differ from viterbi.c, and its viterbi() takes a decoder handle."""
    global _ref_spiral
    if _ref_spiral is None:
        build_ref()
        so = os.path.join(HERE, "_ref", "libdabref_spiral.so")
        if not os.path.exists(so):
            return None
        _ref_spiral = RefSpiral(so)
    return _ref_spiral


def port() -> Port:
    global _port
    if _port is None:
        _port = Port()
    return _port


def ref():
    """The compiled reference, or None where it cannot exist (no /root/reference and no prebuilt .so)."""
    global _ref
    if _ref is None:
        so = build_ref()
        if so is None:
            return None
        _ref = Ref(so)
    return _ref
