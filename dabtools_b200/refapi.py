"""Host-side mirror of the reference's own interface for the receive path (include/dabgpu_ref_abi.h):
ctypes struct layouts identical to src/dab.h / src/input_sdr.h and thin wrappers with the reference's
function names, bound to libdabgpu.so.  The parity tests call these exactly as dab2eti.c / dab.c / fic.c /
misc.c call the reference functions."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _lib

u8p = C.POINTER(C.c_uint8)


class tf_fibs_t(C.Structure):                       # dab.h:21-25
    _fields_ = [("ok_count", C.c_uint8), ("FIB", (C.c_uint8 * 32) * 12), ("FIB_CRC_OK", C.c_uint8 * 12)]


class demapped_transmission_frame_t(C.Structure):   # dab.h:27-33
    _fields_ = [("has_fic", C.c_uint8), ("fic_symbols_demapped", (C.c_uint8 * 3072) * 3), ("fibs", tf_fibs_t),
                ("msc_filter", C.c_uint8 * 72), ("msc_symbols_demapped", (C.c_uint8 * 3072) * 72)]


class subchannel_info_t(C.Structure):               # dab.h:35-47
    _fields_ = [(n, C.c_int) for n in ("id", "eepprot", "slForm", "uep_index", "eep_option", "start_cu", "size",
                                       "bitrate", "eep_protlev", "protlev", "ASCTy")]


class tf_info_t(C.Structure):                       # dab.h:50-61
    _fields_ = [("EId", C.c_uint16), ("CIFCount_hi", C.c_uint8), ("CIFCount_lo", C.c_uint8),
                ("subchans", subchannel_info_t * 64)]


class ens_info_t(C.Structure):                      # dab.h:63-68
    _fields_ = tf_info_t._fields_


ETI_CALLBACK = C.CFUNCTYPE(None, u8p)


class dab_state_t(C.Structure):                     # dab.h:70-89
    _fields_ = [("device_type", C.c_int), ("device_state", C.c_void_p),
                ("tfs", demapped_transmission_frame_t * 5), ("tf_info", tf_info_t), ("ens_info", ens_info_t),
                ("v", C.c_void_p), ("cifs_msc", u8p * 16), ("cifs_fibs", u8p * 16), ("ncifs", C.c_int),
                ("tfidx", C.c_int), ("locked", C.c_int), ("ens_info_shown", C.c_int), ("okcount", C.c_int),
                ("eti_callback", ETI_CALLBACK)]


class CircularBuffer(C.Structure):                  # sdr_fifo.h:27-33
    _fields_ = [("size", C.c_uint32), ("start", C.c_uint32), ("count", C.c_uint32), ("elems", u8p)]


fftw_complex = C.c_double * 2


class sdr_state_t(C.Structure):                     # input_sdr.h:12-41
    _fields_ = [("frequency", C.c_uint32), ("input_buffer", C.c_uint8 * 262144), ("input_buffer_len", C.c_int),
                ("buffer", C.c_uint8 * 393216), ("coarse_timeshift", C.c_int32), ("fine_timeshift", C.c_int32),
                ("coarse_freq_shift", C.c_int32), ("fine_freq_shift", C.c_double), ("fifo", CircularBuffer),
                ("real", C.c_int8 * 196608), ("imag", C.c_int8 * 196608), ("filt", C.c_float * (196608 - 2662)),
                ("dab_frame", C.POINTER(fftw_complex)), ("prs_ifft", C.POINTER(fftw_complex)),
                ("prs_conj_ifft", C.POINTER(fftw_complex)), ("prs_syms", C.POINTER(fftw_complex)),
                ("symbols", (fftw_complex * 2048) * 76), ("symbols_d", C.POINTER(fftw_complex)),
                ("startup_delay", C.c_int32), ("force_timesync", C.c_uint8), ("p_e_prior_dep", C.c_double),
                ("p_e_prior_vitdec", C.c_double), ("p_e_after_vitdec", C.c_double)]


def _p(a, ty=C.c_uint8):
    return a.ctypes.data_as(C.POINTER(ty))


class RefApi:
    """the reference's function names, served by libdabgpu"""

    def __init__(self):
        self.lib = lib = _lib.load()
        lib.viterbi.argtypes = [C.c_void_p, u8p, u8p, C.c_uint]
        lib.fic_depuncture.argtypes = [u8p, u8p]
        lib.uep_depuncture.argtypes = [u8p, u8p, C.POINTER(subchannel_info_t), C.POINTER(C.c_int)]
        lib.eep_depuncture.argtypes = [u8p, u8p, C.POINTER(subchannel_info_t), C.POINTER(C.c_int)]
        lib.dab_descramble_bytes.argtypes = [u8p, C.c_int32]
        lib.check_fib_crc.argtypes = [u8p]
        lib.time_deinterleave.argtypes = [u8p, C.POINTER(u8p)]
        lib.fic_decode.argtypes = [C.POINTER(dab_state_t), C.POINTER(demapped_transmission_frame_t)]
        lib.init_dab_state.argtypes = [C.POINTER(C.POINTER(dab_state_t)), C.c_void_p, ETI_CALLBACK]
        lib.dab_process_frame.argtypes = [C.POINTER(dab_state_t)]
        lib.sdr_init.argtypes = [C.POINTER(sdr_state_t)]
        lib.sdr_demod.argtypes = [C.POINTER(demapped_transmission_frame_t), C.POINTER(sdr_state_t)]
        lib.dab_coarse_time_sync.argtypes = [C.POINTER(C.c_int8), C.POINTER(C.c_float), C.c_uint8]
        lib.dab_coarse_time_sync.restype = C.c_uint32
        lib.dab_fine_time_sync.argtypes = [C.POINTER(fftw_complex)]
        lib.dab_coarse_freq_sync_2.argtypes = [C.POINTER(fftw_complex)]
        lib.dab_fine_freq_corr.argtypes = [C.POINTER(fftw_complex), C.c_int32]
        lib.dab_fine_freq_corr.restype = C.c_double
        lib.init_eti.argtypes = [u8p, C.POINTER(ens_info_t)]
        lib.init_viterbi()

    def viterbi(self, symbols: np.ndarray, nbits: int) -> np.ndarray:
        symbols = np.ascontiguousarray(symbols, dtype=np.uint8)
        out = np.zeros((nbits + 7) // 8, dtype=np.uint8)
        self.lib.viterbi(None, _p(symbols), _p(out), nbits)
        return out

    def fic_depuncture(self, bits):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        out = np.zeros(3096, dtype=np.uint8)
        self.lib.fic_depuncture(_p(out), _p(bits))
        return out

    def uep_depuncture(self, bits, uep_index):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        out = np.zeros(4 * (9216 + 6), dtype=np.uint8)
        sc, n = subchannel_info_t(uep_index=uep_index), C.c_int(0)
        self.lib.uep_depuncture(_p(out), _p(bits), C.byref(sc), C.byref(n))
        return out[: n.value].copy()

    def eep_depuncture(self, bits, protlev, size, bitrate):
        bits = np.ascontiguousarray(bits, dtype=np.uint8)
        out = np.zeros(4 * (9216 + 6), dtype=np.uint8)
        sc, n = subchannel_info_t(protlev=protlev, size=size, bitrate=bitrate), C.c_int(0)
        self.lib.eep_depuncture(_p(out), _p(bits), C.byref(sc), C.byref(n))
        return out[: n.value].copy()

    def descramble(self, buf):
        b = np.array(buf, dtype=np.uint8, copy=True)
        self.lib.dab_descramble_bytes(_p(b), b.size)
        return b

    def check_fib_crc(self, fib):
        fib = np.ascontiguousarray(fib, dtype=np.uint8)
        return int(self.lib.check_fib_crc(_p(fib)))

    def time_deinterleave(self, cifs):
        cifs = [np.ascontiguousarray(c, dtype=np.uint8) for c in cifs]
        arr = (u8p * 16)(*[_p(c) for c in cifs])
        out = np.zeros(55296, dtype=np.uint8)
        self.lib.time_deinterleave(_p(out), arr)
        return out

    def run_backend(self, tfs: np.ndarray):
        """init_dab_state + dab_process_frame per TF, exactly like oracle/ref_harness.c:ref_run_backend"""
        tfs = np.ascontiguousarray(tfs, dtype=np.uint8).reshape(-1, 230400)
        frames = []
        cb = ETI_CALLBACK(lambda p: frames.append(bytes(C.cast(p, C.POINTER(C.c_uint8 * 6144)).contents)))
        dab = C.POINTER(dab_state_t)()
        self.lib.init_dab_state(C.byref(dab), None, cb)
        d = dab.contents
        d.device_type = 1
        fibs = np.zeros((tfs.shape[0], 384), np.uint8)
        crc = np.zeros((tfs.shape[0], 12), np.uint8)
        for t, src in enumerate(tfs):
            tf = d.tfs[d.tfidx]
            tf.has_fic = 1
            C.memmove(tf.fic_symbols_demapped, src.ctypes.data, 9216)
            C.memmove(tf.msc_symbols_demapped, src.ctypes.data + 9216, 221184)
            self.lib.dab_process_frame(dab)
            fibs[t] = np.frombuffer(tf.fibs.FIB, dtype=np.uint8)
            crc[t] = np.frombuffer(tf.fibs.FIB_CRC_OK, dtype=np.uint8)
        eti = np.frombuffer(b"".join(frames), dtype=np.uint8).reshape(-1, 6144)
        return eti, fibs, crc, dict(locked=d.locked, ncifs=d.ncifs, tfidx=d.tfidx)

    def run_iq(self, iq: np.ndarray, chunk: int = 262144):
        """sdr_init + (sdr_demod -> dab_process_frame) per callback, like dab2eti.c:60-130 without the
        tuner (fixed frequency)"""
        iq = np.ascontiguousarray(iq, dtype=np.uint8).ravel()
        frames, trace = [], []
        cb = ETI_CALLBACK(lambda p: frames.append(bytes(C.cast(p, C.POINTER(C.c_uint8 * 6144)).contents)))
        dab = C.POINTER(dab_state_t)()
        sdr = sdr_state_t()
        self.lib.init_dab_state(C.byref(dab), C.byref(sdr), cb)
        d = dab.contents
        d.device_type = 1
        self.lib.sdr_init(C.byref(sdr))
        for pos in range(0, iq.size - chunk + 1, chunk):
            C.memmove(sdr.input_buffer, iq.ctypes.data + pos, chunk)
            sdr.input_buffer_len = chunk
            ok = self.lib.sdr_demod(C.byref(d.tfs[d.tfidx]), C.byref(sdr))
            if ok:
                self.lib.dab_process_frame(dab)
            trace.append((ok, sdr.coarse_timeshift, sdr.fine_timeshift, sdr.coarse_freq_shift, d.locked, len(frames)))
        eti = np.frombuffer(b"".join(frames), dtype=np.uint8).reshape(-1, 6144)
        return eti, trace
