"""ctypes binding of libdabgpu.so (include/dabgpu.h).

Loading never needs a GPU; every compute entry point fails loudly (DabGpuError) when no sm_100
device is usable -- there is deliberately no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# DABGPU_LIB selects another build of the same library (kernel experiments); there is no other back-end
LIB_PATH = os.environ.get("DABGPU_LIB") or os.path.join(HERE, "libdabgpu.so")

u8p = C.POINTER(C.c_uint8)


class DabGpuError(RuntimeError):
    pass


_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if os.environ.get("DABGPU_FORBID_LOAD"):
        # bench.py's reference arm sets this: nothing of the product may be mapped into that process
        raise DabGpuError("libdabgpu.so must not be loaded in this process (DABGPU_FORBID_LOAD is set)")
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise DabGpuError(f"{LIB_PATH} is missing; run `python -m dabtools_b200.build`")
        from . import build as _b
        _b.build()
    lib = C.CDLL(LIB_PATH)
    lib.dabgpu_last_error_string.restype = C.c_char_p
    lib.dabgpu_set_stream.argtypes = [C.c_void_p]
    lib.dabgpu_tab_shape.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32)]
    lib.dabgpu_tab_uep.argtypes = [C.POINTER(C.c_int32)]
    lib.dabgpu_tab_puncture_mask.restype = C.c_uint32
    lib.dabgpu_tab_freq_deint.argtypes = [C.POINTER(C.c_uint16)]
    lib.dabgpu_tab_prs.argtypes = [u8p]
    lib.dabgpu_tab_prbs.argtypes = [u8p, C.c_int]
    lib.dabgpu_viterbi_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                         C.c_int, C.c_int]
    lib.dabgpu_fic_decode_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    lib.dabgpu_last_trellis_steps.restype = C.c_uint64
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        msg = load().dabgpu_last_error_string()
        raise DabGpuError(f"libdabgpu error {rc}: {msg.decode() if msg else '?'}")


def _np_ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def use_torch_stream():
    """Route libdabgpu launches to torch's current CUDA stream (so torch.cuda.Event times them)."""
    import torch
    load().dabgpu_set_stream(C.c_void_p(torch.cuda.current_stream().cuda_stream))


# ---- batched channel decoding -------------------------------------------------------------------
def viterbi_batch(soft: np.ndarray, nbits: int, descramble: bool = False) -> np.ndarray:
    """soft: uint8 [n][4*(nbits+6)] reference soft symbols (host) -> uint8 [n][ceil(nbits/8)]"""
    soft = np.ascontiguousarray(soft, dtype=np.uint8)
    n = soft.shape[0]
    out = np.zeros((n, (nbits + 7) // 8), dtype=np.uint8)
    check(load().dabgpu_viterbi_batch(_np_ptr(soft), soft.shape[1], n, nbits, _np_ptr(out), out.shape[1],
                                      int(descramble), 0))
    return out


def viterbi_soft_batch(soft: np.ndarray, nbits: int, descramble: bool = False) -> np.ndarray:
    """like viterbi_batch, but the symbols are weighted with the reference's metric table (soft decisions)"""
    soft = np.ascontiguousarray(soft, dtype=np.uint8)
    n = soft.shape[0]
    out = np.zeros((n, (nbits + 7) // 8), dtype=np.uint8)
    lib = load()
    lib.dabgpu_viterbi_soft_batch.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                              C.c_int, C.c_int]
    check(lib.dabgpu_viterbi_soft_batch(_np_ptr(soft), soft.shape[1], n, nbits, _np_ptr(out), out.shape[1],
                                        int(descramble), 0))
    return out


def soft_metrics() -> np.ndarray:
    t = np.zeros((2, 256), dtype=np.int32)
    load().dabgpu_tab_soft_metrics(_np_ptr(t))
    return t


def fic_decode_batch(fic_bits: np.ndarray):
    """fic_bits: uint8 [n][2304] hard bits (host) -> (fibs uint8 [n][96], crc_ok uint8 [n][3])"""
    fic_bits = np.ascontiguousarray(fic_bits, dtype=np.uint8).reshape(-1, 2304)
    n = fic_bits.shape[0]
    fibs = np.zeros((n, 96), dtype=np.uint8)
    ok = np.zeros((n, 3), dtype=np.uint8)
    check(load().dabgpu_fic_decode_batch(_np_ptr(fic_bits), n, _np_ptr(fibs), _np_ptr(ok), 0))
    return fibs, ok


def fic_decode_batch_device(fic_bits, fibs, crc_ok):
    """torch uint8 CUDA tensors: [n][2304] -> fibs [n][96], crc_ok [n][3] (in place, async)"""
    n = fic_bits.shape[0]
    assert fic_bits.is_cuda and fic_bits.is_contiguous() and fibs.is_contiguous() and crc_ok.is_contiguous()
    check(load().dabgpu_fic_decode_batch(C.c_void_p(fic_bits.data_ptr()), n, C.c_void_p(fibs.data_ptr()),
                                         C.c_void_p(crc_ok.data_ptr()), 1))


# ---- ETI consumers on the device ----------------------------------------------------------------------
ETI_BAD_SYNC, ETI_BAD_FC, ETI_BAD_HCRC, ETI_BAD_EOF_CRC, ETI_BAD_PADDING = 1, 2, 4, 8, 16


def eti_extract_subchannel(eti: np.ndarray, subchid: int, pitch: int = 4608):
    """eti: uint8 [n][6144] (host) -> (data uint8 [n][pitch], lengths int32 [n]; -1 = not carried)"""
    eti = np.ascontiguousarray(eti, dtype=np.uint8).reshape(-1, 6144)
    n = eti.shape[0]
    out = np.zeros((n, pitch), dtype=np.uint8)
    lens = np.full(n, -1, dtype=np.int32)
    lib = load()
    lib.dabgpu_eti_extract_subchannel.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p,
                                                  C.c_int]
    check(lib.dabgpu_eti_extract_subchannel(_np_ptr(eti), n, subchid, _np_ptr(out), pitch, _np_ptr(lens), 0))
    return out, lens


def eti_check(eti: np.ndarray) -> np.ndarray:
    """eti: uint8 [n][6144] (host) -> uint32 [n] masks of ETI_BAD_* (0 = consistent frame)"""
    eti = np.ascontiguousarray(eti, dtype=np.uint8).reshape(-1, 6144)
    flags = np.zeros(eti.shape[0], dtype=np.uint32)
    lib = load()
    lib.dabgpu_eti_check.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    check(lib.dabgpu_eti_check(_np_ptr(eti), eti.shape[0], _np_ptr(flags), 0))
    return flags


# ---- single-frame front-end ------------------------------------------------------------------------
def sync_frame(frame: np.ndarray, force_timesync: int = 0) -> dict:
    frame = np.ascontiguousarray(frame, dtype=np.uint8).ravel()
    assert frame.size == 393216
    out = (C.c_int32 * 4)()
    ffs = C.c_double(0)
    check(load().dabgpu_sync_frame(_np_ptr(frame), force_timesync, out, C.byref(ffs)))
    return dict(coarse_timeshift=out[0], fine_timeshift=out[1], coarse_freq_shift=out[2], ok=out[3],
                fine_freq_shift=float(ffs.value))


def demod_frame_debug(frame: np.ndarray) -> dict:
    frame = np.ascontiguousarray(frame, dtype=np.uint8).ravel()
    assert frame.size == 393216
    sym = np.zeros((76, 2048), dtype=np.complex64)
    symd = np.zeros((76, 2048), dtype=np.complex64)
    bits = np.zeros(230400, dtype=np.uint8)
    check(load().dabgpu_demod_frame_debug(_np_ptr(frame), _np_ptr(sym), _np_ptr(symd), _np_ptr(bits)))
    return dict(symbols=sym, symbols_d=symd, bits=bits)


def demod_frame_soft(frame: np.ndarray) -> np.ndarray:
    frame = np.ascontiguousarray(frame, dtype=np.uint8).ravel()
    assert frame.size == 393216
    out = np.zeros(230400, dtype=np.uint8)
    check(load().dabgpu_demod_frame_soft(_np_ptr(frame), _np_ptr(out)))
    return out


# ---- batched receiver ---------------------------------------------------------------------------------
class StreamStatus(C.Structure):
    _fields_ = [("locked", C.c_int32), ("okcount", C.c_int32), ("ncifs", C.c_int32), ("tfidx", C.c_int32),
                ("coarse_timeshift", C.c_int32), ("fine_timeshift", C.c_int32),
                ("coarse_freq_shift", C.c_int32), ("last_ok", C.c_int32),
                ("fine_freq_shift", C.c_double), ("frequency", C.c_uint32), ("n_subchannels", C.c_int32),
                ("frames_demodulated", C.c_uint64), ("eti_frames", C.c_uint64), ("fib_crc_errors", C.c_uint64)]


ENGINE_VERBOSE = 1
ENGINE_VIRTUAL_TUNER = 2
ENGINE_SOFT = 4
ENGINE_FOLLOW_RECONFIG = 8


class Engine:
    """S independent ensemble streams in lock-step (dabgpu_engine_* of include/dabgpu.h)."""

    def __init__(self, n_streams: int, tuner_hz: int = 200_000_000, flags: int = 0):
        lib = load()
        lib.dabgpu_engine_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_uint32, C.c_int]
        lib.dabgpu_engine_destroy.argtypes = [C.c_void_p]
        lib.dabgpu_engine_feed_iq.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
        lib.dabgpu_engine_process_demapped.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int]
        lib.dabgpu_engine_eti_count.argtypes = [C.c_void_p]
        lib.dabgpu_engine_eti_device.argtypes = [C.c_void_p]
        lib.dabgpu_engine_eti_device.restype = C.c_void_p
        lib.dabgpu_engine_fetch_eti.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        lib.dabgpu_engine_status.argtypes = [C.c_void_p, C.c_int, C.POINTER(StreamStatus)]
        lib.dabgpu_engine_set_seed.argtypes = [C.c_void_p, C.c_int, C.c_uint]
        lib.dabgpu_engine_trellis_steps.argtypes = [C.c_void_p]
        lib.dabgpu_engine_trellis_steps.restype = C.c_uint64
        lib.dabgpu_launch_count.restype = C.c_uint64
        self._lib = lib
        self.n_streams = n_streams
        h = C.c_void_p()
        check(lib.dabgpu_engine_create(C.byref(h), n_streams, tuner_hz, flags))
        self._h = h

    def close(self):
        if self._h:
            self._lib.dabgpu_engine_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # host buffers (numpy) --------------------------------------------------------------------
    def feed_iq(self, iq: np.ndarray) -> int:
        """iq: uint8 [n_streams][chunk_len] (host).  Returns the number of ETI frames produced."""
        iq = np.ascontiguousarray(iq, dtype=np.uint8)
        assert iq.shape[0] == self.n_streams
        check(self._lib.dabgpu_engine_feed_iq(self._h, _np_ptr(iq), iq.shape[1], iq.shape[1], 0))
        return self._lib.dabgpu_engine_eti_count(self._h)

    def process_demapped(self, tfs: np.ndarray, mask=None) -> int:
        tfs = np.ascontiguousarray(tfs, dtype=np.uint8).reshape(self.n_streams, -1)
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        check(self._lib.dabgpu_engine_process_demapped(self._h, _np_ptr(tfs), tfs.shape[1],
                                                       None if m is None else _np_ptr(m), 0))
        return self._lib.dabgpu_engine_eti_count(self._h)

    def process_wavefinder(self, packets: np.ndarray, n_packets) -> int:
        """packets: uint8 [n_streams][max_packets*524] (host), n_packets: per-stream packet counts"""
        packets = np.ascontiguousarray(packets, dtype=np.uint8).reshape(self.n_streams, -1)
        n = np.ascontiguousarray(n_packets, dtype=np.int32)
        self._lib.dabgpu_engine_process_wavefinder.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        check(self._lib.dabgpu_engine_process_wavefinder(self._h, _np_ptr(packets), packets.shape[1], _np_ptr(n)))
        return self._lib.dabgpu_engine_eti_count(self._h)

    # device buffers (torch CUDA tensors) ------------------------------------------------------
    def feed_iq_device(self, iq) -> int:
        assert iq.is_cuda and iq.dim() == 2 and iq.shape[0] == self.n_streams and iq.stride(1) == 1
        check(self._lib.dabgpu_engine_feed_iq(self._h, C.c_void_p(iq.data_ptr()), iq.stride(0), iq.shape[1], 1))
        return self._lib.dabgpu_engine_eti_count(self._h)

    def process_demapped_device(self, tfs) -> int:
        assert tfs.is_cuda and tfs.dim() == 2 and tfs.shape[0] == self.n_streams and tfs.stride(1) == 1
        check(self._lib.dabgpu_engine_process_demapped(self._h, C.c_void_p(tfs.data_ptr()), tfs.stride(0), None, 1))
        return self._lib.dabgpu_engine_eti_count(self._h)

    def eti_count(self) -> int:
        return self._lib.dabgpu_engine_eti_count(self._h)

    def fetch_eti(self, out: np.ndarray = None):
        n = self.eti_count()
        if out is None:
            out = np.empty((n, 6144), dtype=np.uint8)
        ids = np.empty(max(n, 1), dtype=np.int32)
        if n:
            got = self._lib.dabgpu_engine_fetch_eti(self._h, _np_ptr(out), _np_ptr(ids), n)
            if got < 0:
                check(got)
        return out[:n], ids[:n]

    def status(self, stream: int) -> StreamStatus:
        st = StreamStatus()
        check(self._lib.dabgpu_engine_status(self._h, stream, C.byref(st)))
        return st

    def set_seed(self, stream: int, seed: int):
        check(self._lib.dabgpu_engine_set_seed(self._h, stream, seed))

    def trellis_steps(self) -> int:
        return int(self._lib.dabgpu_engine_trellis_steps(self._h))


def launch_count() -> int:
    lib = load()
    lib.dabgpu_launch_count.restype = C.c_uint64
    return int(lib.dabgpu_launch_count())


KERNEL_NAMES = ("ingest", "fifo_read", "sync", "demod", "fic_prep", "fic_viterbi", "msc_gather", "msc_viterbi",
                "eti_pack")


def _engine_enable_timing(self, on: bool = True):
    self._lib.dabgpu_engine_enable_timing.argtypes = [C.c_void_p, C.c_int]
    check(self._lib.dabgpu_engine_enable_timing(self._h, int(on)))


def _engine_kernel_times(self) -> dict:
    ms = (C.c_double * len(KERNEL_NAMES))()
    n = (C.c_uint64 * len(KERNEL_NAMES))()
    self._lib.dabgpu_engine_kernel_times.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    self._lib.dabgpu_engine_kernel_times(self._h, ms, n, len(KERNEL_NAMES))
    return {k: dict(ms=float(ms[i]), launches=int(n[i])) for i, k in enumerate(KERNEL_NAMES)}


Engine.enable_timing = _engine_enable_timing
Engine.kernel_times = _engine_kernel_times


def _engine_host_times(self) -> dict:
    us = (C.c_double * 4)()
    self._lib.dabgpu_engine_host_times.argtypes = [C.c_void_p, C.c_void_p]
    self._lib.dabgpu_engine_host_times(self._h, us)
    return dict(zip(("pre", "wait_gpu", "fsm", "jobs"), (float(x) for x in us)))


Engine.host_times = _engine_host_times


def _engine_set_msc_batch(self, calls: int):
    self._lib.dabgpu_engine_set_msc_batch.argtypes = [C.c_void_p, C.c_int]
    check(self._lib.dabgpu_engine_set_msc_batch(self._h, calls))


def _engine_flush(self) -> int:
    self._lib.dabgpu_engine_flush.argtypes = [C.c_void_p]
    check(self._lib.dabgpu_engine_flush(self._h))
    return self._lib.dabgpu_engine_eti_count(self._h)


Engine.set_msc_batch = _engine_set_msc_batch
Engine.flush = _engine_flush


def _engine_join(self):
    self._lib.dabgpu_engine_join.argtypes = [C.c_void_p]
    check(self._lib.dabgpu_engine_join(self._h))


Engine.join = _engine_join


def _engine_submit_iq(self, iq: np.ndarray):
    """start the upload of one callback's worth of host IQ ([n_streams][chunk_len], ideally pinned)"""
    assert iq.dtype == np.uint8 and iq.flags.c_contiguous and iq.shape[0] == self.n_streams
    self._lib.dabgpu_engine_submit_iq.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    check(self._lib.dabgpu_engine_submit_iq(self._h, _np_ptr(iq), iq.shape[1], iq.shape[1]))


def _engine_feed_submitted(self) -> int:
    self._lib.dabgpu_engine_feed_submitted.argtypes = [C.c_void_p]
    check(self._lib.dabgpu_engine_feed_submitted(self._h))
    return self._lib.dabgpu_engine_eti_count(self._h)


def _engine_attach_capture(self, iq):
    """iq: torch CUDA uint8 [n_streams][len], row-contiguous; consumed in place by feed_capture()"""
    assert iq.is_cuda and iq.dim() == 2 and iq.shape[0] == self.n_streams and iq.stride(1) == 1
    self._lib.dabgpu_engine_attach_capture.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    check(self._lib.dabgpu_engine_attach_capture(self._h, C.c_void_p(iq.data_ptr()), iq.stride(0), iq.shape[1]))
    self._capture = iq   # keep it alive


def _engine_feed_capture(self, chunk_len: int) -> int:
    self._lib.dabgpu_engine_feed_capture.argtypes = [C.c_void_p, C.c_int]
    check(self._lib.dabgpu_engine_feed_capture(self._h, chunk_len))
    return self._lib.dabgpu_engine_eti_count(self._h)


def _engine_extract_subchannel(self, subchid: int, pitch: int = 4608):
    """sub-channel bytes of the last call's ETI frames, extracted on the device (eti2mpa.c:32-67)"""
    n = self.eti_count()
    out = np.zeros((max(n, 1), pitch), dtype=np.uint8)
    lens = np.full(max(n, 1), -1, dtype=np.int32)
    self._lib.dabgpu_engine_extract_subchannel.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]
    got = self._lib.dabgpu_engine_extract_subchannel(self._h, subchid, _np_ptr(out), pitch, _np_ptr(lens))
    if got < 0:
        check(got)
    return out[:n], lens[:n]


def _engine_check_eti(self) -> np.ndarray:
    n = self.eti_count()
    flags = np.zeros(max(n, 1), dtype=np.uint32)
    self._lib.dabgpu_engine_check_eti.argtypes = [C.c_void_p, C.c_void_p]
    got = self._lib.dabgpu_engine_check_eti(self._h, _np_ptr(flags))
    if got < 0:
        check(got)
    return flags[:n]


def _engine_pump(self, in_fds, out_fds=None, max_callbacks: int = -1) -> int:
    """dabgpu_engine_pump: stream every in_fds[s] through the engine into out_fds[s]; returns ETI frames written"""
    n = self.n_streams
    assert len(in_fds) == n and (out_fds is None or len(out_fds) == n)
    a = (C.c_int * n)(*in_fds)
    b = (C.c_int * n)(*out_fds) if out_fds is not None else None
    self._lib.dabgpu_engine_pump.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong]
    self._lib.dabgpu_engine_pump.restype = C.c_longlong
    got = self._lib.dabgpu_engine_pump(self._h, n, a, b, max_callbacks)
    if got < 0:
        check(int(got))
    return int(got)


def _engine_set_capture_cyclic(self, on: bool = True):
    self._lib.dabgpu_engine_set_capture_cyclic.argtypes = [C.c_void_p, C.c_int]
    check(self._lib.dabgpu_engine_set_capture_cyclic(self._h, int(on)))


def _engine_set_subchannel_mask(self, mask: int, stream: int = -1):
    self._lib.dabgpu_engine_set_subchannel_mask.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
    check(self._lib.dabgpu_engine_set_subchannel_mask(self._h, stream, mask & 0xFFFFFFFFFFFFFFFF))


Engine.set_subchannel_mask = _engine_set_subchannel_mask
Engine.attach_capture = _engine_attach_capture
Engine.set_capture_cyclic = _engine_set_capture_cyclic
Engine.pump = _engine_pump
Engine.extract_subchannel = _engine_extract_subchannel
Engine.check_eti = _engine_check_eti
Engine.feed_capture = _engine_feed_capture
Engine.submit_iq = _engine_submit_iq
Engine.feed_submitted = _engine_feed_submitted
