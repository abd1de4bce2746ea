"""ctypes binding of libdabgpu.so (include/dabgpu.h).

Loading never needs a GPU; every compute entry point fails loudly (DabGpuError) when no sm_100
device is usable -- there is deliberately no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libdabgpu.so")

u8p = C.POINTER(C.c_uint8)


class DabGpuError(RuntimeError):
    pass


_lib = None


def load(build_if_missing: bool = True) -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
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
