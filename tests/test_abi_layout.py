"""include/dabgpu_ref_abi.h reproduces the reference's struct layouts byte for byte (dab.h,
input_sdr.h, sdr_fifo.h), so that an unmodified dab2eti.o can link against libdabgpu."""
import ctypes as C

import numpy as np

# values of the reference build on x86-64 (gcc), re-checked against oracle/_ref when it is present
KNOWN = dict(dab_state=1160304, sdr_state=4314856, tf=230870)


def _ours(dab):
    lib = dab.load()
    off = (C.c_int32 * 15)()
    lib.dabgpu_abi_offsets(off)
    return dict(dab_state=lib.dabgpu_sizeof_dab_state(), sdr_state=lib.dabgpu_sizeof_sdr_state(),
                tf=lib.dabgpu_sizeof_tf()), list(off)


def test_sizes_match_reference(dab, ref):
    sizes, off = _ours(dab)
    rl = ref.lib
    assert sizes == dict(dab_state=rl.ref_sizeof_dab_state(), sdr_state=rl.ref_sizeof_sdr_state(),
                         tf=rl.ref_sizeof_tf())
    roff = (C.c_int32 * 15)()
    rl.ref_abi_offsets(roff)
    assert off == list(roff)


def test_sizes_known(dab):
    sizes, _ = _ours(dab)
    assert sizes == KNOWN
