"""Host-side control logic exported by libdabgpu (no GPU involved): FIG parsing, ensemble merge and
the ETI header builder against the compiled reference (oracle/_ref) on randomised, well-formed input.
fic.c:47-147, misc.c:14-27, misc.c:153-213."""
import ctypes as C

import numpy as np
import pytest

from dabtools_b200 import refapi as R


def _fig00(rng):
    return bytes([0x05, 0x00, rng.integers(256), rng.integers(256), rng.integers(32), rng.integers(250)])


def _fig01(rng, budget):
    body = b""
    while len(body) + 4 <= budget - 2 and rng.random() < 0.8:
        scid, start = int(rng.integers(64)), int(rng.integers(864))
        if rng.random() < 0.5:      # short form: UEP table index
            body += bytes([scid << 2 | start >> 8, start & 0xFF, int(rng.integers(64))])
        else:                       # long form: option 0/1 (the others index past eeptable[] in the reference)
            opt, lvl, size = int(rng.integers(2)), int(rng.integers(4)), int(rng.integers(1, 864))
            body += bytes([scid << 2 | start >> 8, start & 0xFF, 0x80 | opt << 4 | lvl << 2 | size >> 8, size & 0xFF])
    return bytes([len(body) + 1, 0x01]) + body


def _fig02(rng, budget):
    pd = int(rng.random() < 0.3)
    body = b""
    while len(body) + 9 <= budget - 2 and rng.random() < 0.8:
        sid = bytes(rng.integers(0, 256, 4 if pd else 2, dtype=np.uint8))
        ncomp = int(rng.integers(1, 3))
        comps = b""
        for _ in range(ncomp):
            tmid = int(rng.choice([0, 0, 0, 3]))    # (1 and 2 only make the reference print)
            comps += bytes([tmid << 6 | int(rng.integers(64)), int(rng.integers(64)) << 2 | int(rng.integers(4))])
        body += sid + bytes([int(rng.integers(16)) << 4 | ncomp]) + comps
    return bytes([len(body) + 1, pd << 5 | 0x02]) + body


def _other_fig(rng, budget):
    n = int(rng.integers(1, min(budget - 1, 12) + 1))
    return bytes([int(rng.integers(1, 8)) << 5 | n]) + bytes(rng.integers(0, 256, n, dtype=np.uint8))


def _random_fib(rng):
    fib = b""
    while len(fib) < 26 and rng.random() < 0.9:
        budget = 30 - len(fib)
        kind = rng.integers(4)
        fig = (_fig00(rng) if kind == 0 and budget >= 6 else _fig01(rng, budget) if kind == 1 and budget >= 6
               else _fig02(rng, budget) if kind == 2 and budget >= 11 else _other_fig(rng, budget))
        if len(fib) + len(fig) > 30:
            break
        fib += fig
    fib += b"\xff" * (30 - len(fib))
    return fib + b"\x00\x00"


def _fibs(rng):
    f = R.tf_fibs_t()
    for i in range(12):
        raw = _random_fib(rng)
        for j in range(32):
            f.FIB[i][j] = raw[j]
        f.FIB_CRC_OK[i] = int(rng.random() < 0.8)
    f.ok_count = sum(f.FIB_CRC_OK)
    return f


@pytest.fixture(scope="module")
def libs(dab, ref):
    ours = dab.load()
    return ours, ref.lib


def test_fib_decode_and_merge_info_match_the_reference(libs):
    ours, theirs = libs
    rng = np.random.default_rng(2024)
    ens_a, ens_b = R.ens_info_t(), R.ens_info_t()
    for e in (ens_a, ens_b):       # init as dab.c:23-25
        C.memset(C.byref(e), 0, C.sizeof(e))
        for i in range(64):
            e.subchans[i].id = e.subchans[i].ASCTy = -1
        e.CIFCount_hi = e.CIFCount_lo = 0xFF
    for trial in range(300):
        fibs = _fibs(rng)
        a, b = R.tf_info_t(), R.tf_info_t()
        ours.fib_decode(C.byref(a), C.byref(fibs), 12)
        theirs.fib_decode(C.byref(b), C.byref(fibs), 12)
        assert bytes(a) == bytes(b), trial
        ours.merge_info(C.byref(ens_a), C.byref(a))
        theirs.merge_info(C.byref(ens_b), C.byref(b))
        assert bytes(ens_a) == bytes(ens_b), trial


def test_init_eti_matches_the_reference(libs):
    ours, theirs = libs
    rng = np.random.default_rng(7)
    from dabtools_b200 import tables as T
    for trial in range(200):
        e = R.ens_info_t()
        C.memset(C.byref(e), 0, C.sizeof(e))
        for i in range(64):
            e.subchans[i].id = e.subchans[i].ASCTy = -1
        e.EId = int(rng.integers(65536))
        e.CIFCount_hi, e.CIFCount_lo = int(rng.integers(20)), int(rng.integers(250))
        for scid in rng.choice(64, size=int(rng.integers(0, 13)), replace=False):
            sc = e.subchans[int(scid)]
            sc.id = int(scid)
            sc.start_cu = int(rng.integers(864))
            if rng.random() < 0.5:
                idx = int(rng.integers(64))
                sc.slForm = sc.eepprot = 0
                sc.uep_index, sc.size, sc.bitrate, sc.protlev = idx, T.UEP[idx][1], T.UEP[idx][0], T.UEP[idx][2]
            else:
                sc.slForm = sc.eepprot = 1
                sc.protlev = int(rng.integers(8))
                sc.bitrate = 8 * int(rng.integers(1, 48))
                sc.size = int(rng.integers(1, 400))
        a = (C.c_uint8 * 6144)()
        b = (C.c_uint8 * 6144)()
        e2 = R.ens_info_t.from_buffer_copy(bytes(e))
        na = ours.init_eti(a, C.byref(e))
        nb = theirs.init_eti(b, C.byref(e2))
        assert na == nb and bytes(a[:na]) == bytes(b[:nb]), trial
        assert bytes(e) == bytes(e2)


def test_fifo_drop_ins_match_the_reference(libs):
    """cbInit/cbWrite/cbRead/cbIsEmpty/cbIsFull/sdr_read_fifo (sdr_fifo.c:26-61): same buffer contents,
    same FIFO state, including the stale tail of negative shifts and the parked skip bytes of positive
    ones (the quirks the batched FIFO bookkeeping reproduces on the GPU side)."""
    ours, theirs = libs
    rng = np.random.default_rng(99)
    u8p = C.POINTER(C.c_uint8)
    for L in (ours, theirs):
        L.cbInit.argtypes = [C.POINTER(R.CircularBuffer), C.c_uint32]
        L.cbWrite.argtypes = [C.POINTER(R.CircularBuffer), u8p]
        L.cbRead.argtypes = [C.POINTER(R.CircularBuffer), u8p]
        L.cbIsEmpty.argtypes = [C.POINTER(R.CircularBuffer)]
        L.cbIsFull.argtypes = [C.POINTER(R.CircularBuffer)]
        L.sdr_read_fifo.argtypes = [C.POINTER(R.CircularBuffer), C.c_uint32, C.c_int32, u8p]
    size, frame = 4096, 1000
    fa, fb = R.CircularBuffer(), R.CircularBuffer()
    ours.cbInit(C.byref(fa), size)
    theirs.cbInit(C.byref(fb), size)
    buf_a = (C.c_uint8 * frame)()
    buf_b = (C.c_uint8 * frame)()
    for trial in range(60):
        n = int(rng.integers(200, 1500))
        n = min(n, size - fa.count)               # (an overflow only makes the reference print)
        data = rng.integers(0, 256, n, dtype=np.uint8)
        for v in data:
            x = C.c_uint8(int(v))
            ours.cbWrite(C.byref(fa), C.byref(x))
            theirs.cbWrite(C.byref(fb), C.byref(x))
        assert (fa.start, fa.count) == (fb.start, fb.count)
        assert ours.cbIsFull(C.byref(fa)) == theirs.cbIsFull(C.byref(fb))
        shift = int(rng.choice([0, 0, 2, 16, 300, -2, -20, -400]))
        need = frame + max(shift, 0)
        if fa.count >= need:
            ours.sdr_read_fifo(C.byref(fa), frame, shift, buf_a)
            theirs.sdr_read_fifo(C.byref(fb), frame, shift, buf_b)
            assert bytes(buf_a) == bytes(buf_b), (trial, shift)
            assert (fa.start, fa.count) == (fb.start, fb.count)
        assert ours.cbIsEmpty(C.byref(fa)) == theirs.cbIsEmpty(C.byref(fb))
    x, y = C.c_uint8(0), C.c_uint8(0)
    while fa.count:
        ours.cbRead(C.byref(fa), C.byref(x))
        theirs.cbRead(C.byref(fb), C.byref(y))
        assert x.value == y.value
    assert fb.count == 0
