"""include/dabgpu_tables.h (as exported by libdabgpu and by the oracle port) against the reference's
dab_tables.c / sdr_prstab.c (through oracle/_ref) and against the surveyor's known-answer values."""
import numpy as np
import pytest

from dabtools_b200 import tables as T


def test_known_answers(port):
    # PRBS prefix, depunctured lengths, metric table, Syms
    assert T.prbs(16).tobytes().hex() == "07be2e64129da3cf9b15238dab898880"
    assert T.shape_fic()["in_bits"] == 2304 and 4 * (T.shape_fic()["nbits"] + 6) == 3096
    assert 4 * (T.shape_uep(35)["nbits"] + 6) == 12312
    assert 4 * (T.shape_uep(63)["nbits"] + 6) == 36888
    assert T.shape_uep(35)["in_bits"] == 6140                 # 6144 with 4 pad bits
    for lvl in range(4):
        size = {0: 48, 1: 32, 2: 24, 3: 16}[lvl]              # 32 kbit/s at 1-A..4-A
        assert 4 * (T.shape_eep(lvl, size)["nbits"] + 6) == 3096
    assert 4 * (T.shape_eep(1, 8)["nbits"] + 6) == 792        # EEP 2-A 8 kbit/s special case
    mt = port.gen_metrics()
    assert mt[0, 127:130].tolist() == [3, 0, -7] and mt[1, 127:130].tolist() == [-7, 0, 3]


def test_uep_sizes_consistent():
    for i, (bitrate, size, lvl, L, PI, pad) in enumerate(T.UEP):
        sh = T.shape_uep(i)
        assert sh["nbits"] == 24 * bitrate
        assert sh["in_bits"] + pad == 64 * size, i


def test_lib_and_port_tables_agree(port):
    assert np.array_equal(T.freq_deint(), port.freq_deint())
    assert np.array_equal(T.prs(), port.prs())
    for pi in range(1, 25):
        assert T.puncture_mask(pi) == port.puncture_mask(pi)
        assert bin(T.puncture_mask(pi)).count("1") == 8 + pi


def test_against_reference_tables(ref, port):
    assert np.array_equal(T.freq_deint(), ref.freq_deint())
    q = T.prs()
    assert np.array_equal(np.array([1, 1j, -1, -1j])[q], ref.prs())
    pv = ref.pvec()
    for pi in range(1, 25):
        m = T.puncture_mask(pi)
        assert [(m >> i) & 1 for i in range(32)] == pv[pi - 1].tolist()
    for mine, (bitrate, size, lvl, l, pi0, pad) in zip(T.UEP, ref.uep_table()):
        assert mine[0] == bitrate and mine[1] == size and mine[2] == lvl and mine[5] == pad
        assert list(mine[3]) == l
        assert [p - 1 if ll else pi for p, ll, pi in zip(mine[4], l, pi0)] == pi0   # reference is 0-based
    assert np.array_equal(ref.gen_metrics(), port.gen_metrics())
    ref.viterbi(np.full(4 * 14, 128, np.uint8), 8)  # populates Syms[]
    assert ref.syms()[:8].tolist() == [0, 15, 6, 9, 13, 2, 11, 4]


def test_glibc_rand_port(port):
    import ctypes
    libc = ctypes.CDLL("libc.so.6")
    for seed in (1, 7, 12345):
        libc.srand(seed)
        want = [libc.rand() for _ in range(50)]
        assert port.rand_sequence(seed, 50) == want
