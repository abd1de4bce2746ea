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


def test_both_builds_of_the_table_accessors_agree():
    """tables.py reads libdabtables.so (host-only build); libdabgpu.so exports the same dabgpu_tab_*
    functions from the same header -- both must return the same values."""
    import ctypes as C
    from dabtools_b200 import lib
    L = lib.load()
    rev = np.zeros(1536, np.uint16)
    L.dabgpu_tab_freq_deint(rev.ctypes.data_as(C.POINTER(C.c_uint16)))
    assert np.array_equal(rev, T.freq_deint())
    q = np.zeros(1536, np.uint8)
    L.dabgpu_tab_prs(q.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert np.array_equal(q, T.prs())
    pr = np.zeros(1152, np.uint8)
    L.dabgpu_tab_prbs(pr.ctypes.data_as(C.POINTER(C.c_uint8)), 1152)
    assert np.array_equal(pr, T.prbs(1152))
    for pi in range(1, 25):
        assert int(L.dabgpu_tab_puncture_mask(pi)) == T.puncture_mask(pi)
    u = (C.c_int32 * (64 * 12))()
    L.dabgpu_tab_uep(u)
    assert [tuple(np.frombuffer(u, np.int32).reshape(64, 12)[i][:3]) for i in range(64)] == [r[:3] for r in T.UEP]
    for kind, a, b in [(0, 0, 0)] + [(1, i, 0) for i in range(64)] + [(2, lv, sz) for lv, sz in
                                                                     [(0, 12), (1, 8), (2, 90), (4, 27), (7, 30)]]:
        o = (C.c_int32 * 23)()
        assert L.dabgpu_tab_shape(kind, a, b, o) == 0
        sh = T.shape_fic() if kind == 0 else T.shape_uep(a) if kind == 1 else T.shape_eep(a, b)
        assert (o[0], o[1], o[2]) == (sh["nbits"], sh["in_bits"], sh["n_regions"])


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


def _steps_from_soft(soft):
    """the reference depuncturer's output (4 soft symbols per trellis step, 127/128/129) in the
    library's step-byte format: low nibble received bits, high nibble "symbol was transmitted" """
    v = soft.reshape(-1, 4)
    r = ((v > 128).astype(np.uint8) << np.arange(4, dtype=np.uint8)).sum(axis=1)
    e = ((v != 128).astype(np.uint8) << np.arange(4, 8, dtype=np.uint8)).sum(axis=1)
    return (r | e).astype(np.uint8)


def test_gather_period_tables_match_the_reference_depuncturers(port):
    """Host half of msc_gather_periods_kernel (period lists + deposit tables, run on the CPU through
    dabgpu_tab_depuncture_steps) against uep_depuncture / eep_depuncture for every profile."""
    import ctypes as C
    from dabtools_b200 import lib
    L = lib.load()
    L.dabgpu_tab_depuncture_steps.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    rng = np.random.default_rng(8)
    cases = [(1, i, 0, T.UEP[i][1]) for i in range(64)]
    cases += [(2, lv, sz, sz) for lv, sz in [(0, 12), (0, 96), (1, 8), (1, 64), (2, 6), (2, 90), (3, 4), (3, 40),
                                             (4, 27), (4, 54), (5, 21), (5, 84), (6, 18), (6, 72), (7, 15), (7, 30)]]
    for kind, a, b, size_cu in cases:
        sh = T.shape_uep(a) if kind == 1 else T.shape_eep(a, b)
        start_cu = int(rng.integers(0, 864 - size_cu + 1))
        cif = rng.integers(0, 2, 55296, dtype=np.uint8)
        steps = np.full(((sh["nbits"] + 6 + 15) // 16) * 16, 0xEE, dtype=np.uint8)
        n = L.dabgpu_tab_depuncture_steps(kind, a, b, start_cu, cif.ctypes.data, steps.ctypes.data, steps.size)
        assert n == steps.size, (kind, a, b, n)
        sub = cif[64 * start_cu:]
        if kind == 1:
            soft = port.uep_depuncture(sub, a)
        else:
            soft = port.eep_depuncture(sub, a, b, sh["nbits"] // 24)
        want = _steps_from_soft(soft)
        assert want.size == sh["nbits"] + 6
        assert np.array_equal(steps[:want.size], want), (kind, a, b)
        assert not steps[want.size:].any()          # row padding is "nothing transmitted"
