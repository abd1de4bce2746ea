"""The reference's own function signatures, served by libdabgpu (batch of one on the GPU), against the
oracle.  These are the calls dab2eti.c / dab.c / fic.c / misc.c make on this path."""
import numpy as np
import pytest

from dabtools_b200 import synth
from dabtools_b200 import tables as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(gpu):
    from dabtools_b200 import refapi
    return refapi.RefApi()


def test_viterbi_signature(api, port):
    rng = np.random.default_rng(1)
    for nbits in (768, 3072, 8, 100):
        data = rng.integers(0, 256, (nbits + 7) // 8, dtype=np.uint8)
        sym = port.encode(data)[: 4 * (nbits + 6)] if nbits % 8 == 0 else rng.integers(0, 2, 4 * (nbits + 6)).astype(np.uint8)
        soft = (127 + 2 * (sym ^ (rng.random(sym.size) < 0.04))).astype(np.uint8)
        soft[rng.random(sym.size) < 0.3] = 128
        assert np.array_equal(api.viterbi(soft, nbits), port.viterbi(soft, nbits)), nbits


def test_depuncture_signatures(api, port):
    rng = np.random.default_rng(2)
    bits = rng.integers(0, 2, 2304, dtype=np.uint8)
    assert np.array_equal(api.fic_depuncture(bits), port.fic_depuncture(bits))
    for idx in (0, 4, 15, 35, 45, 63):
        bits = rng.integers(0, 2, 64 * T.UEP[idx][1], dtype=np.uint8)
        assert np.array_equal(api.uep_depuncture(bits, idx), port.uep_depuncture(bits, idx)), idx
    for lvl, size, rate in ((0, 48, 32), (1, 8, 8), (1, 16, 16), (2, 48, 64), (3, 16, 32), (4, 27, 32), (7, 60, 128)):
        bits = rng.integers(0, 2, 64 * size, dtype=np.uint8)
        assert np.array_equal(api.eep_depuncture(bits, lvl, size, rate), port.eep_depuncture(bits, lvl, size, rate))


def test_descramble_crc_time_deinterleave(api, port):
    rng = np.random.default_rng(3)
    for n in (1, 96, 384, 1152, 2000):
        buf = rng.integers(0, 256, n, dtype=np.uint8)
        assert np.array_equal(api.descramble(buf), port.descramble(buf))
    null_fib = np.zeros(32, np.uint8)
    null_fib[0], null_fib[30], null_fib[31] = 0xFF, 0xA8, 0xA8
    assert api.check_fib_crc(null_fib) == 1
    for _ in range(8):
        fib = rng.integers(0, 256, 32, dtype=np.uint8)
        assert api.check_fib_crc(fib) == port.check_fib_crc(fib)
    cifs = [rng.integers(0, 2, 55296, dtype=np.uint8) for _ in range(16)]
    assert np.array_equal(api.time_deinterleave(cifs), port.time_deinterleave(cifs))


def test_dab_process_frame_sequence(api, port):
    ens = synth.small_ensemble()
    g = synth.ModeITransmitter(ens).generate(1, 16, seed=21, want_iq=False)
    bits = g["bits"][0].numpy().copy()
    rng = np.random.default_rng(4)
    bits ^= (rng.random(bits.shape) < 0.02).astype(np.uint8)
    eti, fibs, crc, st = api.run_backend(bits)
    want_eti, want_fibs, want_crc = port.run_backend(bits)
    assert np.array_equal(fibs, want_fibs) and np.array_equal(crc, want_crc)
    assert eti.shape == want_eti.shape == (12, 6144) and np.array_equal(eti, want_eti)
    assert st["locked"] == 1


def test_sdr_demod_loop(api, port):
    ens = synth.small_ensemble()
    g = synth.ModeITransmitter(ens).generate(1, 17, seed=22, snr_db=30, tail_samples=262144)
    iq = g["iq"][0].numpy()[2 * 40000:]
    n = iq.size // 262144 * 262144
    eti, trace = api.run_iq(iq[:n])
    want = port.run_iq(iq[:n])
    wt = [(int(a["ok"]), int(a["coarse_timeshift"]), int(a["fine_timeshift"]), int(a["coarse_freq_shift"]),
           int(a["locked"]), int(a["eti_frames"])) for a in want["trace"]]
    assert trace == wt
    assert eti.shape == want["eti"].shape and eti.shape[0] >= 4 and np.array_equal(eti, want["eti"])


def test_sync_signatures(api, port, ref_or_none=None):
    """sdr_sync.h entry points with the reference's argument types (int8 / double-complex arrays)"""
    import ctypes as C
    from dabtools_b200 import refapi
    ens = synth.small_ensemble()
    g = synth.ModeITransmitter(ens).generate(1, 2, seed=23, snr_db=25, cfo_hz=1000.0)
    iq = g["iq"][0].numpy()
    frame = iq[:393216]
    want = port.demod_frame(frame)
    real = (frame[0::2].astype(np.int16) - 127).astype(np.int8)
    imag = (frame[1::2].astype(np.int16) - 127).astype(np.int8)
    filt = np.zeros(196608 - 2662, np.float32)
    lib = api.lib
    assert lib.dab_coarse_time_sync(real.ctypes.data_as(C.POINTER(C.c_int8)), filt.ctypes.data_as(C.POINTER(C.c_float)), 0) == 0
    mis = iq[2 * 60000: 2 * 60000 + 393216]
    real2 = (mis[0::2].astype(np.int16) - 127).astype(np.int8)
    assert lib.dab_coarse_time_sync(real2.ctypes.data_as(C.POINTER(C.c_int8)), filt.ctypes.data_as(C.POINTER(C.c_float)), 0) \
        == port.coarse_time_sync(real2)
    cframe = np.stack([real.astype(np.float64), imag.astype(np.float64)], axis=1).copy()
    pf = cframe.ctypes.data_as(C.POINTER(refapi.fftw_complex))
    assert lib.dab_fine_time_sync(pf) == want["fine_timeshift"]
    assert abs(lib.dab_fine_freq_corr(pf, 0) - want["fine_freq_shift"]) < 0.05
    # the coarse frequency estimator takes the fftshifted spectrum of the symbol at 2656+505+fine_timeshift
    start = 2656 + 505 + want["fine_timeshift"]
    x = cframe[start:start + 2048, 0] + 1j * cframe[start:start + 2048, 1]
    spec = np.fft.fftshift(np.fft.fft(x))
    sp = np.stack([spec.real, spec.imag], axis=1).copy()
    assert lib.dab_coarse_freq_sync_2(sp.ctypes.data_as(C.POINTER(refapi.fftw_complex))) == want["coarse_freq_shift"] == 1


def test_c_host_program_is_a_drop_in(gpu, port, tmp_path):
    """examples/dab2eti_file.c: a plain C host (the reference's call sequence, gcc, only include/*.h and
    -ldabgpu) turns a capture file into the same ETI bytes as the oracle."""
    import os
    import subprocess
    from conftest import ROOT
    exe = tmp_path / "dab2eti_file"
    libdir = os.path.join(ROOT, "dabtools_b200")
    subprocess.check_call(["gcc", "-O2", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "dab2eti_file.c"), "-L" + libdir, "-ldabgpu",
                           "-Wl,-rpath," + libdir, "-o", str(exe)])
    ens = synth.small_ensemble()
    g = synth.ModeITransmitter(ens).generate(1, 17, seed=24, snr_db=30, tail_samples=262144)
    iq = g["iq"][0].numpy()
    n = iq.size // 262144 * 262144
    cap = tmp_path / "capture.iq"
    iq[:n].tofile(cap)
    out = subprocess.run([str(exe), str(cap)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
    eti = np.frombuffer(out, dtype=np.uint8).reshape(-1, 6144)
    want = port.run_iq(iq[:n])["eti"]
    assert eti.shape == want.shape and eti.shape[0] >= 8
    assert np.array_equal(eti, want)


def test_unmodified_reference_dab2eti_links_and_runs_on_libdabgpu(gpu, port, tmp_path):
    """oracle/_ref/dab2eti_gpu is the reference's own src/dab2eti.c, unmodified and compiled against the
    reference's own headers, linked with libdabgpu.so INSTEAD of the reference's objects (plus a
    file-backed librtlsdr, oracle/ref_shim/rtlsdr_file.c; see oracle/Makefile).  Same capture in,
    same ETI bytes out as the same dab2eti.c linked with the reference's objects (dab2eti_ref) and
    as the oracle.  Built in the build container (needs /root/reference); skipped where absent."""
    import os
    import subprocess
    from conftest import ROOT
    exe_gpu = os.path.join(ROOT, "oracle", "_ref", "dab2eti_gpu")
    exe_ref = os.path.join(ROOT, "oracle", "_ref", "dab2eti_ref")
    if not os.path.exists(exe_gpu):
        pytest.skip("oracle/_ref/dab2eti_gpu was not built (no /root/reference at build time)")
    ens = synth.small_ensemble()
    g = synth.ModeITransmitter(ens).generate(1, 19, seed=25, snr_db=30, tail_samples=262144)
    iq = g["iq"][0].numpy()[2 * 4321:]
    n = iq.size // 262144 * 262144
    cap = tmp_path / "capture.iq"
    iq[:n].tofile(cap)
    env = dict(os.environ, RTLSDR_FILE=str(cap))
    r = subprocess.run([exe_gpu, "200000000"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, env=env, timeout=300)
    assert r.returncode == 0, r.stderr.decode()[-2000:]
    eti = np.frombuffer(r.stdout, dtype=np.uint8).reshape(-1, 6144)
    want = port.run_iq(iq[:n])["eti"]
    assert eti.shape == want.shape and eti.shape[0] >= 12
    assert np.array_equal(eti, want)
    assert b"Locked" in r.stderr and b"ENSEMBLE_INFO" in r.stderr      # the reference's own diagnostics
    if os.path.exists(exe_ref):
        r2 = subprocess.run([exe_ref, "200000000"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, env=env,
                            timeout=300)
        assert r2.returncode == 0 and r2.stdout == r.stdout


def test_spiral_mode_callers(api, port):
    """a20: a host built with -DENABLE_SPIRAL_VITERBI calls create_viterbi() and passes the Spiral
    alphabet {0, 128 = erasure, 255} (depuncture.c:36-43, dab.c:27-30).  libdabgpu serves that
    signature with the same decoder (viterbi.c's maximum-likelihood decisions): identical to the
    reference's Spiral SSE2 decoder on clean input, identical to viterbi.c on the same hard decisions
    always, and in agreement with the Spiral decoder on most noisy code words (its 8-bit metrics and
    tie-break differ, SURVEY 3.4 -- it is a CPU baseline, not an oracle)."""
    import ctypes as C
    from oracle import oracle
    lib = api.lib
    lib.create_viterbi.restype = C.c_void_p
    lib.create_viterbi.argtypes = [C.c_int]
    assert lib.create_viterbi(768)
    sp = oracle.ref_spiral()
    rng = np.random.default_rng(12)
    same_sp = n = 0
    for p_flip in (0.0, 0.0, 0.04, 0.04, 0.04, 0.04, 0.04, 0.04):
        for nbits in (768, 3072):
            data = rng.integers(0, 256, nbits // 8, dtype=np.uint8)
            sym = port.encode(data)
            s = sym ^ (rng.random(sym.size) < p_flip).astype(np.uint8)
            erase = rng.random(sym.size) < 0.25
            soft_s = (255 * s).astype(np.uint8)
            soft_s[erase] = 128
            soft_k = (127 + 2 * s).astype(np.uint8)
            soft_k[erase] = 128
            got = api.viterbi(soft_s, nbits)
            assert np.array_equal(got, port.viterbi(soft_k, nbits))
            if p_flip == 0:
                assert np.array_equal(got, data)
            if sp is not None:
                ds = sp.viterbi_spiral(soft_s, nbits)
                if p_flip == 0:
                    assert np.array_equal(got, ds)
                n += 1
                same_sp += np.array_equal(got, ds)
    if sp is not None:
        assert same_sp >= 0.75 * n, (same_sp, n)


def test_viterbi_signature_with_real_soft_symbols(api, port):
    """a caller of the reference-signature viterbi() who passes real soft values gets the
    reference's soft-decision decode (its own metric table), not a sliced one."""
    rng = np.random.default_rng(31)
    nbits = 768
    data = rng.integers(0, 256, nbits // 8, dtype=np.uint8)
    sym = port.encode(data).astype(np.float64)
    v = np.clip(np.rint(128 + (2 * sym - 1) * 3 + rng.normal(0, 2.5, sym.size)), 121, 135).astype(np.uint8)
    got = api.viterbi(v, nbits)
    assert np.array_equal(got, port.viterbi(v, nbits))
    sliced = (127 + 2 * (v > 128)).astype(np.uint8)
    sliced[v == 128] = 128
    assert not np.array_equal(port.viterbi(sliced, nbits), got) or np.array_equal(got, data)
