"""Soft-decision mode (opt-in): symbols instead of bits from the demapper to the decoder.

The decoder's oracle is the reference's own viterbi() (unmodified viterbi.c through the port, which
tests/test_oracle_vs_ref.py pins against oracle/_ref): it already is a soft-decision decoder, the
reference just never feeds it more than three symbol values.  The back-end's oracle is the port with
orc_set_soft(1) (symbols passed through the depuncturers instead of to_viterbi())."""
import numpy as np
import pytest

from dabtools_b200 import synth

pytestmark = pytest.mark.gpu


def _soft_symbols(port, rng, n, nbits, amp, sigma, p_erase):
    data = rng.integers(0, 256, (n, nbits // 8), dtype=np.uint8)
    soft = np.empty((n, 4 * (nbits + 6)), np.uint8)
    for i in range(n):
        sym = port.encode(data[i]).astype(np.float64)
        v = 128 + (2 * sym - 1) * amp + rng.normal(0, sigma, sym.size)
        v = np.clip(np.rint(v), 121, 135)
        v[rng.random(sym.size) < p_erase] = 128
        soft[i] = v.astype(np.uint8)
    return soft, data


@pytest.mark.parametrize("nbits", [8, 104, 768, 1536, 3072, 9216])
@pytest.mark.parametrize("amp,sigma,p_erase", [(1, 0.0, 0.0), (3, 2.0, 0.25), (2, 2.5, 0.4), (0, 4.0, 0.0)])
def test_soft_viterbi_matches_the_reference_decoder(gpu, port, nbits, amp, sigma, p_erase):
    rng = np.random.default_rng(nbits + int(10 * sigma))
    n = 45 if nbits <= 3072 else 33
    soft, data = _soft_symbols(port, rng, n, nbits, amp, sigma, p_erase)
    got = gpu.viterbi_soft_batch(soft, nbits)
    want = np.stack([port.viterbi(soft[i], nbits) for i in range(n)])
    assert np.array_equal(got, want)
    if sigma == 0:
        assert np.array_equal(got, data)
    if nbits % 32 == 0:
        got_s = gpu.viterbi_soft_batch(soft, nbits, descramble=True)
        assert np.array_equal(got_s, np.stack([port.descramble(w) for w in want]))


def test_soft_viterbi_hard_alphabet_and_saturation(gpu, port):
    """127 / 128 / 129 is a special case of the soft decoder: same output as the hard-decision kernel;
    symbols outside the metric table's range saturate to 121 / 135."""
    rng = np.random.default_rng(4)
    nbits = 768
    data = rng.integers(0, 256, nbits // 8, dtype=np.uint8)
    sym = port.encode(data)
    s = sym ^ (rng.random(sym.size) < 0.06).astype(np.uint8)
    soft = (127 + 2 * s).astype(np.uint8)
    soft[rng.random(sym.size) < 0.3] = 128
    assert np.array_equal(gpu.viterbi_soft_batch(soft[None], nbits), gpu.viterbi_batch(soft[None], nbits))
    wild = np.where(s == 1, rng.integers(129, 256, s.size), rng.integers(0, 128, s.size)).astype(np.uint8)
    sat = np.clip(wild, 121, 135).astype(np.uint8)
    assert np.array_equal(gpu.viterbi_soft_batch(wild[None], nbits)[0], port.viterbi(sat, nbits))
    assert np.array_equal(gpu.soft_metrics()[:, 121:136], port.gen_metrics()[:, 121:136])


def _soft_tfs(bits, rng, amp, sigma):
    v = 128 + (2.0 * bits - 1.0) * amp + rng.normal(0, sigma, bits.shape)
    return np.clip(np.rint(v), 121, 135).astype(np.uint8)


@pytest.mark.parametrize("batch", [1, 3])
def test_soft_backend_matches_the_oracle_and_beats_hard_decisions(gpu, port, batch):
    """Back-end of a DABGPU_ENGINE_SOFT engine on noisy symbol frames: ETI byte for byte as the oracle
    with orc_set_soft(1); and the point of the mode -- on the same noisy symbols, sliced to bits, the
    hard-decision receiver makes payload errors (or loses lock) where the soft one does not."""
    ens = synth.small_ensemble()
    S, n_tf = 2, 20
    g = synth.ModeITransmitter(ens).generate(S, n_tf, seed=81, want_iq=False)
    bits = g["bits"].numpy()
    rng = np.random.default_rng(8)
    soft = _soft_tfs(bits, rng, amp=3.0, sigma=1.5)          # symbol error rate ~ 1.4 %
    eng = gpu.Engine(S, 200_000_000, gpu.ENGINE_SOFT)
    eng.set_msc_batch(batch)
    out = [[] for _ in range(S)]
    for t in range(n_tf + 1):
        n = eng.process_demapped(soft[:, t]) if t < n_tf else eng.flush()
        eti, ids = eng.fetch_eti()
        for f, s in zip(eti, ids):
            out[s].append(f.copy())
    eng.close()
    port.set_soft(True)
    try:
        want = [port.run_backend(soft[s])[0] for s in range(S)]
    finally:
        port.set_soft(False)
    nst = len(ens.subchannels)
    off = 12 + 4 * nst + 96
    for s in range(S):
        got = np.array(out[s], dtype=np.uint8).reshape(-1, 6144)
        assert got.shape == want[s].shape and want[s].shape[0] == 4 * (n_tf - 13), (s, got.shape, want[s].shape)
        assert np.array_equal(got, want[s]), s
        # soft decisions: every payload byte is right
        for f in range(got.shape[0]):
            body = got[f][off: off + ens.bytes_per_cif].tobytes()
            assert body == synth.expected_eti_payload(ens, g["payload"], s, (int(got[f][4]) - 3) % 250)
    # the same symbols sliced: the hard-decision receiver has payload errors
    hard = (soft > 128).astype(np.uint8)
    h_eti = port.run_backend(hard[0])[0]
    errs = 0
    for f in h_eti:
        want_body = synth.expected_eti_payload(ens, g["payload"], 0, (int(f[4]) - 3) % 250)
        errs += np.unpackbits(np.frombuffer(f[off: off + ens.bytes_per_cif].tobytes(), np.uint8) ^
                              np.frombuffer(want_body, np.uint8)).sum()
    assert h_eti.shape[0] < want[0].shape[0] or errs > 0


def test_soft_demapper_against_the_oracle_spectra(gpu, port):
    """Soft symbols of one frame against the same quantiser applied to the oracle's float64 DQPSK
    quotients (input_sdr.c:132-144): identical but for values that fall on a rounding boundary
    (float32 vs float64), never more than one level apart; sliced at 128 they are the reference's bits."""
    from dabtools_b200 import tables as T
    ens = synth.small_ensemble()
    g = synth.ModeITransmitter(ens).generate(1, 2, seed=5, snr_db=12)
    frame = g["iq"][0].numpy()[:393216]
    want = port.demod_frame(frame)
    got = gpu.demod_frame_soft(frame).reshape(75, 3072)
    rev = T.freq_deint().astype(np.int64)
    used = np.r_[256:1024, 1025:1793]
    d = want["symbols_d"][1:, used]                       # [75][1536] by carrier index
    q = lambda x: np.clip(np.rint(8.0 * x), -7, 7)
    exp = np.empty((75, 3072), np.int64)
    exp[:, rev] = 128 - q(d.real)
    exp[:, 1536 + rev] = 128 + q(d.imag)
    diff = np.abs(got.astype(np.int64) - exp)
    assert diff.max() <= 1
    assert (diff != 0).mean() < 1e-3
    hard = want["bits"].reshape(75, 3072)
    decided = got != 128
    assert np.array_equal((got > 128)[decided], hard[decided].astype(bool))
    assert len(np.unique(got)) == 15                      # a real spread of levels, not three values
    assert 0.2 < ((got > 122) & (got < 134)).mean() < 0.9


def test_soft_receiver_end_to_end_beats_hard_decisions(gpu, port):
    """Whole path IQ -> ETI at 7.5 dB SNR, same captures through a hard-decision engine (the
    reference's behaviour) and a DABGPU_ENGINE_SOFT engine: both lock; the soft receiver's payload
    bit error rate is at least 5 times lower."""
    ens = synth.small_ensemble()
    S, n_tf = 4, 36
    g = synth.ModeITransmitter(ens, "cuda").generate(S, n_tf, seed=91, snr_db=7.5, tail_samples=262144)
    iq = g["iq"].cpu().numpy()
    n = iq.shape[1] // 262144 * 262144
    payload = {k: v.cpu() for k, v in g["payload"].items()}
    nst = len(ens.subchannels)
    off = 12 + 4 * nst + 96

    def run(flags):
        eng = gpu.Engine(S, 200_000_000, flags)
        out = [[] for _ in range(S)]
        for pos in range(0, n, 262144):
            eng.feed_iq(iq[:, pos: pos + 262144])
            eti, ids = eng.fetch_eti()
            for f, s in zip(eti, ids):
                out[s].append(f.copy())
        eng.close()
        errs = bits = 0
        for s in range(S):
            for f in out[s]:
                want = synth.expected_eti_payload(ens, payload, s, (int(f[4]) - 3) % 250)
                errs += int(np.unpackbits(np.frombuffer(f[off: off + ens.bytes_per_cif].tobytes(), np.uint8) ^
                                          np.frombuffer(want, np.uint8)).sum())
                bits += 8 * ens.bytes_per_cif
        return errs, bits

    e_hard, b_hard = run(0)
    e_soft, b_soft = run(gpu.ENGINE_SOFT)
    assert b_soft >= b_hard > 0.5 * S * 4 * (n_tf - 16) * 8 * ens.bytes_per_cif     # both really decoded frames
    assert e_hard > 50, e_hard                                                         # the noise bites
    assert e_soft / b_soft < 0.2 * e_hard / b_hard, (e_soft, b_soft, e_hard, b_hard)
