"""CUDA Viterbi / FIC decode through the C ABI against the oracle: bit-exact (integer work)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _soft_batch(port, rng, n, nbits, p_flip, p_erase):
    soft = np.empty((n, 4 * (nbits + 6)), np.uint8)
    data = rng.integers(0, 256, (n, nbits // 8), dtype=np.uint8)
    for i in range(n):
        sym = port.encode(data[i])
        s = sym ^ (rng.random(sym.size) < p_flip).astype(np.uint8)
        soft[i] = 127 + 2 * s
        soft[i][rng.random(sym.size) < p_erase] = 128
    return soft, data


@pytest.mark.parametrize("nbits", [8, 24, 40, 104, 192, 768, 1000, 1536, 3072, 3080, 9216])
@pytest.mark.parametrize("p_flip,p_erase", [(0.0, 0.0), (0.03, 0.25), (0.09, 0.5), (0.5, 0.0)])
def test_viterbi_batch_matches_oracle(gpu, port, nbits, p_flip, p_erase):
    rng = np.random.default_rng(nbits * 7 + int(p_flip * 100))
    n = 70 if nbits <= 3072 else 33          # not a multiple of 32: exercises a ragged last warp
    soft, data = _soft_batch(port, rng, n, nbits, p_flip, p_erase)
    got = gpu.viterbi_batch(soft, nbits)
    want = np.stack([port.viterbi(soft[i], nbits) for i in range(n)])
    assert np.array_equal(got, want)
    if p_flip == 0 and p_erase == 0:
        assert np.array_equal(got, data)
    if nbits <= 9216 and nbits % 32 == 0:
        got_s = gpu.viterbi_batch(soft, nbits, descramble=True)
        assert np.array_equal(got_s, np.stack([port.descramble(w) for w in want]))


def test_viterbi_batch_persistent_launch_with_ragged_tail(gpu, port):
    """Enough codewords for the persistent launch shape (one 8-warp CTA per SM, static work lists:
    VitBatch::plan needs >= 8 groups per SM) with a last group of 13 lanes: every codeword must still
    come out as the oracle decodes it, whichever warp's list it landed in."""
    nbits, uniq = 768, 1300
    rng = np.random.default_rng(99)
    soft_u, _ = _soft_batch(port, rng, uniq, nbits, 0.06, 0.3)
    want_u = np.stack([port.viterbi(soft_u[i], nbits) for i in range(uniq)])
    n = 32 * 1200 + 13
    idx = rng.permutation(n) % uniq
    got = gpu.viterbi_batch(np.ascontiguousarray(soft_u[idx]), nbits)
    assert got.shape == (n, nbits // 8)
    assert np.array_equal(got, want_u[idx])


def test_viterbi_adversarial(gpu, port):
    nbits = 768
    n = 4 * (nbits + 6)
    rng = np.random.default_rng(5)
    cases = [np.full(n, 128, np.uint8), np.full(n, 127, np.uint8), np.full(n, 129, np.uint8),
             np.tile(np.array([127, 129], np.uint8), n // 2),
             np.tile(np.array([127, 128, 129, 128], np.uint8), n // 4)]
    cases += [rng.choice(np.array([127, 128, 129], np.uint8), n) for _ in range(40)]
    # tie-heavy: mostly erased input, so that many path metrics are equal
    for _ in range(40):
        s = np.full(n, 128, np.uint8)
        idx = rng.integers(0, n, 40)
        s[idx] = rng.choice(np.array([127, 129], np.uint8), idx.size)
        cases.append(s)
    soft = np.stack(cases)
    got = gpu.viterbi_batch(soft, nbits)
    want = np.stack([port.viterbi(s, nbits) for s in soft])
    assert np.array_equal(got, want)


def test_viterbi_golden(gpu):
    gold = np.load(os.path.join(GOLDEN, "reference_v1.npz"))
    for k, (nbits, _, _) in enumerate(gold["vit_meta"]):
        got = gpu.viterbi_batch(gold[f"vit_in_{k}"][None, :], int(nbits))
        assert np.array_equal(got[0], gold[f"vit_out_{k}"]), k


def test_empty_batch(gpu):
    assert gpu.viterbi_batch(np.zeros((0, 4 * 774), np.uint8), 768).shape == (0, 96)
    f, c = gpu.fic_decode_batch(np.zeros((0, 2304), np.uint8))
    assert f.shape == (0, 96) and c.shape == (0, 3)


def test_fic_decode_batch(gpu, port):
    """BASELINE config 2 in miniature: 1/4 clean, 3/4 with 1/4/8 % bit flips; bit-exact FIBs and CRCs."""
    from dabtools_b200 import synth
    ens = synth.reference_ensemble()
    g = synth.ModeITransmitter(ens).generate(1, 64, seed=3, want_iq=False)
    fic = g["bits"][0, :, :9216].numpy().reshape(-1, 2304).copy()     # 256 groups
    rng = np.random.default_rng(8)
    n = fic.shape[0]
    for q, p in enumerate((0.0, 0.01, 0.04, 0.08)):
        sl = slice(q * n // 4, (q + 1) * n // 4)
        fic[sl] ^= (rng.random(fic[sl].shape) < p).astype(np.uint8)
    fibs, ok = gpu.fic_decode_batch(fic)
    want_f = np.empty_like(fibs)
    want_ok = np.empty_like(ok)
    for t in range(n // 4):
        f, c, _ = port.fic_decode(fic[4 * t: 4 * t + 4].reshape(-1))
        want_f[4 * t: 4 * t + 4] = f.reshape(4, 96)
        want_ok[4 * t: 4 * t + 4] = c.reshape(4, 3)
    assert np.array_equal(fibs, want_f)
    assert np.array_equal(ok, want_ok)
    assert ok[: n // 4].all()
    assert np.array_equal(fibs[: n // 4].reshape(-1), g["fibs"][: n // 4].numpy().reshape(-1))


def test_fic_decode_device_pointers(gpu, port):
    import torch
    rng = np.random.default_rng(9)
    fic = rng.integers(0, 2, (96, 2304), dtype=np.uint8)
    d_in = torch.from_numpy(fic).cuda()
    d_f = torch.zeros((96, 96), dtype=torch.uint8, device="cuda")
    d_ok = torch.zeros((96, 3), dtype=torch.uint8, device="cuda")
    gpu.use_torch_stream()
    gpu.fic_decode_batch_device(d_in, d_f, d_ok)
    torch.cuda.synchronize()
    f2, ok2 = gpu.fic_decode_batch(fic)
    assert np.array_equal(d_f.cpu().numpy(), f2) and np.array_equal(d_ok.cpu().numpy(), ok2)
