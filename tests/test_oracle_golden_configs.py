"""The oracle port at the sizes BASELINE.json states, against digests of the unmodified reference's
outputs (tests/golden/baseline_configs_v1.npz, made by tests/golden/make_golden_configs.py)."""
import hashlib
import os
import sys
import zlib

import numpy as np
import pytest

from conftest import GOLDEN

sys.path.insert(0, GOLDEN)
import make_golden_configs as mg  # noqa: E402  (the committed generator: same seeds, same helpers)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "baseline_configs_v1.npz"))


def port_fic_groups(port, bits):
    n = bits.shape[0]
    fibs = np.empty((n, 96), np.uint8)
    ok = np.empty((n, 3), np.uint8)
    for g in range(0, n, 4):
        f, c, _ = port.fic_decode(bits[g:g + 4].reshape(-1))
        fibs[g:g + 4] = f.reshape(4, 96)
        ok[g:g + 4] = c.reshape(4, 3)
    return fibs, ok


def test_config2_16384_fic_groups(gold, port):
    from dabtools_b200 import synth
    bits, sent = synth.fic_groups(mg.CFG2["n_groups"], mg.CFG2["seed"])
    assert zlib.crc32(bits.tobytes()) == int(gold["cfg2_in_crc"][0])
    fibs, ok = port_fic_groups(port, bits)
    assert np.array_equal(np.packbits(ok), gold["cfg2_ok"])
    got = [zlib.crc32(fibs[i:i + 1024].tobytes()) for i in range(0, fibs.shape[0], 1024)]
    assert got == gold["cfg2_fibs_crc32_per_1024"].tolist()
    assert hashlib.sha256(fibs.tobytes()).digest() == gold["cfg2_fibs_sha256"].tobytes()
    assert np.array_equal(fibs[:4096], sent[:4096]) and ok[:4096].all()


def test_config1_80_tf_capture(gold, port):
    iq = mg.cfg1_capture()
    if zlib.crc32(iq.tobytes()) != int(gold["cfg1_iq_crc"][0]):
        pytest.skip("synthetic capture differs in the last bit on this torch/numpy build")
    r = port.run_iq(iq)
    tr = r["trace"]
    got = np.stack([tr["ok"], tr["coarse_timeshift"], tr["fine_timeshift"], tr["coarse_freq_shift"],
                    tr["locked"], tr["eti_frames"]], axis=1)
    assert np.array_equal(got, gold["cfg1_trace_int"])
    assert r["eti"].shape[0] == gold["cfg1_eti_sha256"].shape[0] >= 255
    assert np.array_equal(mg.frame_digests(r["eti"]), gold["cfg1_eti_sha256"])
