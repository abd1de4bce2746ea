"""BASELINE.json's configurations at their stated sizes, CUDA path (through the C ABI) against the
oracle and against digests of the unmodified reference's outputs.

  config 1  one ~80-TF capture of the 10-sub-channel ensemble -> ETI, bit-exact
  config 2  16384 FIC groups, 1/4 clean + 3/4 at 1/4/8 % flips -> FIBs + CRC flags, bit-exact
  config 3  1024 streams of the reference ensemble on one GPU (the benchmarked workload: capture mode /
            MSC batches of 2 and host path / batches of 4): randomly chosen streams against the oracle
  config 4  SNR x carrier-offset grid with the virtual tuner: lock rate and post-Viterbi BER against
            the oracle within stated tolerances
"""
import hashlib
import os
import sys
import zlib

import numpy as np
import pytest

from conftest import GOLDEN
from dabtools_b200 import synth

sys.path.insert(0, GOLDEN)
sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "tools"))
import make_golden_configs as mg  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "baseline_configs_v1.npz"))


def test_config2_16384_fic_groups(gpu, port, gold):
    from test_oracle_golden_configs import port_fic_groups
    bits, sent = synth.fic_groups(mg.CFG2["n_groups"], mg.CFG2["seed"])
    fibs, ok = gpu.fic_decode_batch(bits)
    want_f, want_ok = port_fic_groups(port, bits)
    assert np.array_equal(ok, want_ok)
    assert np.array_equal(fibs, want_f)
    assert ok[:4096].all() and np.array_equal(fibs[:4096], sent[:4096])
    assert 0 < (ok[12288:] == 0).sum() < 0.1 * ok[12288:].size      # the 8 % quarter really is noisy
    if zlib.crc32(bits.tobytes()) == int(gold["cfg2_in_crc"][0]):    # ... and against the reference itself
        assert np.array_equal(np.packbits(ok), gold["cfg2_ok"])
        assert hashlib.sha256(fibs.tobytes()).digest() == gold["cfg2_fibs_sha256"].tobytes()


def _engine_eti(gpu, iq, flags=0, batch=1):
    S = iq.shape[0]
    eng = gpu.Engine(S, 200_000_000, flags)
    eng.set_msc_batch(batch)
    out = [[] for _ in range(S)]
    for pos in range(0, iq.shape[1] - 262144 + 1, 262144):
        eng.feed_iq(iq[:, pos: pos + 262144])
        eti, ids = eng.fetch_eti()
        for f, s in zip(eti, ids):
            out[s].append(f.copy())
    if eng.flush():
        eti, ids = eng.fetch_eti()
        for f, s in zip(eti, ids):
            out[s].append(f.copy())
    st = [eng.status(s) for s in range(S)]
    eng.close()
    return [np.array(o, dtype=np.uint8).reshape(-1, 6144) for o in out], st


def test_config1_80_tf_capture(gpu, port, gold):
    iq = mg.cfg1_capture()
    got, _ = _engine_eti(gpu, iq[None, :])
    want = port.run_iq(iq)["eti"]
    assert got[0].shape == want.shape and want.shape[0] >= 255
    assert np.array_equal(got[0], want)
    if zlib.crc32(iq.tobytes()) == int(gold["cfg1_iq_crc"][0]):
        assert np.array_equal(mg.frame_digests(got[0]), gold["cfg1_eti_sha256"])


def test_config3_1024_streams(gpu, port):
    """The benchmarked workload at its size: S = 1024 streams of the reference ensemble, device
    resident.  Engine A consumes the capture in place with MSC batches of 2 (bench `value`), engine B
    takes the same samples from host memory with MSC batches of 4 (bench `e2e`); 8 randomly chosen
    streams of each are compared byte for byte with the oracle's receive loop."""
    import torch
    import bench
    S, n_tf = 1024, 18
    dev = torch.device("cuda", 0)
    data, ens = bench.generate_dataset(S, n_tf, dev, seed=77)
    n_calls = data.shape[1] // 262144
    rng = np.random.default_rng(3)
    chosen = sorted(rng.choice(S, 8, replace=False).tolist())

    def collect(eng, feed):
        out = {s: [] for s in chosen}
        total = 0
        for c in range(n_calls):
            n = feed(c)
            total += n
            if n:
                eti, ids = eng.fetch_eti()
                for s in chosen:
                    out[s].extend(f.copy() for f in eti[ids == s])
        if eng.flush():
            eti, ids = eng.fetch_eti()
            total += len(eti)
            for s in chosen:
                out[s].extend(f.copy() for f in eti[ids == s])
        return out, total

    eng = gpu.Engine(S)
    eng.set_msc_batch(2)
    eng.attach_capture(data)
    a, total_a = collect(eng, lambda c: eng.feed_capture(262144))
    locked = sum(eng.status(s).locked for s in range(S))
    eng.close()
    host = data.cpu().numpy()
    eng = gpu.Engine(S)
    eng.set_msc_batch(4)
    b, total_b = collect(eng, lambda c: eng.feed_iq(host[:, c * 262144:(c + 1) * 262144]))
    eng.close()
    assert locked == S and total_a == total_b and total_a >= S * 4 * (n_tf - 15)
    for s in chosen:
        want = port.run_iq(host[s, : n_calls * 262144])["eti"]
        for got in (a[s], b[s]):
            got = np.array(got, dtype=np.uint8).reshape(-1, 6144)
            assert got.shape == want.shape and want.shape[0] >= 8, (s, got.shape, want.shape)
            assert np.array_equal(got, want), s


SNRS, CFOS, PER_CELL = (8, 12, 20), (0, 400, -2300), 4


def test_config4_snr_cfo_grid(gpu, port):
    """3 SNR x 3 CFO x 4 streams, small ensemble, virtual tuner on both sides.  Sync parity is
    statistical (float32 vs float64 estimators, SURVEY 7 hard part 4), so per cell:
      * streams locked at the end: equal to the oracle's count +- 1;
      * post-Viterbi BER of the emitted frames: <= oracle BER * 1.5 + 2e-4;
    over the grid: at least 90 % of the streams byte-identical with the oracle; at 20 dB every
    stream locks and decodes without a payload bit error."""
    import torch
    from snr_cfo_sweep import payload_ber
    ens = synth.small_ensemble()
    tx = synth.ModeITransmitter(ens, "cuda")
    cells = [(snr, cfo) for snr in SNRS for cfo in CFOS]
    S, n_tf = len(cells) * PER_CELL, 44
    rows, payloads = [], []
    for s in range(S):
        snr, cfo = cells[s % len(cells)]
        g = tx.generate(1, n_tf, seed=5000 + s, snr_db=float(snr), cfo_hz=float(cfo), tail_samples=262144)
        cut = 2 * (7919 * (s + 1) % 190000)
        rows.append(g["iq"][0, cut: cut + (n_tf - 1) * 393216].cpu().numpy())
        payloads.append({k: v.cpu() for k, v in g["payload"].items()})
    n = min(r.size for r in rows) // 262144 * 262144
    iq = np.stack([r[:n] for r in rows])
    got, st = _engine_eti(gpu, iq, flags=gpu.ENGINE_VIRTUAL_TUNER)
    same = 0
    for ci, (snr, cfo) in enumerate(cells):
        ss = range(ci, S, len(cells))
        lg = lr = eg = bg = er = br = 0
        for s in ss:
            r = port.run_iq(iq[s], seed=1)
            lg += st[s].locked
            lr += int(r["trace"]["locked"][-1])
            e, b = payload_ber(ens, payloads[s], 0, got[s])
            eg, bg = eg + e, bg + b
            e, b = payload_ber(ens, payloads[s], 0, r["eti"])
            er, br = er + e, br + b
            same += int(got[s].shape == r["eti"].shape and np.array_equal(got[s], r["eti"]))
        assert abs(lg - lr) <= 1, (snr, cfo, lg, lr)
        ber_g, ber_r = eg / max(bg, 1), er / max(br, 1)
        assert ber_g <= ber_r * 1.5 + 2e-4, (snr, cfo, ber_g, ber_r)
        if snr >= 20:
            assert lg == PER_CELL and bg > 0 and eg == 0, (snr, cfo, lg, eg)
    assert same >= 0.9 * S, same


def test_sync_misses_while_locked_do_not_walk_the_slot_ring(gpu, port):
    """A locked stream whose signal is replaced by noise for 12 transmission frames: every frame in
    the burst fails dab_coarse_time_sync (19 dropped frames in a row), lock and tuner stay as they
    are.  Frames that never reach dab_process_frame must leave tfidx and the CIF window alone
    (dab2eti.c:68-71, dab.c:97) -- the engine's 14-slot ring must not advance for them -- so the
    ETI after the burst equals the reference's, also with MSC batches queued."""
    ens = synth.small_ensemble()
    S, n_tf = 2, 44
    g = synth.ModeITransmitter(ens).generate(S, n_tf, seed=31, snr_db=30, tail_samples=262144)
    iq = g["iq"].numpy().copy()
    rng = np.random.default_rng(5)
    a = 18 * 393216 + 100000
    iq[0, a: a + 12 * 393216] = np.clip(np.round(rng.normal(127, 30, 12 * 393216)), 0, 255).astype(np.uint8)
    n = iq.shape[1] // 262144 * 262144
    want = [port.run_iq(iq[s, :n]) for s in range(S)]
    tr = want[0]["trace"]
    assert (tr["coarse_timeshift"] != 0).sum() >= 12 and tr["locked"][-1] == 1      # the scenario happened
    assert set(tr["frequency"].tolist()) == {200_000_000}
    for batch in (1, 2, 4):
        got, _ = _engine_eti(gpu, iq[:, :n], batch=batch)
        for s in range(S):
            assert got[s].shape == want[s]["eti"].shape and want[s]["eti"].shape[0] >= 60, (batch, s)
            assert np.array_equal(got[s], want[s]["eti"]), (batch, s)
