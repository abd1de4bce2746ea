"""Back-end on the GPU (FIC decode -> lock FSM -> time de-interleave -> depuncture -> Viterbi ->
descramble -> ETI assembly) against the oracle: ETI frames bit-exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from dabtools_b200 import synth

pytestmark = pytest.mark.gpu


def _run_engine(gpu, bits, msc_batch=1):
    """bits: [S][n_tf][230400] -> list of per-stream ETI arrays"""
    S, n_tf = bits.shape[0], bits.shape[1]
    eng = gpu.Engine(S)
    eng.set_msc_batch(msc_batch)
    out = [[] for _ in range(S)]
    for t in range(n_tf + 1):
        n = eng.process_demapped(bits[:, t]) if t < n_tf else eng.flush()
        eti, ids = eng.fetch_eti()
        assert eti.shape[0] == n
        for f, s in zip(eti, ids):
            out[s].append(f.copy())
    st = [eng.status(s) for s in range(S)]
    eng.close()
    return [np.array(o, dtype=np.uint8).reshape(-1, 6144) for o in out], st


@pytest.mark.parametrize("ens_name,flip", [("small", 0.0), ("small", 0.03), ("reference", 0.0), ("reference", 0.05)])
def test_backend_eti_matches_oracle(gpu, port, ens_name, flip):
    ens = synth.small_ensemble() if ens_name == "small" else synth.reference_ensemble()
    S, n_tf = 3, 18
    g = synth.ModeITransmitter(ens).generate(S, n_tf, seed=11, want_iq=False)
    bits = g["bits"].numpy().copy()
    if flip:
        rng = np.random.default_rng(1)
        bits ^= (rng.random(bits.shape) < flip).astype(np.uint8)
        bits[1, 14, :9216] ^= (rng.random(9216) < 0.3).astype(np.uint8)   # stream 1 loses lock at TF 14
    got, st = _run_engine(gpu, bits)
    for s in range(S):
        want, _, _ = port.run_backend(bits[s])
        assert got[s].shape == want.shape, (s, got[s].shape, want.shape)
        assert np.array_equal(got[s], want), s
    if not flip:
        assert all(x.shape[0] == 4 * (n_tf - 13) for x in got)
        assert all(x.locked == 1 for x in st)
        # and the decoded payload is what was transmitted
        nst = len(ens.subchannels)
        off = 12 + 4 * nst + 96
        body = got[0][0][off: off + ens.bytes_per_cif].tobytes()
        assert body == synth.expected_eti_payload(ens, g["payload"], 0, 36)


@pytest.mark.parametrize("batch", [2, 3, 4])
def test_deferred_msc_batches_give_identical_frames(gpu, port, batch):
    """MSC decoding lagging by up to 4 TFs (one Viterbi launch per batch) must not change a byte,
    including across a lock loss while frames are still queued."""
    ens = synth.reference_ensemble()
    S, n_tf = 2, 24
    g = synth.ModeITransmitter(ens).generate(S, n_tf, seed=31, want_iq=False)
    bits = g["bits"].numpy().copy()
    rng = np.random.default_rng(2)
    bits ^= (rng.random(bits.shape) < 0.03).astype(np.uint8)
    bits[1, 17, :9216] ^= (rng.random(9216) < 0.3).astype(np.uint8)
    got, _ = _run_engine(gpu, bits, msc_batch=batch)
    for s in range(S):
        want, _, _ = port.run_backend(bits[s])
        assert got[s].shape == want.shape and np.array_equal(got[s], want), (batch, s)


def test_backend_golden(gpu):
    gold = np.load(os.path.join(GOLDEN, "reference_v1.npz"))
    bits = np.unpackbits(gold["be_bits"]).reshape(1, 15, 230400)
    got, _ = _run_engine(gpu, bits)
    assert np.array_equal(got[0], gold["be_eti"])


def test_masked_streams_and_ragged_progress(gpu, port):
    """streams advance independently: stream 1 only receives every other call"""
    ens = synth.small_ensemble()
    S, n_tf = 2, 34
    g = synth.ModeITransmitter(ens).generate(S, n_tf, seed=12, want_iq=False)
    bits = g["bits"].numpy()
    eng = gpu.Engine(S)
    out = [[] for _ in range(S)]
    fed = [0, 0]
    for call in range(n_tf):
        mask = np.array([1, call % 2], dtype=np.uint8)
        frame = np.stack([bits[0, fed[0]], bits[1, fed[1]]])
        eng.process_demapped(frame, mask)
        fed[0] += 1
        fed[1] += int(mask[1])
        eti, ids = eng.fetch_eti()
        for f, s in zip(eti, ids):
            out[s].append(f.copy())
    eng.close()
    for s in range(S):
        want, _, _ = port.run_backend(bits[s, : fed[s]])
        assert np.array_equal(np.array(out[s]).reshape(-1, 6144), want)


def _profile_ensembles():
    """Ensembles that together use every UEP profile (64) and every EEP level with several sizes,
    i.e. every puncturing index 1..24 in every region position the standard allows."""
    from dabtools_b200 import tables as T
    sizes = [T.UEP[i][1] for i in range(64)]
    bins = []
    for i in sorted(range(64), key=lambda i: -sizes[i]):
        for b in bins:
            if b["cu"] + sizes[i] <= 864 and len(b["idx"]) < 10 and b["bytes"] + T.shape_uep(i)["nbits"] // 8 <= 5400:
                break
        else:
            b = {"idx": [], "cu": 0, "bytes": 0}
            bins.append(b)
        b["idx"].append(i)
        b["cu"] += sizes[i]
        b["bytes"] += T.shape_uep(i)["nbits"] // 8
    out = []
    for b in bins:
        subs, cu = [], 0
        for k, i in enumerate(b["idx"]):
            subs.append(synth.SubChannel(id=3 * k + 1, start_cu=cu, uep_index=i))
            cu += sizes[i]
        out.append(synth.Ensemble(subs))
    # EEP: (level, size in CU); the second ensemble has the odd sizes
    eep_a = [(0, 12), (1, 8), (2, 6), (3, 4), (4, 27), (5, 21), (6, 18), (7, 15), (0, 96), (2, 48)]
    eep_b = [(1, 64), (3, 40), (4, 54), (5, 84), (6, 72), (7, 30), (2, 90), (0, 24)]
    for spec in (eep_a, eep_b):
        subs, cu = [], 0
        for k, (lv, sz) in enumerate(spec):
            subs.append(synth.SubChannel(id=5 * k + 2, start_cu=cu, eep_level=lv, size_cu=sz))
            cu += sz
        out.append(synth.Ensemble(subs))
    return out


def test_every_protection_profile(gpu, port):
    """One stream per ensemble, all in one engine (different multiplex layouts side by side): the
    gather's per-region deposit tables and the Viterbi's length classes against the oracle for all
    64 UEP profiles and all 8 EEP levels."""
    ensembles = _profile_ensembles()
    n_tf = 16
    bits = np.stack([synth.ModeITransmitter(e).generate(1, n_tf, seed=40 + k, want_iq=False)["bits"].numpy()[0]
                     for k, e in enumerate(ensembles)])
    rng = np.random.default_rng(3)
    bits[:, :, 9216:] ^= (rng.random(bits[:, :, 9216:].shape) < 0.02).astype(np.uint8)   # MSC only: stay locked
    for batch in (1, 3):
        got, st = _run_engine(gpu, bits, msc_batch=batch)
        for s in range(len(ensembles)):
            want, _, _ = port.run_backend(bits[s])
            assert got[s].shape == want.shape and want.shape[0] == 4 * (n_tf - 13), (s, got[s].shape, want.shape)
            assert np.array_equal(got[s], want), (batch, s)


def _crc16(data: bytes, crc: int = 0xFFFF) -> int:
    for b in data:
        crc ^= b << 8
        for _ in range(8):
            crc = ((crc << 1) ^ 0x1021) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return crc


def _parse_eti(frame):
    """ETI(NI) frame -> (nst, {SubChId: (stl words, payload bytes)}, header/EOF CRC ok)"""
    f = bytes(frame)
    nst = f[5] & 0x7F
    fl = ((f[6] & 7) << 8) | f[7]
    eoh = 8 + 4 * nst
    hcrc_ok = (~_crc16(f[4:eoh + 2]) & 0xFFFF) == (f[eoh + 2] << 8 | f[eoh + 3])
    pos = eoh + 4 + 96
    subs = {}
    for j in range(nst):
        w = f[8 + 4 * j: 12 + 4 * j]
        scid = w[0] >> 2
        stl = ((w[2] & 3) << 8) | w[3]
        subs[scid] = (w, f[pos: pos + 8 * stl])
        pos += 8 * stl
    mst = f[eoh + 4: pos]
    eof_ok = (~_crc16(mst) & 0xFFFF) == (f[pos] << 8 | f[pos + 1])
    assert fl == nst + 1 + (pos - (eoh + 4)) // 4
    return nst, subs, f[eoh + 4: eoh + 4 + 96], hcrc_ok and eof_ok, f[pos + 8:]


def test_subchannel_filter(gpu, port):
    """dabgpu_engine_set_subchannel_mask: the selected sub-channels come out exactly as in the
    unfiltered ETI (STC words, payload), the frame is a consistent ETI(NI) frame (NST, FL, both CRCs,
    padding), and the full mask is the reference's behaviour."""
    ens = synth.small_ensemble()            # SubChIds 3, 7, 12
    S, n_tf = 2, 18
    bits = synth.ModeITransmitter(ens).generate(S, n_tf, seed=31, want_iq=False)["bits"].numpy()
    full, _ = _run_engine(gpu, bits, msc_batch=2)
    want, _, _ = port.run_backend(bits[0])
    assert np.array_equal(full[0], want)

    def run(mask0, change_at=None, mask1=None):
        eng = gpu.Engine(S)
        eng.set_msc_batch(2)
        eng.set_subchannel_mask(mask0, stream=0)          # stream 1 keeps everything
        out = [[] for _ in range(S)]
        for t in range(n_tf + 1):
            if t == change_at:
                eng.set_subchannel_mask(mask1, stream=0)
            n = eng.process_demapped(bits[:, t]) if t < n_tf else eng.flush()
            eti, ids = eng.fetch_eti()
            for f, s in zip(eti, ids):
                out[s].append(f.copy())
        eng.close()
        return [np.array(o, dtype=np.uint8).reshape(-1, 6144) for o in out]

    keep = (1 << 3) | (1 << 12)
    got = run(keep)
    assert np.array_equal(got[1], full[1])                 # the other stream is untouched
    assert got[0].shape == full[0].shape
    for fr, ref_fr in zip(got[0], full[0]):
        nst, subs, fic, ok, pad = _parse_eti(fr)
        nst_r, subs_r, fic_r, ok_r, _ = _parse_eti(ref_fr)
        assert ok and ok_r and nst == 2 and nst_r == 3
        assert sorted(subs) == [3, 12] and fic == fic_r
        for scid in subs:
            assert subs[scid] == subs_r[scid]
        assert bytes(fr[:5]) == bytes(ref_fr[:5])          # ERR, FSYNC, FCT
        assert set(pad) == {0x55}
    # nothing selected: header + FIC only
    none = run(0)
    nst, subs, fic, ok, pad = _parse_eti(none[0][0])
    assert nst == 0 and ok and not subs
    # a change while frames are queued takes effect with the next transmission frame
    mixed = run(~0, change_at=16, mask1=keep)
    nsts = [_parse_eti(fr)[0] for fr in mixed[0]]
    assert nsts[:8] == [3] * 8 and nsts[-4:] == [2] * 4 and sorted(set(nsts)) == [2, 3]
    assert all(_parse_eti(fr)[3] for fr in mixed[0])


@pytest.mark.parametrize("batch", [1, 2, 4])
def test_multiplex_reorganisation_mid_stream(gpu, port, batch):
    """SURVEY 8f-2 / TODO.md:3.  The sub-channel table changes while the receiver is locked and (with
    batch > 1) while ETI frames of the old layout are still queued for a deferred MSC launch: one
    sub-channel moves and changes protection, a new one appears, two streams switch at different
    transmission frames and a third never does.  The reference has no reconfiguration handling: FIG
    0/1 entries overwrite ens_info as they arrive (fic.c:62-93, misc.c:14-27) and every ETI frame is
    built with the table of the moment (misc.c:153-278) -- the engine must do exactly that, frame
    for frame, so the oracle fed the same FIBs is the expected output."""
    A = synth.small_ensemble()
    B = synth.Ensemble([synth.SubChannel(id=3, start_cu=0, uep_index=35),
                        synth.SubChannel(id=7, start_cu=250, eep_level=1, size_cu=64),
                        synth.SubChannel(id=20, start_cu=400, uep_index=16)])
    n_a = (22, 19, 36)
    n_tf = 36
    rows = []
    for s, na in enumerate(n_a):
        ga = synth.ModeITransmitter(A).generate(1, na, seed=51 + s, want_iq=False)
        parts = [ga["bits"][0].numpy()]
        if na < n_tf:
            gb = synth.ModeITransmitter(B).generate(1, n_tf - na, seed=61 + s, want_iq=False, first_cif=4 * na)
            parts.append(gb["bits"][0].numpy())
        rows.append(np.concatenate(parts))
    bits = np.stack(rows)
    got, st = _run_engine(gpu, bits, msc_batch=batch)
    for s in range(3):
        want, _, _ = port.run_backend(bits[s])
        assert got[s].shape == want.shape and want.shape[0] == 4 * (n_tf - 13), (s, got[s].shape, want.shape)
        assert np.array_equal(got[s], want), s
    nst = [[int(f[5] & 0x7F) for f in g] for g in got]
    assert set(nst[0]) == {3, 4} and set(nst[1]) == {3, 4} and set(nst[2]) == {3}     # the switch happened
    assert nst[0].index(4) != nst[1].index(4)
    assert all(x.locked == 1 for x in st)


def test_wavefinder_packets_as_second_producer(gpu, port):
    """the Psion Wavefinder's USB packets (input_wf.c:23-115) as a producer of demapped
    transmission frames for the same back-end.  Stream 0 loses a FIC symbol in one frame (has_fic = 0
    -> NULL FIBs, fic.c:167-175), two MSC symbols in another (the frame buffer keeps what it held
    five frames earlier) and all FIC symbols in a third; stream 1 is clean; stream 2 skips a frame
    altogether.  ETI byte for byte as do_wf_decode (dab2eti.c:251-272) produces it in the oracle,
    which tests/test_oracle_vs_ref.py pins against the reference's unmodified input_wf.c."""
    ens = synth.small_ensemble()
    S, n_tf = 3, 23
    g = synth.ModeITransmitter(ens).generate(S, n_tf, seed=71, want_iq=False)
    bits = g["bits"].numpy()
    drops = {0: {16: (3,), 18: (40, 41), 19: (2, 3, 4)}, 1: {}, 2: {}}
    skip = {2: {17}}                                    # stream 2 delivers no frame in call 17
    per_stream = [[None if t in skip.get(s, ()) else synth.wavefinder_packets(bits[s, t], drop=drops[s].get(t, ()))
                   for t in range(n_tf)] for s in range(S)]
    eng = gpu.Engine(S)
    out = [[] for _ in range(S)]
    for t in range(1, n_tf):                            # the first frame is read and discarded (dab2eti.c:262)
        pk = np.zeros((S, 77 * 524), np.uint8)
        n = np.zeros(S, np.int32)
        for s in range(S):
            p = per_stream[s][t]
            if p is None:
                continue
            n[s] = p.shape[0]
            pk[s, : p.size] = p.reshape(-1)
        k = eng.process_wavefinder(pk, n)
        eti, ids = eng.fetch_eti()
        assert eti.shape[0] == k
        for f, s in zip(eti, ids):
            out[s].append(f.copy())
    eng.close()
    for s in range(S):
        stream = np.concatenate([p for p in per_stream[s] if p is not None])
        want = port.run_wf(stream)
        got = np.array(out[s], dtype=np.uint8).reshape(-1, 6144)
        assert got.shape == want.shape and want.shape[0] >= 28, (s, got.shape, want.shape)
        assert np.array_equal(got, want), s


@pytest.mark.parametrize("batch", [1, 4])
def test_follow_signalled_reconfiguration(gpu, port, batch):
    """SURVEY 8f-2, the part the reference does not have (TODO.md:3): DABGPU_ENGINE_FOLLOW_RECONFIG.
    The multiplex announces a reconfiguration the way EN 300 401 does (FIG 0/0 change flags +
    occurrence change, next configuration in FIG 0/1 with C/N = 1) six frames ahead; at the signalled
    CIF one sub-channel moves and changes protection, one disappears and a new one appears.  A
    following engine decodes EVERY logical frame -- also the 16 around the change, whose bits are
    spread over CIFs of both configurations -- with the table that was current for that frame: payload
    equal to what was transmitted, NST / STC of the right configuration.  The default engine (= the
    reference: merges the announcement at once, never drops a sub-channel, applies the newest table to
    15-CIF-old frames) garbles the frames around the change, which is what the mode is for."""
    A = synth.small_ensemble()
    B = synth.Ensemble([synth.SubChannel(id=3, start_cu=0, uep_index=35),
                        synth.SubChannel(id=7, start_cu=250, eep_level=1, size_cu=64),
                        synth.SubChannel(id=20, start_cu=400, uep_index=16)])
    S, n_tf, sw = 2, 34, 22
    g = synth.generate_reconfiguration(A, B, S, n_tf, sw, seed=3)
    bits = g["bits"].numpy()
    N = g["switch_cif"]

    def run(flags):
        eng = gpu.Engine(S, 200_000_000, flags)
        eng.set_msc_batch(batch)
        out = [[] for _ in range(S)]
        for t in range(n_tf + 1):
            n = eng.process_demapped(bits[:, t]) if t < n_tf else eng.flush()
            eti, ids = eng.fetch_eti()
            for f, s in zip(eti, ids):
                out[s].append(f.copy())
        eng.close()
        return [np.array(o, dtype=np.uint8).reshape(-1, 6144) for o in out]

    def check(frames, s):
        """-> (frames whose header and payload are right for their own CIF, total)"""
        good = 0
        for f in frames:
            cif = (int(f[4]) - 3) % 250                         # FCT leads the content by 3
            ens, pl = (A, g["payload_a"]) if cif < N else (B, g["payload_b"])
            nst = int(f[5] & 0x7F)
            ids = [int(f[8 + 4 * i] >> 2) for i in range(nst)]
            off = 12 + 4 * nst + 96
            want = synth.expected_eti_payload(ens, pl, s, cif)
            ok = ids == [sc.id for sc in ens.subchannels] and f[off: off + len(want)].tobytes() == want
            good += ok
        return good, frames.shape[0]

    got = run(gpu.ENGINE_FOLLOW_RECONFIG)
    for s in range(S):
        good, total = check(got[s], s)
        assert total == 4 * (n_tf - 13) and good == total, (s, good, total)
        assert not gpu.eti_check(got[s]).any()
        cifs = [(int(f[4]) - 3) % 250 for f in got[s]]
        assert min(cifs) < N - 16 and max(cifs) > N + 16        # frames on both sides and across the change
    ref_like = run(0)
    want = port.run_backend(bits[0])[0]
    assert np.array_equal(ref_like[0], want)                    # default mode: still the reference, byte for byte
    good, total = check(ref_like[0], 0)
    assert good < total - 16                                    # ... which loses the frames around the change
