"""Device-side ETI consumers: sub-channel extraction against the reference's own
eti2mpa.c (compiled unmodified into oracle/_ref/eti2mpa_ref) and against its restatement below;
the frame checker against frames the oracle built and against deliberately damaged ones."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from dabtools_b200 import synth

pytestmark = pytest.mark.gpu


def eti2mpa_restated(eti, subchid):
    """eti2mpa.c:32-67 per frame (the reference latches offset/length from the first frame; the
    multiplex does not change here, so both agree)"""
    out = []
    for f in eti:
        b = [int(x) for x in f[:8 + 4 * 64]]
        ficf, nst = b[5] >> 7, b[5] & 0x7F
        off, length = 0, -1
        for i in range(nst):
            scid = b[8 + 4 * i] >> 2
            stl = ((b[8 + 4 * i + 2] & 3) << 8) | b[8 + 4 * i + 3]
            if scid == subchid:
                length = stl * 8
                break
            off += stl * 8
        if length < 0:
            out.append(None)
        else:
            s0 = 12 + 4 * nst + ficf * 96 + off
            out.append(f[s0:s0 + length].copy())
    return out


@pytest.fixture(scope="module")
def eti_frames(gpu, port):
    ens = synth.reference_ensemble()
    g = synth.ModeITransmitter(ens).generate(1, 20, seed=41, want_iq=False)
    eti, _, _ = port.run_backend(g["bits"][0].numpy())
    assert eti.shape[0] >= 20
    return ens, g, eti


def test_extract_subchannel_matches_eti2mpa(gpu, eti_frames, tmp_path):
    ens, g, eti = eti_frames
    exe = os.path.join(ROOT, "oracle", "_ref", "eti2mpa_ref")
    eti_file = tmp_path / "frames.eti"       # eti2mpa read()s 6144 bytes at a time: a file, not a pipe
    eti.tofile(eti_file)
    for sc in ens.subchannels:
        data, lens = gpu.eti_extract_subchannel(eti, sc.id)
        want = eti2mpa_restated(eti, sc.id)
        assert lens.tolist() == [sc.nbytes] * eti.shape[0]
        for f in range(eti.shape[0]):
            assert np.array_equal(data[f, :lens[f]], want[f]), (sc.id, f)
            assert not data[f, lens[f]:].any()
        if os.path.exists(exe):      # the reference's own program on the same frames
            with open(eti_file, "rb") as fh:
                r = subprocess.run([exe, str(sc.id)], stdin=fh, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
            assert r.stdout == b"".join(data[f, :lens[f]].tobytes() for f in range(eti.shape[0])), sc.id
    # the payload really is what was transmitted (logical CIF of a frame: FCT leads it by 3)
    sc = ens.subchannels[2]
    data, lens = gpu.eti_extract_subchannel(eti, sc.id)
    for f in range(eti.shape[0]):
        cif = (int(eti[f, 4]) - 3) % 250
        assert data[f, :lens[f]].tobytes() == bytes(g["payload"][sc.id][0, cif].numpy())
    # a SubChId the multiplex does not carry
    _, lens = gpu.eti_extract_subchannel(eti, 40)
    assert (lens == -1).all()
    assert gpu.eti_extract_subchannel(np.zeros((0, 6144), np.uint8), 1)[1].size == 0


def test_check_eti(gpu, eti_frames):
    _, _, eti = eti_frames
    assert not gpu.eti_check(eti).any()
    bad = eti.copy()
    nst = int(eti[0, 5] & 0x7F)
    e1 = 12 + 4 * nst
    bad[1, 2] ^= 0x10                  # FSYNC
    bad[2, 9] ^= 0x01                  # an STC byte: header CRC (and, being a start address, nothing else)
    bad[3, e1 + 200] ^= 0x80           # a payload bit: end-of-frame CRC
    bad[4, 6143] = 0x54                # padding
    bad[5, 7] ^= 0x01                  # FL: FC and header CRC
    bad[6, 4] ^= 0x01                  # FCT parity no longer matches FSYNC; header CRC
    flags = gpu.eti_check(bad)
    assert flags[0] == 0 and not flags[7:].any()
    assert flags[1] == gpu.ETI_BAD_SYNC
    assert flags[2] == gpu.ETI_BAD_HCRC
    assert flags[3] == gpu.ETI_BAD_EOF_CRC
    assert flags[4] == gpu.ETI_BAD_PADDING
    assert flags[5] == gpu.ETI_BAD_FC | gpu.ETI_BAD_HCRC
    assert flags[6] == gpu.ETI_BAD_SYNC | gpu.ETI_BAD_HCRC


def test_engine_consumers_work_where_the_frames_lie(gpu, port):
    """dabgpu_engine_extract_subchannel / check_eti on the last call's frames in HBM: only the
    extracted bytes cross PCIe; equal to extracting from the fetched frames."""
    ens = synth.small_ensemble()
    S = 3
    g = synth.ModeITransmitter(ens).generate(S, 18, seed=42, want_iq=False)
    bits = g["bits"].numpy()
    eng = gpu.Engine(S)
    seen = 0
    for t in range(bits.shape[1]):
        n = eng.process_demapped(bits[:, t])
        if not n:
            continue
        eti, ids = eng.fetch_eti()
        assert not eng.check_eti().any()
        for sc in ens.subchannels:
            data, lens = eng.extract_subchannel(sc.id)
            want = eti2mpa_restated(eti, sc.id)
            assert lens.tolist() == [sc.nbytes] * n
            for f in range(n):
                assert np.array_equal(data[f, :lens[f]], want[f])
        seen += n
    eng.close()
    assert seen >= 4 * S * 3


def test_empty_and_degenerate_inputs(gpu, tmp_path):
    """empty batches, a source that ends before its first callback, frames without packets: no frames, no error"""
    import os
    assert gpu.eti_check(np.zeros((0, 6144), np.uint8)).size == 0
    assert gpu.viterbi_soft_batch(np.zeros((0, 4 * 774), np.uint8), 768).shape == (0, 96)
    garbage = np.full((3, 6144), 0xA5, np.uint8)                      # not ETI at all: flagged, nothing extracted
    assert (gpu.eti_check(garbage) & gpu.ETI_BAD_SYNC).all()
    data, lens = gpu.eti_extract_subchannel(garbage, 5)
    assert (lens == -1).all() and not data.any()
    eng = gpu.Engine(2)
    short = tmp_path / "short.iq"
    np.zeros(1000, np.uint8).tofile(short)
    fds = [os.open(short, os.O_RDONLY) for _ in range(2)]
    assert eng.pump(fds, None) == 0
    for fd in fds:
        os.close(fd)
    assert eng.process_wavefinder(np.zeros((2, 524), np.uint8), np.zeros(2, np.int32)) == 0
    assert eng.extract_subchannel(3)[1].size == 0 and eng.check_eti().size == 0
    eng.close()
