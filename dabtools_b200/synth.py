"""Synthetic DAB Mode I transmitter (test and benchmark input generator).

Not part of the receive hot path: it only manufactures inputs for it.  Conventions follow
SURVEY.md 8(d): random payload -> energy dispersal -> K=7 r=1/4 convolutional code -> UEP/EEP/FIC
puncturing -> 16-CIF time interleaving -> frequency interleaving -> pi/4-DQPSK against the phase
reference symbol -> 2048-point IFFT + 504-sample guard -> 2656-sample null symbol -> AWGN, carrier
offset, timing offset -> uint8 I/Q (value+127), i.e. exactly what an RTL-SDR hands to dab2eti.

Written in torch so that the same code makes a 20-frame test capture on the CPU and a
1024-stream benchmark batch on the GPU (torch.fft is used for *generation only*).
The transmit-side tables come from libdabgpu's C tables (include/dabgpu_tables.h) via
dabtools_b200.tables, which tests/test_tables.py pins against the compiled reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

from . import tables as T


@dataclass
class SubChannel:
    id: int
    start_cu: int
    uep_index: Optional[int] = None      # short form (UEP)
    eep_level: Optional[int] = None      # long form: 0..3 = 1-A..4-A, 4..7 = 1-B..4-B
    size_cu: Optional[int] = None

    def __post_init__(self):
        if self.uep_index is not None:
            self.shape = T.shape_uep(self.uep_index)
            self.size_cu = T.UEP[self.uep_index][1]
            self.bitrate = T.UEP[self.uep_index][0]
        else:
            self.shape = T.shape_eep(self.eep_level, self.size_cu)
            self.bitrate = self.shape["nbits"] // 24
        self.nbytes = self.shape["nbits"] // 8


@dataclass
class Ensemble:
    subchannels: List[SubChannel]
    eid: int = 0xCE15

    def __post_init__(self):
        self.subchannels = sorted(self.subchannels, key=lambda s: s.id)
        used = np.zeros(864, dtype=bool)
        for s in self.subchannels:
            assert s.start_cu + s.size_cu <= 864
            assert not used[s.start_cu:s.start_cu + s.size_cu].any(), "overlapping sub-channels"
            used[s.start_cu:s.start_cu + s.size_cu] = True

    @property
    def bytes_per_cif(self):
        return sum(s.nbytes for s in self.subchannels)

    @property
    def bits_per_frame(self):
        """decoded bits per ETI frame: all sub-channels plus the FIC"""
        return 8 * self.bytes_per_cif + 768

    @property
    def steps_per_frame(self):
        return sum(s.shape["nbits"] + 6 for s in self.subchannels) + 774


def reference_ensemble() -> Ensemble:
    """The 10-sub-channel 'CrystalPalace-like' multiplex of SURVEY.md 8(d): 750 of 864 CU,
    1008 kbit/s; per ETI frame 24 960 decoded bits, 11 codewords, 25 026 trellis steps."""
    spec = [(1, 35), (2, 35), (3, 45), (4, 35), (5, 16), (6, 35), (7, 21), (8, 16)]
    subs, cu = [], 0
    for sid, idx in spec:
        subs.append(SubChannel(id=sid, start_cu=cu, uep_index=idx))
        cu += T.UEP[idx][1]
    subs.append(SubChannel(id=9, start_cu=cu, eep_level=2, size_cu=48))   # EEP 3-A 64 kbit/s
    cu += 48
    subs.append(SubChannel(id=10, start_cu=cu, eep_level=4, size_cu=27))  # EEP 1-B 32 kbit/s
    return Ensemble(subs)


def small_ensemble() -> Ensemble:
    """1 UEP + 2 EEP sub-channels (the surveyor's back-end probe layout)."""
    return Ensemble([
        SubChannel(id=3, start_cu=0, uep_index=35),
        SubChannel(id=7, start_cu=200, eep_level=2, size_cu=48),
        SubChannel(id=12, start_cu=300, eep_level=5, size_cu=42),
    ])


# --------------------------------------------------------------------------------------
def _crc16(data: bytes, crc: int = 0xFFFF) -> int:
    for b in data:
        crc ^= b << 8
        for _ in range(8):
            crc = ((crc << 1) ^ 0x1021) & 0xFFFF if crc & 0x8000 else (crc << 1) & 0xFFFF
    return crc


def _fib(figs: bytes) -> bytes:
    assert len(figs) <= 30
    body = bytearray(figs)
    if len(body) < 30:
        body.append(0xFF)
    body.extend(b"\x00" * (30 - len(body)))
    crc = (~_crc16(bytes(body))) & 0xFFFF
    return bytes(body) + bytes([crc >> 8, crc & 0xFF])


def _fig01_entries(ens: Ensemble):
    entries = []
    for s in ens.subchannels:
        b0, b1 = (s.id << 2) | (s.start_cu >> 8), s.start_cu & 0xFF
        if s.uep_index is not None:
            entries.append(bytes([b0, b1, s.uep_index & 0x3F]))
        else:
            opt, lvl = s.eep_level >> 2, s.eep_level & 3
            entries.append(bytes([b0, b1, 0x80 | (opt << 4) | (lvl << 2) | (s.size_cu >> 8), s.size_cu & 0xFF]))
    return entries


def build_fibs(ens: Ensemble, cif_count: int, next_ens: Optional[Ensemble] = None,
               occurrence: Optional[int] = None) -> bytes:
    """3 FIBs (96 bytes) for one CIF: FIG 0/0 + FIG 0/1 sub-channel organisation.  With `next_ens` /
    `occurrence` the CIF also announces a multiplex reconfiguration the way EN 300 401 clause 6.4
    does: change flags and occurrence change (lower CIF count at which it takes effect) in FIG 0/0,
    and the next configuration's sub-channel organisation as FIG 0/1 with the C/N flag set."""
    hi, lo = (cif_count // 250) % 20, cif_count % 250
    if occurrence is None:
        fig00 = bytes([0x05, 0x00, ens.eid >> 8, ens.eid & 0xFF, hi, lo])
    else:
        fig00 = bytes([0x06, 0x00, ens.eid >> 8, ens.eid & 0xFF, 0x40 | hi, lo, occurrence % 250])
    groups = [(0x01, _fig01_entries(ens))]
    if next_ens is not None:
        groups.append((0x81, _fig01_entries(next_ens)))          # C/N = 1: next configuration
    fibs, cur = [], bytearray(fig00)
    for hdr, entries in groups:
        group = bytearray()
        for e in entries:
            if len(cur) + 2 + len(group) + len(e) > 30:
                if group:
                    cur.extend(bytes([len(group) + 1, hdr]) + group)
                    group = bytearray()
                fibs.append(_fib(bytes(cur)))
                cur = bytearray()
            group.extend(e)
        if group:
            if len(cur) + 2 + len(group) > 30:
                fibs.append(_fib(bytes(cur)))
                cur = bytearray()
            cur.extend(bytes([len(group) + 1, hdr]) + group)
    fibs.append(_fib(bytes(cur)))
    while len(fibs) < 3:
        fibs.append(_fib(b""))
    assert len(fibs) == 3, "ensemble does not fit 3 FIBs"
    return b"".join(fibs)


def generate_reconfiguration(ens_a: Ensemble, ens_b: Ensemble, n_streams: int, n_tf: int, switch_tf: int,
                             seed: int = 0, announce_tfs: int = 6):
    """Demapped transmission frames (bits, no IQ) of a multiplex that changes from `ens_a` to `ens_b` at the
    first CIF of transmission frame `switch_tf`, signalled like a real one: for `announce_tfs` frames
    before the change FIG 0/0 carries change flags + occurrence change and FIG 0/1 also lists the next
    configuration (C/N = 1).  Logical frames before the change are coded with ens_a's layout, the
    others with ens_b's; the time interleaver runs across the change (each logical frame's bits keep
    their own positions).  Returns dict(bits [S][n_tf][230400], payload_a, payload_b, fibs, switch_cif)."""
    ta, tb = ModeITransmitter(ens_a), ModeITransmitter(ens_b)
    S, n_cif, N = n_streams, 4 * n_tf, 4 * switch_tf
    rng = np.random.default_rng(0xDAB0000 + seed)
    logical = torch.zeros((S, n_cif, 55296), dtype=torch.uint8)
    payloads = []
    for tx, ens, lo, hi in ((ta, ens_a, 0, N), (tb, ens_b, N, n_cif)):
        pl = {}
        for s in ens.subchannels:
            p = torch.from_numpy(rng.integers(0, 256, (S, n_cif, s.nbytes), dtype=np.uint8))
            pl[s.id] = p
            coded = tx._code_block(p[:, lo:hi], s.id)
            logical[:, lo:hi, s.start_cu * 64: s.start_cu * 64 + coded.shape[-1]] = coded
        payloads.append(pl)
    tx_bits = torch.zeros_like(logical)
    for m in range(16):
        dl = int(T.TDI_DELAY[m])
        tx_bits[:, dl:, m::16] = logical[:, : n_cif - dl, m::16]
    fib_rows = []
    for c in range(n_cif):
        if c < N:
            ann = c >= N - 4 * announce_tfs
            fib_rows.append(build_fibs(ens_a, c, ens_b if ann else None, N if ann else None))
        else:
            fib_rows.append(build_fibs(ens_b, c))
    fibs = torch.from_numpy(np.frombuffer(b"".join(fib_rows), dtype=np.uint8).reshape(n_cif, 96).copy())
    fic = ta._code_block(fibs, "fic").reshape(n_tf, 9216).unsqueeze(0).expand(S, n_tf, 9216)
    bits = torch.cat([fic, tx_bits.reshape(S, n_tf, 4 * 55296)], dim=-1).contiguous()
    return dict(bits=bits, payload_a=payloads[0], payload_b=payloads[1], fibs=fibs, switch_cif=N)


# --------------------------------------------------------------------------------------
class ModeITransmitter:
    def __init__(self, ens: Ensemble, device="cpu"):
        self.ens = ens
        self.dev = torch.device(device)
        d = self.dev
        self.prbs_bits = torch.from_numpy(np.unpackbits(T.prbs(1152))).to(d)
        self.rev = torch.from_numpy(T.freq_deint().astype(np.int64)).to(d)
        self.prs_q = torch.from_numpy(T.prs().astype(np.int64)).to(d)
        # carrier index c -> FFT bin
        c = np.arange(1536)
        k = np.where(c < 768, c - 768, c - 767)
        self.bins = torch.from_numpy((k % 2048).astype(np.int64)).to(d)
        self.keep = {}
        for s in list(ens.subchannels) + ["fic"]:
            sh = T.shape_fic() if s == "fic" else s.shape
            key = "fic" if s == "fic" else s.id
            self.keep[key] = torch.from_numpy(T.kept_positions(sh)).to(d)
        self.delay = torch.tensor(T.TDI_DELAY, device=d)

    # -- channel coding -------------------------------------------------------------------
    def _encode(self, data_bits: torch.Tensor) -> torch.Tensor:
        """[..., nbits] 0/1 -> mother code [..., 4*(nbits+6)] (viterbi.c:322-347 semantics)"""
        nb = data_bits.shape[-1]
        pad = torch.zeros(data_bits.shape[:-1] + (6,), dtype=data_bits.dtype, device=data_bits.device)
        x = torch.cat([pad, data_bits, pad], dim=-1)  # x[6+t] = bit t
        out = torch.zeros(data_bits.shape[:-1] + (nb + 6, 4), dtype=torch.uint8, device=data_bits.device)
        for j, poly in enumerate(T.POLYS):
            acc = torch.zeros(data_bits.shape[:-1] + (nb + 6,), dtype=torch.uint8, device=data_bits.device)
            for kbit in range(7):
                if (poly >> kbit) & 1:
                    acc ^= x[..., 6 - kbit: 6 - kbit + nb + 6]
            out[..., j] = acc
        return out.reshape(data_bits.shape[:-1] + (4 * (nb + 6),))

    def _code_block(self, payload: torch.Tensor, key) -> torch.Tensor:
        """payload bytes [..., nbytes] -> punctured channel bits [..., in_bits]"""
        nbytes = payload.shape[-1]
        shifts = torch.arange(7, -1, -1, device=payload.device, dtype=torch.uint8)
        bits = ((payload.unsqueeze(-1) >> shifts) & 1).reshape(payload.shape[:-1] + (8 * nbytes,))
        bits = bits ^ self.prbs_bits[: 8 * nbytes]
        mother = self._encode(bits)
        return mother[..., self.keep[key]]

    # -- one batch ------------------------------------------------------------------------
    def generate(self, n_streams: int, n_tf: int, seed: int = 0, snr_db: Optional[float] = 40.0,
                 cfo_hz=0.0, lead_samples=0, amplitude: float = 32.0, want_iq: bool = True,
                 first_cif: int = 0, tail_samples: int = 0):
        """Returns dict with
             payload : {subch id: uint8 [S][n_cif][nbytes]}   (logical CIF order)
             fibs    : uint8 [n_cif][96]                        (same for all streams)
             bits    : uint8 [S][n_tf][230400]  ideal demapped hard bits (fic 9216 + msc 221184)
             iq      : uint8 [S][2*(lead + n_tf*196608 + tail)]   (if want_iq)
        """
        d, ens = self.dev, self.ens
        S, n_cif = n_streams, 4 * n_tf
        rng = np.random.default_rng(0xDAB0000 + seed)
        payload = {}
        logical = torch.zeros((S, n_cif, 55296), dtype=torch.uint8, device=d)
        for s in ens.subchannels:
            p = torch.from_numpy(rng.integers(0, 256, (S, n_cif, s.nbytes), dtype=np.uint8)).to(d)
            payload[s.id] = p
            coded = self._code_block(p, s.id)
            logical[:, :, s.start_cu * 64: s.start_cu * 64 + coded.shape[-1]] = coded
        # time interleaving: tx[c][i] = logical[c - delay[i & 15]][i]
        tx = torch.zeros_like(logical)
        for m in range(16):
            dl = int(T.TDI_DELAY[m])
            if dl < n_cif:
                tx[:, dl:, m::16] = logical[:, : n_cif - dl, m::16]
        # FIC (identical for every stream)
        fibs_np = np.frombuffer(b"".join(build_fibs(ens, first_cif + c) for c in range(n_cif)), dtype=np.uint8)
        fibs = torch.from_numpy(fibs_np.reshape(n_cif, 96).copy()).to(d)
        fic = self._code_block(fibs, "fic")                           # [n_cif][2304]
        fic = fic.reshape(n_tf, 9216).unsqueeze(0).expand(S, n_tf, 9216)
        bits = torch.cat([fic, tx.reshape(S, n_tf, 4 * 55296)], dim=-1).contiguous()  # [S][n_tf][230400]
        out = dict(payload=payload, fibs=fibs, bits=bits)
        if not want_iq:
            return out

        # frequency interleaving + pi/4-DQPSK; phases in units of pi/4
        sym = bits.reshape(S, n_tf, 75, 3072)
        b0 = sym[..., self.rev].to(torch.int64)            # [S][n_tf][75][1536] by carrier index
        b1 = sym[..., 1536 + self.rev].to(torch.int64)
        # (b0,b1): 00->1, 10->3, 11->5, 01->7 eighth-turns
        step = 1 + 2 * b0 + 6 * b1 - 4 * b0 * b1
        ph = torch.cumsum(step, dim=2) + 2 * self.prs_q     # [S][n_tf][75][1536]
        ph = torch.cat([(2 * self.prs_q).expand(S, n_tf, 1, 1536), ph], dim=2) & 7   # 76 symbols
        ang = ph.to(torch.float32) * (math.pi / 4)
        spec = torch.zeros((S, n_tf, 76, 2048), dtype=torch.complex64, device=d)
        spec[..., self.bins] = torch.polar(torch.ones_like(ang), ang)
        x = torch.fft.ifft(spec, dim=-1)                     # 1/N normalised
        gain = amplitude * 2048.0 / math.sqrt(768.0)         # sigma of each real component = amplitude
        x = x * gain
        symt = torch.cat([x[..., 2048 - 504:], x], dim=-1)   # guard interval
        frame = torch.cat([torch.zeros((S, n_tf, 2656), dtype=torch.complex64, device=d),
                           symt.reshape(S, n_tf, 76 * 2552)], dim=-1)
        sig = frame.reshape(S, n_tf * 196608)
        if lead_samples or tail_samples:
            sig = torch.cat([torch.zeros((S, lead_samples), dtype=torch.complex64, device=d), sig,
                             torch.zeros((S, tail_samples), dtype=torch.complex64, device=d)], dim=-1)
        n_total = sig.shape[-1]
        cfo = torch.as_tensor(cfo_hz, dtype=torch.float64, device=d).reshape(-1)
        if cfo.numel() == 1:
            cfo = cfo.expand(S)
        if bool((cfo != 0).any()):
            n = torch.arange(n_total, device=d, dtype=torch.float64)
            phase = (2 * math.pi / 2048000.0) * cfo[:, None] * n[None, :]
            sig = sig * torch.polar(torch.ones_like(phase), phase).to(torch.complex64)
        if snr_db is not None:
            g = torch.Generator(device=d)
            g.manual_seed(0x5EED0000 + seed)
            sigma = math.sqrt(amplitude * amplitude / (10.0 ** (snr_db / 10.0)))
            noise = torch.randn((S, n_total, 2), generator=g, device=d, dtype=torch.float32) * sigma
            sig = sig + torch.view_as_complex(noise)
        iq = torch.view_as_real(sig)
        iq = torch.clamp(torch.round(iq) + 127.0, 0, 255).to(torch.uint8).reshape(S, 2 * n_total)
        out["iq"] = iq
        return out


def _crc16_rows(rows: np.ndarray) -> np.ndarray:
    """CRC-16-CCITT (init 0xffff, no inversion) of every row of a uint8 [n][k] array"""
    tab = np.zeros(256, dtype=np.uint16)
    for b in range(256):
        c = b << 8
        for _ in range(8):
            c = ((c << 1) ^ 0x1021) & 0xFFFF if c & 0x8000 else (c << 1) & 0xFFFF
        tab[b] = c
    crc = np.full(rows.shape[0], 0xFFFF, dtype=np.uint16)
    for j in range(rows.shape[1]):
        crc = ((crc << 8) & 0xFFFF) ^ tab[(crc >> 8) ^ rows[:, j]]
    return crc


def fic_groups(n_groups: int, seed: int = 0, flips=(0.0, 0.01, 0.04, 0.08)):
    """BASELINE config 2 input: `n_groups` FIC groups (one CIF's 3 FIBs = one 768-bit codeword each).
    Every FIB carries 30 random bytes and a valid CRC, is scrambled, encoded and punctured like the
    FIC (fic.c:160-208 in reverse); quarter q of the groups then gets i.i.d. bit flips with
    probability flips[q].  Returns (bits uint8 [n][2304] hard bits, fibs uint8 [n][96] as sent)."""
    rng = np.random.default_rng(0xF1C0000 + seed)
    fibs = rng.integers(0, 256, (n_groups * 3, 32), dtype=np.uint8)
    crc = ~_crc16_rows(fibs[:, :30]) & 0xFFFF
    fibs[:, 30] = crc >> 8
    fibs[:, 31] = crc & 0xFF
    fibs = fibs.reshape(n_groups, 96)
    tx = ModeITransmitter(Ensemble([]), "cpu")
    bits = np.empty((n_groups, 2304), dtype=np.uint8)
    for i in range(0, n_groups, 2048):
        bits[i:i + 2048] = tx._code_block(torch.from_numpy(fibs[i:i + 2048]), "fic").numpy()
    n = n_groups
    for q, p in enumerate(flips):
        sl = slice(q * n // len(flips), (q + 1) * n // len(flips))
        if p > 0:
            bits[sl] ^= (rng.random(bits[sl].shape) < p).astype(np.uint8)
    return bits, fibs


def wavefinder_packets(bits_tf: np.ndarray, drop=()) -> np.ndarray:
    """One transmission frame of ideal demapped bits (uint8 [230400]: 3 FIC + 72 MSC symbols of 3072)
    as the Psion Wavefinder delivers it over USB (input_wf.c:23-115 in reverse): 524-byte packets,
    byte 2 = symbol number (1 = PRS, 2..76 data, 0 = the NULL symbol that ends the frame), bytes 12..395
    = 192 little-endian words of 8 DQPSK decisions each in carrier order.  `drop`: symbol numbers lost."""
    rev = T.freq_deint().astype(np.int64)
    sym = np.asarray(bits_tf, dtype=np.uint8).reshape(75, 3072)
    b0 = sym[:, rev].reshape(75, 192, 8).astype(np.uint16)          # carrier order: dst[rev[q]] = b0
    b1 = sym[:, 1536 + rev].reshape(75, 192, 8).astype(np.uint16)
    sh0 = (15 - 2 * np.arange(8)).astype(np.uint16)
    words = np.ascontiguousarray(((b0 << sh0) | (b1 << (sh0 - 1))).sum(axis=2).astype("<u2"))
    order = [1] + [n for n in range(2, 77) if n not in drop] + [0]
    pk = np.zeros((len(order), 524), dtype=np.uint8)
    for i, n in enumerate(order):
        pk[i, 2] = n
        if n >= 2:
            pk[i, 12:396] = words[n - 2].view(np.uint8)
    return pk


def expected_eti_payload(ens: Ensemble, payload: dict, stream: int, logical_cif: int) -> bytes:
    """MST sub-channel bytes of the ETI frame that carries logical CIF `logical_cif`."""
    return b"".join(bytes(payload[s.id][stream, logical_cif].cpu().numpy()) for s in ens.subchannels)
