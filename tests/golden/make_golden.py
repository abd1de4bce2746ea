"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libdabref.so).

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The fixtures let the oracle port (and through it the CUDA path) be checked against reference
outputs on machines where the reference itself is not available.
"""
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dabtools_b200 import synth  # noqa: E402
from dabtools_b200 import tables as T  # noqa: E402
from oracle import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    ref = oracle.ref()
    assert ref is not None, "needs the compiled reference"
    rng = np.random.default_rng(20261017)
    out = {}

    # 1. Viterbi: (soft input, decoded bytes) for clean / noisy / erased / adversarial inputs
    vit_meta = []
    k = 0
    for nbits in (192, 768, 1536, 3072):
        for p_flip, p_erase in [(0, 0), (0.03, 0.25), (0.08, 0.45), (0.5, 0)]:
            data = rng.integers(0, 256, nbits // 8, dtype=np.uint8)
            sym = ref.encode(data)
            s = sym ^ (rng.random(sym.size) < p_flip).astype(np.uint8)
            soft = (127 + 2 * s).astype(np.uint8)
            soft[rng.random(sym.size) < p_erase] = 128
            out[f"vit_in_{k}"] = soft
            out[f"vit_out_{k}"] = ref.viterbi(soft, nbits)
            out[f"vit_data_{k}"] = data
            vit_meta.append((nbits, p_flip, p_erase))
            k += 1
    for fill in (127, 128, 129):
        soft = np.full(4 * 774, fill, np.uint8)
        out[f"vit_in_{k}"] = soft
        out[f"vit_out_{k}"] = ref.viterbi(soft, 768)
        out[f"vit_data_{k}"] = np.zeros(96, np.uint8)
        vit_meta.append((768, -1, -1))
        k += 1
    out["vit_meta"] = np.array(vit_meta, dtype=np.float64)

    # 2. depuncture: CRC32 of the reference output for every profile on a fixed bit pattern
    pattern = rng.integers(0, 2, 64 * 416, dtype=np.uint8)
    out["dep_pattern"] = np.packbits(pattern)
    out["dep_fic"] = ref.fic_depuncture(pattern[:2304])
    uep = []
    for idx in range(64):
        o = ref.uep_depuncture(pattern[: 64 * T.UEP[idx][1]], idx)
        uep.append((o.size, zlib.crc32(o.tobytes())))
    out["dep_uep"] = np.array(uep, dtype=np.int64)
    eep = []
    for lvl in range(8):
        mul = (12, 8, 6, 4, 27, 21, 18, 15)[lvl]
        for n in (1, 2, 5, 12):
            size, bitrate = mul * n, n * (8 if lvl < 4 else 32)
            if bitrate > 384:
                continue
            o = ref.eep_depuncture(pattern[: 64 * size], lvl, size, bitrate)
            eep.append((lvl, size, bitrate, o.size, zlib.crc32(o.tobytes())))
    out["dep_eep"] = np.array(eep, dtype=np.int64)

    # 3. descramble / CRC / time de-interleave
    buf = rng.integers(0, 256, 1152, dtype=np.uint8)
    out["scr_in"] = buf
    out["scr_out"] = ref.descramble(buf)
    fibs = rng.integers(0, 256, (16, 32), dtype=np.uint8)
    fibs[0] = 0
    fibs[0, 0], fibs[0, 30], fibs[0, 31] = 0xFF, 0xA8, 0xA8
    out["crc_fibs"] = fibs
    out["crc_ok"] = np.array([ref.check_fib_crc(f) for f in fibs], dtype=np.uint8)
    cifs = [rng.integers(0, 2, 55296, dtype=np.uint8) for _ in range(16)]
    out["tdi_in"] = np.packbits(np.stack(cifs))
    out["tdi_out"] = np.packbits(ref.time_deinterleave(cifs))

    # 4. back-end: demapped TFs -> ETI (3 sub-channels, 15 TFs, light bit errors)
    ens = synth.small_ensemble()
    g = synth.ModeITransmitter(ens).generate(1, 15, seed=77, want_iq=False)
    bits = g["bits"][0].numpy().copy()
    bits ^= (rng.random(bits.shape) < 0.02).astype(np.uint8)
    eti, fb, crc = ref.run_backend(bits)
    out["be_bits"] = np.packbits(bits)
    out["be_eti"] = eti
    out["be_fibs"] = fb
    out["be_crc"] = crc

    # 5. front-end: one IQ frame -> sync estimates, a few spectra rows, demapped bits
    g = synth.ModeITransmitter(ens).generate(1, 1, seed=78, snr_db=22)
    frame = g["iq"][0].numpy()[:393216]
    d = ref.demod_frame(frame)
    out["fe_frame"] = frame
    out["fe_scalars"] = np.array([d["ok"], d["coarse_timeshift"], d["fine_timeshift"], d["coarse_freq_shift"],
                                  d["fine_freq_shift"]], dtype=np.float64)
    rows = np.array([0, 1, 2, 40, 75])
    out["fe_rows"] = rows
    out["fe_symbols"] = d["symbols"][rows]
    out["fe_symbols_d"] = d["symbols_d"][rows[1:]]
    out["fe_bits"] = np.packbits(d["bits"])

    # 6. whole path on a noiseless capture regenerated from seeds: ETI digest + integer trace
    g = synth.ModeITransmitter(ens).generate(1, 20, seed=79, snr_db=None, tail_samples=262144)
    iq = g["iq"][0].numpy()[2 * 31337:]
    r = ref.run_iq(iq)
    tr = r["trace"]
    out["e2e_trace_int"] = np.stack([tr["ok"], tr["coarse_timeshift"], tr["fine_timeshift"],
                                     tr["coarse_freq_shift"], tr["locked"], tr["eti_frames"]], axis=1)
    out["e2e_eti"] = r["eti"]
    out["e2e_iq_crc"] = np.array([zlib.crc32(iq.tobytes())], dtype=np.int64)

    path = os.path.join(HERE, "reference_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
