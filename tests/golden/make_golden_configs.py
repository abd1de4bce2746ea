"""Golden outputs of the UNMODIFIED reference (oracle/_ref/libdabref.so) at the sizes BASELINE.json's
configs state, as digests (the inputs are regenerated from seeds, the outputs are too large to store):

  cfg1  one ~80-TF capture of the 10-sub-channel reference ensemble, 30 dB SNR -> ETI (configs[0])
  cfg2  16384 FIC groups, 1/4 clean + 3/4 with 1 / 4 / 8 % bit flips -> FIBs + CRC flags (configs[1])

Run in the build container (needs /root/reference):   python tests/golden/make_golden_configs.py
"""
import hashlib
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dabtools_b200 import synth  # noqa: E402
from oracle import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CFG1 = dict(n_tf=84, seed=2026, snr_db=30.0, cut_bytes=2 * 41234)
CFG2 = dict(n_groups=16384, seed=2)


def cfg1_capture():
    ens = synth.reference_ensemble()
    g = synth.ModeITransmitter(ens).generate(1, CFG1["n_tf"], seed=CFG1["seed"], snr_db=CFG1["snr_db"],
                                             tail_samples=262144)
    iq = g["iq"][0].numpy()[CFG1["cut_bytes"]:]
    return iq[: iq.size // 262144 * 262144]


def frame_digests(eti: np.ndarray) -> np.ndarray:
    return np.stack([np.frombuffer(hashlib.sha256(f.tobytes()).digest(), dtype=np.uint8) for f in eti])


def ref_fic_groups(ref, bits):
    """fic.c:185-206 per group through the reference's own functions"""
    n = bits.shape[0]
    fibs = np.empty((n, 96), np.uint8)
    ok = np.empty((n, 3), np.uint8)
    for g in range(n):
        f = ref.descramble(ref.viterbi(ref.fic_depuncture(bits[g]), 768))
        fibs[g] = f
        ok[g] = [ref.check_fib_crc(f[32 * k: 32 * k + 32]) for k in range(3)]
    return fibs, ok


def main():
    ref = oracle.ref()
    assert ref is not None, "needs the compiled reference"
    out = {}
    iq = cfg1_capture()
    r = ref.run_iq(iq)
    tr = r["trace"]
    out["cfg1_iq_crc"] = np.array([zlib.crc32(iq.tobytes())], dtype=np.int64)
    out["cfg1_eti_sha256"] = frame_digests(r["eti"])
    out["cfg1_trace_int"] = np.stack([tr["ok"], tr["coarse_timeshift"], tr["fine_timeshift"],
                                      tr["coarse_freq_shift"], tr["locked"], tr["eti_frames"]], axis=1)
    print("cfg1:", iq.size // 393216, "TFs ->", r["eti"].shape[0], "ETI frames")

    bits, sent = synth.fic_groups(CFG2["n_groups"], CFG2["seed"])
    fibs, ok = ref_fic_groups(ref, bits)
    out["cfg2_in_crc"] = np.array([zlib.crc32(bits.tobytes())], dtype=np.int64)
    out["cfg2_fibs_sha256"] = np.frombuffer(hashlib.sha256(fibs.tobytes()).digest(), dtype=np.uint8)
    out["cfg2_fibs_crc32_per_1024"] = np.array([zlib.crc32(fibs[i:i + 1024].tobytes())
                                                for i in range(0, fibs.shape[0], 1024)], dtype=np.int64)
    out["cfg2_ok"] = np.packbits(ok)
    q = CFG2["n_groups"] // 4
    print("cfg2: CRC-ok rate per quarter:", [float(ok[i * q:(i + 1) * q].mean()) for i in range(4)],
          "clean quarter equals what was sent:", bool(np.array_equal(fibs[:q], sent[:q])))
    path = os.path.join(HERE, "baseline_configs_v1.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
