"""Where does a stream of the config-4 sweep leave the reference?  Re-generates the given sweep streams
(same seeds as tools/snr_cfo_sweep.py), runs the GPU engine and the oracle callback by callback and
prints the first callback whose synchroniser outputs / tuner frequency differ.

    python tools/diag_divergence.py 13 85 157 229 23 95 167 239
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from snr_cfo_sweep import CFOS, SNRS


def main():
    from dabtools_b200 import lib, synth
    from oracle import oracle
    port = oracle.ref() or oracle.port()
    cells = [(snr, cfo) for snr in SNRS for cfo in CFOS]
    ens = synth.small_ensemble()
    tx = synth.ModeITransmitter(ens, "cuda")
    lib.check(lib.load().dabgpu_set_device(0))
    tfs = 50
    for s in [int(a) for a in sys.argv[1:]]:
        snr, cfo = cells[s % len(cells)]
        g = tx.generate(1, tfs, seed=1000 + s, snr_db=float(snr), cfo_hz=float(cfo), tail_samples=262144)
        cut = 2 * (7919 * (s + 1) % 190000)
        iq = g["iq"][0, cut:][: (tfs - 1) * 393216].cpu().numpy()
        n = iq.size // 262144 * 262144
        iq = iq[:n]
        os.dup2(os.open(os.devnull, os.O_WRONLY), 2)
        r = port.run_iq(iq, seed=1)
        tr = r["trace"]
        eng = lib.Engine(1, 200_000_000, lib.ENGINE_VIRTUAL_TUNER)
        eng.set_seed(0, 1)
        first = None
        frames = 0
        for k, pos in enumerate(range(0, n, 262144)):
            frames += eng.feed_iq(iq[None, pos:pos + 262144])
            st = eng.status(0)
            mine = (st.last_ok, st.coarse_timeshift, st.fine_timeshift, st.coarse_freq_shift, st.frequency)
            want = (int(tr["ok"][k]), int(tr["coarse_timeshift"][k]), int(tr["fine_timeshift"][k]),
                    int(tr["coarse_freq_shift"][k]), int(tr["frequency"][k]))
            if mine != want and first is None:
                first = (k, mine, want, st.fine_freq_shift, float(tr["fine_freq_shift"][k]))
        eng.close()
        print(f"stream {s} (SNR {snr} dB, CFO {cfo} Hz): frames GPU {frames} / ref {r['eti'].shape[0]}; "
              f"first differing callback: {first}")


if __name__ == "__main__":
    main()
