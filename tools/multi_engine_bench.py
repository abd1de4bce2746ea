"""Experiment: G engines of S/G streams each on ONE GPU, one host thread and one CUDA stream per
engine, so that the host state machines of one engine overlap the GPU work of the others.

    python tools/multi_engine_bench.py [streams] [engines] [steps] [msc_batch]
"""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from dabtools_b200 import lib

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
G = int(sys.argv[2]) if len(sys.argv) > 2 else 2
K = int(sys.argv[3]) if len(sys.argv) > 3 else 12
MB = int(sys.argv[4]) if len(sys.argv) > 4 else 2
W = 4
lib.check(lib.load().dabgpu_set_device(0))
dev = torch.device("cuda", 0)
setup = bench.SETUP_TFS // 2
data, ens = bench.generate_dataset(S, 2 * (setup + W + K + 1), dev, seed=1)
torch.cuda.synchronize()
step_bytes = 3 * bench.CALL_BYTES
per = S // G
start = threading.Barrier(G + 1)
done = threading.Barrier(G + 1)
frames = [0] * G


def worker(g):
    torch.cuda.set_device(0)
    lib.check(lib.load().dabgpu_set_device(0))
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        lib.use_torch_stream()
        eng = lib.Engine(per)
        eng.set_msc_batch(MB)
        mine = data[g * per:(g + 1) * per]

        def step(i):
            n = 0
            for c in range(3):
                off = i * step_bytes + c * bench.CALL_BYTES
                n += eng.feed_iq_device(mine[:, off: off + bench.CALL_BYTES])
            return n

        for i in range(setup + W):
            step(i)
        st.synchronize()
        start.wait()
        n = 0
        for i in range(K):
            n += step(setup + W + i)
        n += eng.flush()
        eng.join()
        st.synchronize()
        frames[g] = n
        done.wait()
        eng.close()


threads = [threading.Thread(target=worker, args=(g,)) for g in range(G)]
for t in threads:
    t.start()
start.wait()
t0 = time.perf_counter()
done.wait()
t1 = time.perf_counter()
for t in threads:
    t.join()
total = sum(frames)
print(f"engines={G} streams={S} msc_batch={MB} steps={K}: {total} frames in {(t1 - t0) * 1e3:.2f} ms "
      f"= {total / (t1 - t0):.0f} frames/s ({(t1 - t0) * 1e3 / K:.3f} ms/step)")
