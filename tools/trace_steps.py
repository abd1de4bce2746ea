"""GPU timeline of a few steady-state steps (DABGPU_TRACE): which kernels overlap, where the gaps are.

    DABGPU_TRACE=gpurun_out/trace.txt python tools/trace_steps.py [streams] [steps] [msc_batch]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from dabtools_b200 import lib

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 6
B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
lib.check(lib.load().dabgpu_set_device(0))
lib.use_torch_stream()
setup = bench.SETUP_TFS // 2
data, ens = bench.generate_dataset(S, 2 * (setup + N + 6), torch.device("cuda", 0), seed=1)
eng = lib.Engine(S)
eng.set_msc_batch(B)
eng.attach_capture(data)
W = 5
for i in range(3 * (setup + W)):
    eng.feed_capture(bench.CALL_BYTES)
eng.flush()          # same sequence as bench.py's device-resident pass
eng.join()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
for i in range(3 * N):
    eng.feed_capture(bench.CALL_BYTES)
eng.flush()
eng.join()
torch.cuda.synchronize()
print(f"{N} steps in {(time.perf_counter() - t0) * 1e3:.2f} ms wall")
eng.join()
torch.cuda.synchronize()
eng.close()
path = os.environ.get("DABGPU_TRACE")
if path and os.path.exists(path):
    rows = [l.split() for l in open(path)]
    # last N steps only
    t_end = max(float(r[2]) for r in rows)
    t0 = t_end - (N - 1) * 2.7
    busy = []
    for r in rows:
        a, b = float(r[1]), float(r[2])
        if b >= t0:
            busy.append((max(a, t0), b, r[0]))
    busy.sort()
    # union of busy intervals
    tot = 0.0
    cur_a, cur_b = busy[0][0], busy[0][1]
    gaps = []
    for a, b, _ in busy[1:]:
        if a > cur_b:
            tot += cur_b - cur_a
            gaps.append((cur_b, a))
            cur_a, cur_b = a, b
        else:
            cur_b = max(cur_b, b)
    tot += cur_b - cur_a
    span = cur_b - busy[0][0]
    print(f"window {span:.3f} ms, union of traced kernels busy {tot:.3f} ms ({100 * tot / span:.1f} %), {len(gaps)} gaps")
    for a, b in sorted(gaps, key=lambda g: g[0] - g[1])[:12]:
        print(f"  gap {b - a:.3f} ms at {a:.3f}")
