"""Per-callback wall-clock trace of the pinned-host path (submit_iq / feed_submitted / fetch_eti)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from dabtools_b200 import lib

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
NSTEP = int(sys.argv[2]) if len(sys.argv) > 2 else 8
BATCH = int(sys.argv[3]) if len(sys.argv) > 3 else 4
AHEAD = int(sys.argv[4]) if len(sys.argv) > 4 else 2
lib.check(lib.load().dabgpu_set_device(0))
lib.use_torch_stream()
dev = torch.device("cuda", 0)
setup = bench.SETUP_TFS // 2
data, ens = bench.generate_dataset(S, 2 * (setup + NSTEP + 1), dev, seed=1)
eng = lib.Engine(S)
eng.set_msc_batch(BATCH)
sb = 3 * bench.CALL_BYTES
for i in range(setup):
    for c in range(3):
        eng.feed_iq_device(data[:, i * sb + c * bench.CALL_BYTES: i * sb + (c + 1) * bench.CALL_BYTES])
host_in = torch.empty((NSTEP, 3, S, bench.CALL_BYTES), dtype=torch.uint8, pin_memory=True)
for i in range(NSTEP):
    for c in range(3):
        off = (setup + i) * sb + c * bench.CALL_BYTES
        host_in[i, c].copy_(data[:, off: off + bench.CALL_BYTES])
host_out = torch.empty((S * 4 * (BATCH + 1), 6144), dtype=torch.uint8, pin_memory=True).numpy()
calls = [host_in[i, c].numpy() for i in range(NSTEP) for c in range(3)]
torch.cuda.synchronize()
t0 = time.perf_counter()
log = []
for k in range(min(AHEAD, len(calls))):
    eng.submit_iq(calls[k])
for k in range(len(calls)):
    a = time.perf_counter()
    if k + AHEAD < len(calls):
        eng.submit_iq(calls[k + AHEAD])
    b = time.perf_counter()
    n = eng.feed_submitted()
    c = time.perf_counter()
    if n:
        eng.fetch_eti(host_out)
    d = time.perf_counter()
    log.append((k, n, (a - t0) * 1e3, (b - a) * 1e3, (c - b) * 1e3, (d - c) * 1e3))
n = eng.flush()
if n:
    eng.fetch_eti(host_out)
torch.cuda.synchronize()
t1 = time.perf_counter()
for r in log:
    print("call %2d frames %6d  t=%8.2f ms  submit %6.2f  feed %6.2f  fetch %6.2f" % r)
tot = sum(r[1] for r in log) + n
print(f"total {tot} frames in {(t1 - t0) * 1e3:.1f} ms = {tot / (t1 - t0):.0f} frames/s; "
      f"ideal upload time at 52 GB/s: {len(calls) * S * bench.CALL_BYTES / 52e9 * 1e3:.1f} ms")
