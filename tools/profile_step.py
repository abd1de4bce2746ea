"""Steady-state steps of the batched receiver for profiling under ncu (profile range = the last steps).

    ncu --profile-from-start off ... python tools/profile_step.py [streams] [profiled_steps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from dabtools_b200 import lib

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
NPROF = int(sys.argv[2]) if len(sys.argv) > 2 else 1
lib.check(lib.load().dabgpu_set_device(0))
lib.use_torch_stream()
setup = bench.SETUP_TFS // 2
data, ens = bench.generate_dataset(S, 2 * (setup + 1 + NPROF), torch.device("cuda", 0), seed=1)
bench.CALLS_PER_STEP = bench.CALLS_PER_2TF
eng = lib.Engine(S)
eng.set_msc_batch(2)
step_bytes = 3 * bench.CALL_BYTES


def step(i):
    n = 0
    for c in range(3):
        off = i * step_bytes + c * bench.CALL_BYTES
        n += eng.feed_iq_device(data[:, off: off + bench.CALL_BYTES])
    return n


for i in range(setup + 1):
    n = step(i)
assert n > 0, n   # steady state: MSC batches of 2 (occasionally 3) transmission frames per stream
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(NPROF):
    step(setup + 1 + i)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", NPROF, "steps of", S, "streams;", lib.launch_count(), "launches in total")
