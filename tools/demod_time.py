"""Per-kernel times of the batched receiver on whatever libdabgpu build DABGPU_LIB selects (kernel
experiments: a build whose results are deliberately wrong still runs its front-end kernels).

    [DABGPU_LIB=...] python tools/demod_time.py [streams]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from dabtools_b200 import lib

S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
lib.check(lib.load().dabgpu_set_device(0))
lib.use_torch_stream()
data, _ = bench.generate_dataset(S, bench.CAPTURE_TFS, dev, seed=1)
cap_calls = bench.CAPTURE_TFS // 2 * bench.CALLS_PER_2TF
chunk = lambda c: data[:, (c % cap_calls) * bench.CALL_BYTES: (c % cap_calls + 1) * bench.CALL_BYTES]
eng = lib.Engine(S)
eng.set_msc_batch(2)
for c in range(33):
    eng.feed_iq_device(chunk(c))
eng.enable_timing(True)
for c in range(33, 33 + 24):
    eng.feed_iq_device(chunk(c))
kt = eng.kernel_times()
eng.enable_timing(False)
print(json.dumps({"lib": os.environ.get("DABGPU_LIB", "default"), "locked": sum(eng.status(s).locked for s in range(S)),
                  "ms_per_launch": {k: round(v["ms"] / v["launches"], 4) for k, v in kt.items() if v["launches"]}}))
eng.close()
