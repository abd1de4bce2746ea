"""Quick device-resident timing of the batched FIC decode (BASELINE config 2 shape)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dabtools_b200 import lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
lib.check(lib.load().dabgpu_set_device(0))
lib.use_torch_stream()
rng = np.random.default_rng(0)
fic = torch.from_numpy(rng.integers(0, 2, (n, 2304), dtype=np.uint8)).cuda()
fibs = torch.zeros((n, 96), dtype=torch.uint8, device="cuda")
ok = torch.zeros((n, 3), dtype=torch.uint8, device="cuda")
for _ in range(3):
    lib.fic_decode_batch_device(fic, fibs, ok)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
e0.record()
for _ in range(K):
    lib.fic_decode_batch_device(fic, fibs, ok)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
steps = n * 774
print(f"n={n} {ms:.3f} ms/call  {n*768/ms/1e3:.1f} Mbit/s decoded  {steps*64/ms/1e9*1e3/1e3:.2f} GACS/s")
