"""Device-resident timing of dabgpu_viterbi_batch over batch sizes (throughput scaling check)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from dabtools_b200 import lib

nbits = int(sys.argv[1]) if len(sys.argv) > 1 else 3072
L = lib.load()
lib.check(L.dabgpu_set_device(0))
lib.use_torch_stream()
for n in (4096, 16384, 32768, 65536, 131072):
    soft = torch.randint(127, 130, (n, 4 * (nbits + 6)), dtype=torch.uint8, device="cuda")
    out = torch.zeros((n, nbits // 8), dtype=torch.uint8, device="cuda")
    def run():
        lib.check(L.dabgpu_viterbi_batch(C.c_void_p(soft.data_ptr()), soft.shape[1], n, nbits,
                                         C.c_void_p(out.data_ptr()), out.shape[1], 1, 1))
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 5
    e0.record()
    for _ in range(K):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"nbits={nbits} n={n:7d} warps={n//32:5d}  {ms:8.3f} ms  {n*(nbits+6)*64/ms/1e9:8.1f} GACS/s  {n*nbits/ms/1e3:9.1f} Mbit/s")
    del soft, out
