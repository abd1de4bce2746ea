"""Concurrent pinned host->device copy probe, one process per GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_probe.py

Every rank copies the bench's own end-to-end traffic pattern -- 268 MB chunks (1024 streams x 262144
bytes, pinned) host->device with cudaMemcpyAsync, three in flight, plus the ETI stream back -- with no
kernels and no engine: what the box can feed N GPUs at once.  bench.py's `e2e` cannot beat
98304 bytes per ETI frame at this rate.  Prints one JSON line (rank 0): GB/s per rank, aggregate, and
the frames/s ceiling they imply.  Variants: pinned memory allocated before / after binding the process
to the GPU's NUMA node, and write-combined pinned memory (cudaHostAllocWriteCombined)."""
import ctypes
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench

CHUNK = 1024 * 262144
N_CHUNKS, REPS = 4, 12


def copy_rate(host_chunks, dev, back_host=None, back_dev=None):
    """GB/s of REPS x len(host_chunks) chunk uploads, optionally with a concurrent D2H stream"""
    s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
    for h in host_chunks[:2]:
        dev.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(s_up):
        e0.record()
        for r in range(REPS):
            for k, h in enumerate(host_chunks):
                dev[k % 3].copy_(h, non_blocking=True)
                if back_host is not None and k % 3 == 2:
                    with torch.cuda.stream(s_down):
                        back_host.copy_(back_dev, non_blocking=True)
        e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    return REPS * len(host_chunks) * CHUNK / ms / 1e6


def wc_pinned(nbytes):
    """cudaHostAlloc(..., cudaHostAllocWriteCombined) as a torch uint8 tensor (never freed: probe only)"""
    rt = ctypes.CDLL("libcudart.so.12")
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(4))
    if rc != 0:
        return None
    buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.uint8)


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29611")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.empty((3, CHUNK), dtype=torch.uint8, device="cuda")
    back_dev = torch.empty(1024 * 16 * 6144, dtype=torch.uint8, device="cuda")   # 16 ETI frames per stream
    res = {}
    # (a) pinned memory allocated wherever the process happened to start
    host = [torch.empty(CHUNK, dtype=torch.uint8, pin_memory=True) for _ in range(N_CHUNKS)]
    for h in host:
        h.fill_(1)
    back_host = torch.empty(back_dev.numel(), dtype=torch.uint8, pin_memory=True)
    res["h2d_unbound"] = copy_rate(host, dev)
    # (b) after binding to the GPU's NUMA node (what bench.py does), freshly allocated
    numa = bench.bind_to_gpu_numa_node(local)
    host_b = [torch.empty(CHUNK, dtype=torch.uint8, pin_memory=True) for _ in range(N_CHUNKS)]
    for h in host_b:
        h.fill_(2)
    res["h2d_numa_bound"] = copy_rate(host_b, dev)
    res["h2d_numa_bound_with_d2h"] = copy_rate(host_b, dev, back_host, back_dev)
    # (c) write-combined pinned memory
    wc = [wc_pinned(CHUNK) for _ in range(N_CHUNKS)]
    if all(w is not None for w in wc):
        for w in wc:
            w.fill_(3)
        res["h2d_write_combined"] = copy_rate(wc, dev)
    keys = sorted(res)
    t = torch.tensor([res[k] for k in keys], dtype=torch.float64, device="cuda")
    allr = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(allr, t)
    if rank == 0:
        out = {"ranks": world, "chunk_mb": CHUNK / 1e6, "numa": numa, "cores": len(os.sched_getaffinity(0))}
        for i, k in enumerate(keys):
            per = [float(a[i]) for a in allr]
            out[k] = {"per_rank_gbs": [round(x, 1) for x in per], "aggregate_gbs": round(sum(per), 1),
                      "min_gbs": round(min(per), 1),
                      "frames_per_s_ceiling": round(sum(per) * 1e9 / 98304)}
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
