#!/usr/bin/env python
"""bench.py -- ETI frames/s of the dabtools receive hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--streams S] [--impl ours|reference]

Workload (BASELINE.json configs[2], and configs[4] when N > 1): S = 1024 independent synthetic
Mode I ensemble streams per GPU (10 UEP/EEP sub-channels, dabtools_b200.synth.reference_ensemble),
each fed exactly like dab2eti feeds the reference: 262144-byte rtlsdr callbacks, run through the
whole path: FIFO read -> synchronisers -> 76 FFTs + DQPSK + demap -> FIC Viterbi + CRC -> lock state
machine -> time de-interleave + depuncture -> MSC Viterbi + descramble -> ETI assembly.
One *step* is one pass over a batch of --tf-per-step (default 128) transmission frames per stream
(192 callbacks, 512 ETI frames per stream; 51.5 GB of samples at S = 1024), so that the driver's
20 steps are a timed region of seconds, not milliseconds.  The samples of a step come from a
24-TF capture per stream that is resident in HBM (9.7 GB) and consumed cyclically.

  value     frames/s with the IQ already resident in HBM and the ETI left in HBM
  e2e       same call path with the IQ in pinned host memory and every ETI frame copied back
  roofline  the FFT/demod kernel (HBM-bound by design; 155 904 algorithmic bytes per ETI frame)
  viterbi   ACS/s and decoded Mbit/s of the MSC Viterbi kernel (issue-bound, not HBM-bound)
  parity    after the timed passes: randomly chosen streams of the same dataset through the oracle
  cpu_baseline / --impl reference: the unmodified reference (oracle/_ref) on the host cores

Under torchrun (N > 1) every rank decodes its own S streams (weak scaling, no collective on the
data path; NCCL is only used for the barrier and the max-over-ranks of the timing).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TF_BYTES = 393216
CALL_BYTES = 262144
CALLS_PER_2TF = 3           # 3 x 262144 bytes = 2 transmission frames
FRAMES_PER_TF = 4
CAPTURE_TFS = 24            # per-stream capture resident in HBM (36 callbacks), consumed cyclically
ALGO_BYTES_PER_FRAME = 155904   # 98304 B IQ in + 57600 B demapped bits out
SETUP_TFS = 18                  # 1 start-up + 10 to lock + 4 to fill the window, plus slack


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def bench_config(S, tf_per_step):
    """`config` of the JSON line -- the same dict in both arms (ours and --impl reference): it names
    the workload, not how an arm samples it."""
    return {
        "workload": (f"BASELINE configs[2]: {S} independent Mode I ensemble streams per GPU, 10 UEP/EEP "
                     f"sub-channels (24960 decoded bits, 25026 trellis steps, 11 codewords per ETI frame), "
                     f"30 dB SNR, fed as 262144-byte rtlsdr callbacks"),
        "streams_per_gpu": S,
        "tf_per_step_per_stream": tf_per_step,
        "frames_per_step_per_gpu": S * tf_per_step * FRAMES_PER_TF,
        "snr_db": 30,
        "capture": f"{CAPTURE_TFS} TF per stream ({S * CAPTURE_TFS * TF_BYTES / 1e9:.1f} GB per GPU), consumed cyclically",
        "l2": "every step reads tens of GB of distinct samples, far beyond the 126 MB L2; no explicit flush",
    }


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def ncu_counters(kernel_prefix, grid=None):
    """Counters of a kernel from the committed `ncu --set full` capture (profiles/ncu_traffic.json,
    written by tools/summarize_ncu.py): pipe utilisations are hardware counters, not models."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["kernels"]
        rows = [r for k, v in d.items() if k.startswith(kernel_prefix) for r in v
                if grid is None or r.get("grid") == grid]
        return rows or None
    except Exception:
        return None


def ncu_traffic(kernel_prefix):
    """DRAM bytes per launch (read + write) of a kernel from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json, written by tools/summarize_ncu.py); None if absent."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["kernels"]
        rows = [r for k, v in d.items() if k.startswith(kernel_prefix) for r in v]
        return sum(r["dram_bytes_per_launch"] for r in rows) if rows else None
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(gpu_index: int):
    """Run this process (and the pinned buffers it allocates from now on) on the CPUs next to its GPU:
    pinned memory on the far socket halves the host<->device bandwidth of the end-to-end pass."""
    info = {"node": None}
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = gpu_index
        if vis:
            try:
                idx = int(vis.split(",")[gpu_index])
            except (ValueError, IndexError):
                idx = gpu_index
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bdf = bus.lower()[-12:]          # 00000000:17:00.0 -> 0000:17:00.0
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        info["node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
    except Exception as e:  # topology not exposed (containers): run unbound
        info["error"] = type(e).__name__
    return info


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).

    NVML is queried in-process (pynvml) from a thread that is already running before the warm-up, so
    that neither a process start-up nor NVML initialisation falls into the timed region; begin()/end()
    mark the region and only samples taken inside it are reported.  Falls back to an `nvidia-smi -lms`
    child started equally early."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
               ("sw_power_cap", 0x4))
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, period_s: float = 0.004):
        self.gpu = gpu_index
        self.period = period_s
        self.samples = []          # (t, sm_mhz, reasons bitmask or set)
        self.window = [None, None]
        self.stop_flag = False
        self.max_mhz = None
        self.mode = None
        self.thread = None
        self.proc = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a list of indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = self.gpu
            if vis:
                try:
                    idx = int(vis.split(",")[self.gpu])
                except (ValueError, IndexError):
                    idx = self.gpu
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.mode = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.mode = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.mode = "nvidia-smi"
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((time.perf_counter(), mhz, mask))
            except Exception:
                pass
            time.sleep(self.period)

    def _read_smi(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                mhz, mx = float(f[1]), float(f[2])
            except ValueError:
                continue
            self.max_mhz = mx
            mask = 0
            for (name, bit), v in zip(self.REASONS, f[4:8]):
                if v.lower().startswith("active"):
                    mask |= bit
            self.samples.append((time.perf_counter(), mhz, mask))

    def begin(self):
        self.window[0] = time.perf_counter()

    def end(self):
        self.window[1] = time.perf_counter()

    def stop(self) -> dict:
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        if self.thread:
            self.thread.join(timeout=1.0)
        if self.mode is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML and nvidia-smi unavailable"], "samples": 0}
        t0, t1 = self.window
        inside = [x for x in self.samples if t0 is not None and t1 is not None and t0 <= x[0] <= t1]
        note = None
        if not inside and self.samples and t0 is not None:
            # region shorter than the sampling period: take the samples nearest to it
            inside = sorted(self.samples, key=lambda x: abs(x[0] - 0.5 * (t0 + t1)))[:3]
            note = "nearest samples (region shorter than the sampling period)"
        mask = 0
        for x in inside:
            mask |= x[2]
        out = {"sm_mhz": float(np.median([x[1] for x in inside])) if inside else None, "sm_max_mhz": self.max_mhz,
               "reasons": [name for name, bit in self.REASONS if mask & bit], "samples": len(inside),
               "source": self.mode}
        if note:
            out["note"] = note
        return out


# ------------------------------------------------------------------------------------------------
# The reference on the host cores: one receiver (one stream) per worker process, state kept across
# steps (oracle ref_stream_* / orc_stream_*), so that every step is a bounded sample of the workload
# in the steady state (locked, interleaver window full) -- exactly where the GPU arm is timed.
def _ref_worker(conn, path, kind, setup_tfs, seed):
    sys.path.insert(0, ROOT)
    from oracle import oracle
    dec = {"reference": oracle.ref, "spiral": oracle.ref_spiral, "port": oracle.port}[kind]()
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 2)   # the reference prints "Locked", ensemble dumps, ...
    iq = np.load(path, mmap_mode="r")
    n_calls = iq.size // CALL_BYTES
    h = dec.stream_open(200_000_000, seed)
    pos = 0

    def feed(n_tf):
        nonlocal pos
        frames = 0
        for _ in range(n_tf * CALLS_PER_2TF // 2):
            c = pos % n_calls
            k, _e = dec.stream_feed(h, iq[c * CALL_BYTES:(c + 1) * CALL_BYTES], want_eti=False)
            frames += k
            pos += 1
        return frames

    feed(setup_tfs)
    conn.send(("ready", int(dec.stream_locked(h))))
    while True:
        msg = conn.recv()
        if msg is None:
            break
        t = time.perf_counter()
        frames = feed(msg)
        conn.send((time.perf_counter() - t, frames))
    dec.stream_close(h)
    conn.close()


class ReferencePool:
    """`cores` worker processes, each running the reference receive loop over its own copy of a
    cyclic capture of the benchmark's ensemble.  step(n_tf) = every worker decodes n_tf more
    transmission frames; returns (wall seconds, ETI frames of all workers)."""

    def __init__(self, cores, kind, setup_tfs=SETUP_TFS):
        import multiprocessing as mp
        from dabtools_b200 import synth
        ens = synth.reference_ensemble()
        g = synth.ModeITransmitter(ens, "cpu").generate(1, CAPTURE_TFS, seed=4242, snr_db=30.0)
        iq = g["iq"][0].numpy()
        tmp = tempfile.NamedTemporaryFile(suffix=".npy", delete=False,
                                          dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
        np.save(tmp, iq)
        tmp.close()
        self.path, self.kind, self.cores = tmp.name, kind, cores
        ctx = mp.get_context("spawn")
        self.conns, self.procs = [], []
        for w in range(cores):
            a, b = ctx.Pipe()
            pr = ctx.Process(target=_ref_worker, args=(b, tmp.name, kind, setup_tfs, 1 + w), daemon=True)
            pr.start()
            self.conns.append(a)
            self.procs.append(pr)
        self.locked = sum(c.recv()[1] for c in self.conns)

    def step(self, n_tf):
        t = time.perf_counter()
        for c in self.conns:
            c.send(n_tf)
        res = [c.recv() for c in self.conns]
        return time.perf_counter() - t, sum(r[1] for r in res)

    def close(self):
        for c in self.conns:
            try:
                c.send(None)
            except Exception:
                pass
        for pr in self.procs:
            pr.join(timeout=10)
        try:
            os.unlink(self.path)
        except OSError:
            pass


def reference_kind():
    from oracle import oracle
    return "reference" if oracle.ref() is not None else "port"


def cpu_reference_rate(cores, kind, n_steps, tf_per_step, warmup=1):
    """steady-state ETI frames/s of the reference on `cores` host cores (one stream per core):
    `n_steps` timed steps of `tf_per_step` transmission frames per core after `warmup` untimed ones"""
    pool = ReferencePool(cores, kind)
    try:
        for _ in range(warmup):
            pool.step(tf_per_step)
        times, frames = [], 0
        for _ in range(n_steps):
            dt, f = pool.step(tf_per_step)
            times.append(dt)
            frames += f
    finally:
        pool.close()
    total = sum(times)
    label = {"reference": "the unmodified reference (oracle/_ref, plain viterbi.c)",
             "spiral": "the unmodified reference built with its Spiral SSE2 Viterbi (oracle/_ref)",
             "port": "the oracle port (oracle/dab_oracle.c)"}[kind]
    return dict(value=frames / total, unit="frames/s", cores=cores, kind="port" if kind == "port" else "reference",
                per_core=frames / total / cores, cpu_model=cpu_model(), locked_streams=pool.locked,
                steps=n_steps, step_s=float(np.median(times)), total_s=total, frames=frames,
                sample=f"{label}: 1 stream per core, {cores} cores, {n_steps} steps x {tf_per_step} TF per core "
                       f"in the steady state (locked, interleaver window full) of a cyclic {CAPTURE_TFS}-TF "
                       f"capture of the same ensemble")


# ------------------------------------------------------------------------------------------------
def shard_streams(total_streams: int, world: int, rank: int):
    """Streams are independent end to end: rank r decodes the contiguous block
    [r*total/world, (r+1)*total/world) and nothing is exchanged between ranks."""
    per = total_streams // world
    rem = total_streams % world
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def reduce_over_ranks(times_ms, frames, device, world):
    """timing = MAX over ranks, work = SUM over ranks (the only collectives of the benchmark; the
    data path has none).  Works on any backend (nccl on the GPU box, gloo in the CPU tests)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(times_ms), dtype=torch.float64, device=device)
    f = torch.tensor(list(frames), dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(f, op=dist.ReduceOp.SUM)
    return [float(x) for x in t], [float(x) for x in f]


def generate_dataset(S, n_tf, device, seed):
    """[S][n_tf*393216] uint8 I/Q on `device`, distinct payload and noise per stream"""
    import torch
    from dabtools_b200 import synth
    ens = synth.reference_ensemble()
    tx = synth.ModeITransmitter(ens, device)
    total = n_tf * TF_BYTES
    out = torch.empty((S, total), dtype=torch.uint8, device=device)
    chunk = 32
    for s0 in range(0, S, chunk):
        n = min(chunk, S - s0)
        g = tx.generate(n, n_tf, seed=seed * 100003 + s0, snr_db=30.0)
        out[s0:s0 + n] = g["iq"]
        del g
    return out, ens


def parity_check(lib, data, S, n_streams, seed):
    """BASELINE config 3 parity, outside every timed region: `n_streams` randomly chosen streams of
    the benchmark's own dataset, decoded (a) by an engine that consumes the capture in place with
    MSC batches of 2 -- the `value` configuration -- and (b) by an engine fed from host memory with
    MSC batches of 4 -- the `e2e` configuration --, compared byte for byte with the oracle's receive
    loop (oracle.port().run_iq) over the same samples."""
    from oracle import oracle
    port = oracle.port()
    n_calls = data.shape[1] // CALL_BYTES
    rng = np.random.default_rng(seed)
    chosen = sorted(rng.choice(S, min(n_streams, S), replace=False).tolist())

    def collect(eng, feed):
        out = {s: [] for s in chosen}
        for c in range(n_calls):
            if feed(c):
                eti, ids = eng.fetch_eti()
                for s in chosen:
                    out[s].extend(f.copy() for f in eti[ids == s])
        if eng.flush():
            eti, ids = eng.fetch_eti()
            for s in chosen:
                out[s].extend(f.copy() for f in eti[ids == s])
        eng.close()
        return out

    eng = lib.Engine(S)
    eng.set_msc_batch(2)
    eng.attach_capture(data)
    a = collect(eng, lambda c: eng.feed_capture(CALL_BYTES))
    host = {s: data[s].cpu().numpy() for s in chosen}
    import torch
    pin = torch.empty((S, CALL_BYTES), dtype=torch.uint8, pin_memory=True)
    eng = lib.Engine(S)
    eng.set_msc_batch(4)

    def feed_host(c):
        pin.copy_(data[:, c * CALL_BYTES:(c + 1) * CALL_BYTES])
        torch.cuda.synchronize()
        return eng.feed_iq(pin.numpy())

    b = collect(eng, feed_host)
    frames = 0
    for s in chosen:
        want = port.run_iq(host[s][: n_calls * CALL_BYTES])["eti"]
        for name, got in (("capture/batch2", a[s]), ("host/batch4", b[s])):
            got = np.array(got, dtype=np.uint8).reshape(-1, 6144)
            if got.shape != want.shape or not np.array_equal(got, want):
                raise RuntimeError(f"parity: stream {s} ({name}): ETI differs from the oracle "
                                   f"({got.shape[0]} vs {want.shape[0]} frames)")
        frames += want.shape[0]
    return {"parity_checked_streams": len(chosen), "streams": chosen, "eti_frames_compared": frames,
            "engines": ["attach_capture + msc_batch 2 (value path)", "host feed_iq + msc_batch 4 (e2e path)"],
            "against": "oracle.port().run_iq on the same samples", "result": "byte-identical"}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from dabtools_b200 import lib, synth

    numa = bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    lib.check(lib.load().dabgpu_set_device(local_rank))
    lib.use_torch_stream()
    S, K, W = args.streams, args.steps, args.warmup
    TFS = args.tf_per_step
    TFS += TFS % 2
    calls_per_step = TFS // 2 * CALLS_PER_2TF
    setup_calls = SETUP_TFS // 2 * CALLS_PER_2TF
    cap_calls = CAPTURE_TFS // 2 * CALLS_PER_2TF
    t_gen = time.time()
    data, ens = generate_dataset(S, CAPTURE_TFS, dev, seed=1 + rank)
    torch.cuda.synchronize()
    t_gen = time.time() - t_gen

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def chunk(c):
        off = (c % cap_calls) * CALL_BYTES
        return data[:, off: off + CALL_BYTES]

    # ---------------- device-resident throughput (value) ----------------
    # The samples are resident in HBM before the timed region.  With --ingest capture (default) the
    # engine consumes them in place (dabgpu_engine_attach_capture / feed_capture, cyclic); with
    # --ingest copy every callback is first copied into the engine's own FIFO ring (feed_iq).
    capture = args.ingest == "capture"

    def measure_value():
        eng = lib.Engine(S)
        eng.set_msc_batch(args.msc_batch)
        if capture:
            eng.attach_capture(data)
            eng.set_capture_cyclic(True)
        fed = [0]

        def feed_calls(n):
            frames = 0
            for _ in range(n):
                frames += eng.feed_capture(CALL_BYTES) if capture else eng.feed_iq_device(chunk(fed[0]))
                fed[0] += 1
            return frames

        feed_calls(setup_calls)
        locked = sum(eng.status(s).locked for s in range(S))
        if locked != S:
            raise RuntimeError(f"only {locked}/{S} streams locked after set-up")
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        n = 0
        for i in range(W):
            n += feed_calls(calls_per_step)
        assert n >= S * FRAMES_PER_TF * (W * TFS - 4), f"steady state not reached: {n} frames in warm-up"
        eng.flush()   # start from an empty pipeline so that the frames counted are exactly the frames fed
        eng.join()
        launches0 = lib.launch_count()
        host_t0 = eng.host_times()
        barrier()
        sampler.begin()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        frames = 0
        for i in range(K):
            frames += feed_calls(calls_per_step)
        frames += eng.flush()   # frames still queued for a deferred MSC batch belong to these steps
        eng.join()              # ... and so does the MSC stream's last batch
        e1.record()
        barrier()
        sampler.end()
        clocks = sampler.stop() if rank == 0 else None
        ms = e0.elapsed_time(e1)
        launches = lib.launch_count() - launches0
        host_t = eng.host_times()
        eng.close()
        return dict(ms=ms, frames=frames, clocks=clocks, launches=launches, host_t=host_t, host_t0=host_t0)

    res = measure_value()
    # a run that saw thermal / hardware slowdown is rejected and measured again, once
    bad_reasons = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    redo = torch.zeros(1, dtype=torch.int32, device=dev)
    if rank == 0 and res["clocks"] and (bad_reasons & set(res["clocks"].get("reasons", []))
                                        or os.environ.get("DABGPU_BENCH_FORCE_REDO")):   # (test hook)
        redo += 1
    if world > 1:
        dist.broadcast(redo, src=0)
    if int(redo.item()):
        rejected = res["clocks"]
        res = measure_value()
        if rank == 0 and res["clocks"] is not None:
            res["clocks"]["rejected_first_run"] = rejected
    ms, frames, clocks, launches = res["ms"], res["frames"], res["clocks"], res["launches"]
    host_t, host_t0 = res["host_t"], res["host_t0"]

    # ---------------- per-kernel timing pass (copying path: it has the ingest kernel) ----------------
    k_timing = 8
    eng = lib.Engine(S)
    eng.set_msc_batch(args.msc_batch)
    for c in range(setup_calls + 6):
        eng.feed_iq_device(chunk(c))
    eng.enable_timing(True)
    for c in range(setup_calls + 6, setup_calls + 6 + k_timing * CALLS_PER_2TF):
        eng.feed_iq_device(chunk(c))
    kt = eng.kernel_times()
    eng.enable_timing(False)
    eng.close()
    del eng

    # ---------------- end to end: pinned host IQ in, ETI out to host ----------------
    eng = lib.Engine(S)
    eng.set_msc_batch(args.e2e_msc_batch)
    for c in range(setup_calls):
        eng.feed_iq_device(chunk(c))
    host_cap = torch.empty((cap_calls, S, CALL_BYTES), dtype=torch.uint8, pin_memory=True)
    for c in range(cap_calls):
        host_cap[c].copy_(chunk(c))
    torch.cuda.synchronize()
    host_out_t = torch.empty((S * FRAMES_PER_TF * (args.e2e_msc_batch + 1), 6144), dtype=torch.uint8,
                             pin_memory=True)
    host_out = host_out_t.numpy()
    host_calls = [host_cap[c].numpy() for c in range(cap_calls)]
    e2e_tfs = args.e2e_tfs - args.e2e_tfs % (2 * args.e2e_msc_batch)   # whole MSC batches in the window
    w_calls, c_calls = 2 * CALLS_PER_2TF * args.e2e_msc_batch // 2, CALLS_PER_2TF
    t_calls = e2e_tfs // 2 * CALLS_PER_2TF
    n_calls_e2e = w_calls + t_calls + c_calls
    AHEAD = 2   # uploads in flight ahead of the callback being processed

    # Public API, software-pipelined.  The timed callbacks sit between warm-up and cool-down
    # callbacks that are fed the same way, so the timed window starts and ends in the same pipeline
    # state (uploads in flight at both ends) and holds exactly its own share of the work: 4 ETI
    # frames per transmission frame and stream.
    def call(k):
        return host_calls[(setup_calls + k) % cap_calls]

    first_timed, last_timed = w_calls, w_calls + t_calls
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    d2h = 0
    for k in range(min(AHEAD, n_calls_e2e)):
        eng.submit_iq(call(k))
    for k in range(n_calls_e2e):
        if k == first_timed:
            if world > 1:
                dist.barrier()
            f0.record()
        if k + AHEAD < n_calls_e2e:
            eng.submit_iq(call(k + AHEAD))
        n = eng.feed_submitted()
        if n:
            eng.fetch_eti(host_out)
            if first_timed <= k < last_timed:
                d2h += n * 6144
        if k == last_timed - 1:
            f1.record()
    if eng.flush():
        eng.fetch_eti(host_out)
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    e2e_frames = e2e_tfs * FRAMES_PER_TF * S
    # spot check: the frames really are ETI (sync word, padding) -- guards against timing nothing
    assert host_out[0, 0] == 0xFF and host_out[0, 1] in (0x07, 0xF8) and host_out[0, -1] == 0x55
    eng.close()
    del host_cap, host_calls

    # ---------------- BASELINE config 2: FIC-only decode, 16384 groups, device resident ----------------
    fic_cfg = None
    if rank == 0:
        n_grp = 16384
        fic_np, sent = synth.fic_groups(n_grp, seed=2)     # 1/4 clean, 3/4 at 1 / 4 / 8 % bit flips
        fic_bits = torch.from_numpy(fic_np).to(dev)
        fibs = torch.zeros((n_grp, 96), dtype=torch.uint8, device=dev)
        okf = torch.zeros((n_grp, 3), dtype=torch.uint8, device=dev)
        for _ in range(3):
            lib.fic_decode_batch_device(fic_bits, fibs, okf)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(20):
            lib.fic_decode_batch_device(fic_bits, fibs, okf)
        c1.record()
        torch.cuda.synchronize()
        fic_ms = c0.elapsed_time(c1) / 20
        fic_cfg = {"groups": n_grp, "ms_per_batch": fic_ms, "decoded_mbit_s": n_grp * 768 / fic_ms / 1e3,
                   "acs_per_s": n_grp * 774 * 64 / (fic_ms * 1e-3),
                   "input": "dabtools_b200.synth.fic_groups(16384, seed=2): 1/4 clean, 3/4 with 1 / 4 / 8 % bit flips",
                   "note": "depuncture + Viterbi + descramble + FIB CRC of 16384 FIC groups (BASELINE configs[1])"}
        if not args.no_parity:
            # outside the timed loop: every FIB and CRC flag against the oracle (fic.c:185-206)
            from oracle import oracle
            port = oracle.port()
            got_f, got_ok = fibs.cpu().numpy(), okf.cpu().numpy()
            for g in range(0, n_grp, 4):
                f, c, _ = port.fic_decode(fic_np[g:g + 4].reshape(-1))
                if not (np.array_equal(f.reshape(4, 96), got_f[g:g + 4]) and
                        np.array_equal(c.reshape(4, 3), got_ok[g:g + 4])):
                    raise RuntimeError(f"fic_only: group {g} differs from the oracle")
            assert np.array_equal(got_f[: n_grp // 4], sent[: n_grp // 4])
            fic_cfg["parity"] = f"all {n_grp} groups (FIBs and CRC flags) byte-identical with oracle.port().fic_decode"
            fic_cfg["crc_ok_rate_per_quarter"] = [float(got_ok[q * n_grp // 4:(q + 1) * n_grp // 4].mean())
                                                  for q in range(4)]
        del fic_bits, fibs, okf

    # ---------------- opt-in soft-decision mode: same call path, a short pass ----------------
    soft_cfg = None
    if rank == 0 and not args.no_soft:
        eng = lib.Engine(S, 200_000_000, lib.ENGINE_SOFT)
        eng.set_msc_batch(args.msc_batch)
        for c in range(setup_calls + 6):
            eng.feed_iq_device(chunk(c))
        eng.flush()
        eng.join()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        n_soft = 0
        for c in range(setup_calls + 6, setup_calls + 6 + 8 * CALLS_PER_2TF):
            n_soft += eng.feed_iq_device(chunk(c))
        n_soft += eng.flush()
        eng.join()
        s1.record()
        torch.cuda.synchronize()
        soft_cfg = {"value": n_soft / (s0.elapsed_time(s1) * 1e-3), "unit": "frames/s", "frames": n_soft,
                    "locked_streams": sum(eng.status(s).locked for s in range(S)),
                    "note": "DABGPU_ENGINE_SOFT (soft demapper, symbol data path, 16-bit-metric Viterbi), copying "
                            "feed_iq path, 16 TF per stream; not the headline configuration"}
        eng.close()
        del eng

    # ---------------- BASELINE config 3 parity on this very dataset ----------------
    parity = None
    if rank == 0 and not args.no_parity:
        parity = parity_check(lib, data, S, args.parity_streams, seed=7)

    # ---------------- reduce over ranks ----------------
    (ms_max, e2e_ms_max), (frames_all, e2e_frames_all) = reduce_over_ranks(
        [ms, e2e_ms], [frames, e2e_frames], dev, world)
    if rank != 0:
        return None

    hbm_peak, peak_src = peaks()
    demod = kt["demod"]
    frames_per_demod = S * FRAMES_PER_TF
    frames_per_msc = frames_per_demod * args.msc_batch
    demod_ms = demod["ms"] / max(demod["launches"], 1)
    achieved = ALGO_BYTES_PER_FRAME * frames_per_demod / (demod_ms * 1e-3) / 1e9 if demod_ms > 0 else 0.0
    vit = kt["msc_viterbi"]
    vit_ms = vit["ms"] / max(vit["launches"], 1)
    msc_steps_per_launch = (ens.steps_per_frame - 774) * frames_per_msc
    msc_bits_per_launch = (ens.bits_per_frame - 768) * frames_per_msc
    demod_ncu = ncu_counters("demod_kernel")
    vit_ncu = ncu_counters("viterbi_kernel")
    step_frames = S * TFS * FRAMES_PER_TF
    out = {
        "metric": "ETI frames/s (Mode I)",
        "value": frames_all / (ms_max * 1e-3),
        "unit": "frames/s",
        "n_gpus": world,
        "steps": K,
        "warmup": W,
        "ms_per_step": ms_max / K,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u8 in / fp32 FFT / u8 path metrics",
        "data": "synthetic",
        "config": bench_config(S, TFS),
        "run": {
            "timed_region_s": ms_max / 1e3,
            "frames_per_step": step_frames * world,
            "msc_batch_tf": args.msc_batch,
            "ingest": ("in place: dabgpu_engine_attach_capture + feed_capture, cyclic (the samples are consumed "
                       "where they lie)" if capture else
                       "copy: dabgpu_engine_feed_iq (every callback is copied into the engine's FIFO ring first)"),
            "timing": "CUDA events around the K steps (pipeline empty at both ends: flush + join), barrier + "
                      "synchronize on both sides, max over ranks",
            "kernel_timing": f"per-kernel CUDA events over {k_timing} x 2 TF following a set-up on a second engine, "
                             "engine streams serialised so that each kernel runs alone",
            "dataset_gen_s": round(t_gen, 1),
            "numa_binding": numa,
        },
        "clocks": clocks,
        "gpu_launches": int(launches),
        "e2e": {
            "value": e2e_frames_all / (e2e_ms_max * 1e-3),
            "unit": "frames/s",
            "h2d_bytes_per_step": S * calls_per_step * CALL_BYTES,
            "d2h_bytes_per_step": int(d2h * TFS / max(e2e_tfs, 1)),
            "timed_tf_per_stream": e2e_tfs,
            "timed_region_s": e2e_ms_max / 1e3,
            "steps": e2e_tfs / TFS,
            "msc_batch_tf": args.e2e_msc_batch,
            "api": "dabgpu_engine_submit_iq / feed_submitted / fetch_eti from pinned host memory (uploads run two "
                   "callbacks ahead); h2d/d2h bytes are per step of `config`, the timed window is "
                   "timed_tf_per_stream transmission frames per stream",
        },
        "roofline": {
            "kernel": "demod_kernel (FFT2048 x76 + DQPSK + freq de-interleave + slicing)",
            "bound": "hbm",
            "achieved": achieved,
            "peak": hbm_peak,
            "unit": "GB/s",
            "frac": achieved / hbm_peak,
            "frac_of_nominal_8tbs": achieved / 8000.0,
            # demod is launched as two grids per frame (FIC symbols, CIF symbols): both together
            "traffic": sum(r["dram_bytes_per_launch"] for r in demod_ncu) if demod_ncu else None,
            "traffic_source": "profiles/ncu_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)",
            "peak_source": peak_src,
            "note": "designed HBM-bound (155 904 algorithmic bytes per ETI frame); in practice bound by the "
                    "shared-memory data pipe of the FFT exchanges (see ncu)",
            "ms_per_launch": demod_ms,
            "algorithmic_bytes_per_launch": ALGO_BYTES_PER_FRAME * frames_per_demod,
            "ncu": demod_ncu,
        },
        "viterbi": {
            "kernel": "viterbi_kernel (MSC)",
            "ms_per_launch": vit_ms,
            "acs_per_s": 64.0 * msc_steps_per_launch / (vit_ms * 1e-3) if vit_ms > 0 else 0.0,
            "decoded_mbit_s": msc_bits_per_launch / (vit_ms * 1e-3) / 1e6 if vit_ms > 0 else 0.0,
            "ncu": vit_ncu,   # ALU-pipe / issue-slot utilisation: hardware counters of the committed capture
        },
        "fic_only": fic_cfg,
        "soft_mode": soft_cfg,
        "parity": parity,
        "host_ms_per_2tf": {k: (host_t[k] - host_t0[k]) / 1e3 / (K * TFS / 2) for k in host_t},
        "kernel_ms_per_launch": {k: (v["ms"] / v["launches"] if v["launches"] else None) for k, v in kt.items()},
    }
    return out


def run_reference(args, rank, world):
    """--impl reference: the unmodified reference (oracle/_ref) on all host cores, same metric and
    config; every step is a bounded sample of the workload (REF_TF_PER_STEP transmission frames per
    core in the steady state).  Nothing of the product is imported or mapped here."""
    if rank != 0:
        return None
    os.environ["DABGPU_FORBID_LOAD"] = "1"   # inherited by the workers: libdabgpu.so stays unmapped
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    K, W = args.steps, args.warmup
    tf = int(os.environ.get("DABGPU_BENCH_REF_TFS", str(args.ref_tf_per_step)))
    tf += tf % 2
    kind = reference_kind()
    t_wall = time.perf_counter()
    r = cpu_reference_rate(cores, kind, n_steps=K, tf_per_step=tf, warmup=W)
    extra = {}
    if not args.no_spiral and kind == "reference":
        try:
            sp = cpu_reference_rate(cores, "spiral", n_steps=max(2, min(K, 5)), tf_per_step=tf, warmup=min(W, 1))
            extra["spiral_sse2"] = {k: sp[k] for k in ("value", "unit", "cores", "per_core", "steps", "sample")}
        except Exception as ex:
            extra["spiral_sse2"] = {"value": None, "error": str(ex)}
    return {
        "impl": "reference",
        "metric": "ETI frames/s (Mode I)",
        "value": r["value"],
        "unit": "frames/s",
        "n_gpus": world,
        "steps": K,
        "warmup": W,
        "ms_per_step": r["total_s"] * 1e3 / max(K, 1),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "u8 in / f64 FFT / long path metrics",
        "data": "synthetic",
        "config": bench_config(args.streams, args.tf_per_step + args.tf_per_step % 2),
        "run": {"timed_region_s": r["total_s"], "wall_s": time.perf_counter() - t_wall,
                "frames_per_step": r["frames"] / max(K, 1), "tf_per_step_per_core": tf,
                "reference_arm": "the unmodified reference receive loop (sdr_demod -> dab_process_frame, compiled "
                                 "from /root/reference/src into oracle/_ref) on the host cores; every step is a "
                                 "bounded sample of the workload of `config`: " + r["sample"]},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "per_core", "cpu_model", "sample")},
        "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        **extra,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--streams", type=int, default=1024, help="ensemble streams per GPU")
    ap.add_argument("--tf-per-step", type=int, default=128,
                    help="transmission frames per stream and step (one step = one pass over that batch)")
    ap.add_argument("--e2e-tfs", type=int, default=288,
                    help="transmission frames per stream in the timed window of the pinned-host pass")
    ap.add_argument("--msc-batch", type=int, default=2,
                    help="transmission frames per MSC Viterbi launch (dabgpu_engine_set_msc_batch)")
    ap.add_argument("--e2e-msc-batch", type=int, default=4,
                    help="same for the end-to-end pass (larger batches amortise the PCIe round trips)")
    ap.add_argument("--ingest", default="capture", choices=["capture", "copy"],
                    help="device-resident pass: consume the samples in place, or copy each callback into the FIFO ring")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-tf-per-step", type=int, default=8,
                    help="reference arm / cpu_baseline: transmission frames per core and step")
    ap.add_argument("--parity-streams", type=int, default=8)
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-spiral", action="store_true")
    ap.add_argument("--no-soft", action="store_true", help="skip the short pass in the opt-in soft-decision mode")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        out = run_reference(args, rank, world)
        if out is not None:
            print(json.dumps(out))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL is only used for the barrier and the max-over-ranks; keep its version banner (printed to
        # stdout under NCCL_DEBUG=VERSION) away from the one JSON line
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    out = run_ours(args, rank, world, local_rank)
    if rank == 0:
        if not args.no_cpu_baseline:
            cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
            try:
                kind = reference_kind()
                r = cpu_reference_rate(cores, kind, n_steps=3, tf_per_step=args.ref_tf_per_step)
                out["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "per_core", "cpu_model",
                                                         "sample")}
                if kind == "reference" and not args.no_spiral:
                    sp = cpu_reference_rate(cores, "spiral", n_steps=3, tf_per_step=args.ref_tf_per_step)
                    out["cpu_baseline"]["spiral_sse2"] = {k: sp[k] for k in ("value", "per_core", "sample")}
            except Exception as ex:  # the baseline is reported context, never the product path
                out["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": cores, "kind": "unavailable",
                                       "sample": f"failed: {ex}"}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
