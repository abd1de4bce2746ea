"""Build libdabgpu.so (CUDA, sm_100a) in-tree with nvcc.

    python -m dabtools_b200.build [--force]

Objects go to dabtools_b200/build/, the library to dabtools_b200/libdabgpu.so (git-ignored, but
shipped to the GPU box with the working tree).  Only sm_100a SASS is generated: there is no PTX
fallback, no other architecture and no CPU path.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libdabgpu.so")
TABLES_LIB = os.path.join(HERE, "libdabtables.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
]


def _newer(src_list, target):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in src_list)


def build_variant(name: str, extra_flags: list[str]) -> str:
    """Kernel experiments: a second library built with extra nvcc flags (e.g. -DDABGPU_VIT_WARPS=8),
    selected at run time with DABGPU_LIB=<path>.  Not used by the product."""
    obj = os.path.join(HERE, "build", "variant_" + name)
    os.makedirs(obj, exist_ok=True)
    lib = os.path.join(HERE, f"libdabgpu_{name}.so")
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    objs = [os.path.join(obj, os.path.basename(s)[:-3] + ".o") for s in sources]

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            raise RuntimeError("nvcc failed")

    with ThreadPoolExecutor(max_workers=8) as ex:
        list(ex.map(run, [[NVCC] + NVCC_FLAGS + extra_flags + ["-c", s, "-o", o] for s, o in zip(sources, objs)]))
    run([NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                               "-Xlinker", "-Bsymbolic", "-lpthread"])
    return lib


def build_tables(force: bool = False) -> str:
    """libdabtables.so: the dabgpu_tab_* accessors alone, host-only (gcc, no CUDA); see csrc/tables_host.c"""
    src = os.path.join(CSRC, "tables_host.c")
    hdr = os.path.join(HERE, "..", "include", "dabgpu_tables.h")
    if force or _newer([src, hdr], TABLES_LIB):
        subprocess.check_call([os.environ.get("CC", "gcc"), "-O2", "-fPIC", "-shared", "-fvisibility=hidden",
                               "-std=gnu11", "-o", TABLES_LIB, src])
    return TABLES_LIB


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    build_tables(force)
    sources = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    headers = sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) +
                     glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    jobs = []
    objs = []
    for s in sources:
        o = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or _newer([s] + headers, o):
            jobs.append([NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for " + cmd[-3])
        return r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(run, jobs))
    stale_objs = set(glob.glob(os.path.join(OBJ, "*.o"))) - set(objs)
    for o in stale_objs:
        os.remove(o)
    if jobs or stale_objs or force or _newer(objs, LIB):
        run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                  "-Xlinker", "-Bsymbolic", "-lpthread"])
    return LIB


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--variant":
        print(build_variant(sys.argv[2], sys.argv[3:]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
