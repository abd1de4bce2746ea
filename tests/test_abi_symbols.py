"""libdabgpu.so loads without a GPU and exports every function include/*.h declares."""
import ctypes
import glob
import os
import re

import pytest

from conftest import ROOT


def _declared_functions():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        src = re.sub(r"//[^\n]*", "", src)
        src = re.sub(r"#[^\n]*", "", src)
        # drop static inline definitions (tables header) -- they are not exported symbols
        src = re.sub(r"static\s+inline[^{;]*\{", "{", src)
        depth, flat = 0, []
        for ch in src:
            if ch == "{":
                depth += 1
            elif ch == "}":
                depth -= 1
            elif depth == 0 or (depth == 1 and 'extern "C"' in src):
                flat.append(ch)
        text = "".join(flat)
        for m in re.finditer(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}()]*(?:\([^()]*\)[^;{}()]*)*\)\s*;", text):
            n = m.group(1)
            if n not in ("sizeof", "defined", "__attribute__"):
                names.add(n)
    return sorted(names)


def test_library_loads_and_exports_everything(dab):
    lib = dab.load()
    missing = []
    for name in _declared_functions():
        try:
            getattr(lib, name)
        except AttributeError:
            missing.append(name)
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_no_cpu_fallback(dab):
    """Without a device every compute entry point must fail loudly, never compute on the CPU."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(dab.DabGpuError):
        dab.fic_decode_batch(np.zeros((1, 2304), np.uint8))
    with pytest.raises(dab.DabGpuError):
        dab.viterbi_batch(np.full((1, 4 * 774), 128, np.uint8), 768)
