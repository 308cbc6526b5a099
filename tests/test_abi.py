"""The C-ABI library builds, loads and exports every symbol include/ngb200.h declares."""
import ctypes
import os
import re
import subprocess
from parity_util import ROOT


def test_header_symbols_exported():
    import importlib
    g = importlib.import_module("__graft_entry__")
    g.build()
    so = os.path.join(ROOT, "ngspice-sf-mirror_b200", "libngb200.so")
    assert os.path.exists(so)
    hdr = open(os.path.join(ROOT, "include", "ngb200.h")).read()
    names = set(re.findall(r"\b(ngb[A-Z]\w+)\s*\(", hdr))
    assert len(names) > 25
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (\w+)", out))
    missing = sorted(names - exported)
    assert not missing, missing
    # device code for sm_100a is inside
    sass = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    assert "sm_100a" in sass


def test_product_library_has_no_cpu_path():
    """without a GPU the product library refuses to create a batch (no fallback)"""
    import numpy as np
    from parity_util import pkg, ngt, GOLDEN
    import torch
    if torch.cuda.is_available():
        return
    lib = pkg.library()
    assert lib.backend == "cuda-sm_100a"
    flat = ngt.read(f"{GOLDEN}/ro17.flat.ngt")
    circ = pkg.Circuit.from_flat(lib, flat)
    try:
        pkg.Batch(circ, 1)
    except pkg.NgbError as e:
        assert "no CPU fallback" in str(e) or "CUDA" in str(e)
    else:
        raise AssertionError("batch creation must fail without a CUDA device")
