"""The alternative code paths of the hot path give the same bits as the default ones:
  * phase-split BSIM4 load (four kernels, B4W fields through the scratch array; NGB_B4_SPLIT=1)
    against the recorded reference calls, and its generated field lists are current;
  * first LU packing (NGB_LU_V1=1, kept for comparison) against KLU's recorded factors;
  * warp-per-sample LU launch geometry (S >= 64) on the GPU."""
import os
import subprocess
import sys
import numpy as np
import pytest
import test_load_parity
import test_lu_parity
from parity_util import GOLDEN, ngt, pkg, first_pattern, replay_load, trace_calls

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_field_lists_are_current():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "b4_split_lists.py"), "--check"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("name", ["ro17", "inv", "mix"])
def test_split_load_hostsim_matches_reference(hostsim_lib, monkeypatch, name):
    monkeypatch.setenv("NGB_B4_SPLIT", "1")
    if name in test_load_parity.DUAL:
        test_load_parity._check(hostsim_lib, name, tol_state=0.0, tol_mat=1.0, tol_scaled=1e-12)
    else:
        test_load_parity._check(hostsim_lib, name, tol_state=0.0, tol_mat=1e-14)


def test_split_load_hostsim_same_bits_as_one_kernel(hostsim_lib, monkeypatch):
    """every array the load writes, split against unsplit, on every recorded call of the 17-stage oscillator"""
    lib = hostsim_lib
    flat = ngt.read(f"{GOLDEN}/ro17.flat.ngt"); trace = ngt.read(f"{GOLDEN}/ro17.trace.ngt.gz")
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=first_pattern(trace))
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("NGB_B4_SPLIT", mode)
        batch = None
        for call in trace_calls(trace):
            batch, ours, ref, maps = replay_load(lib, circ, flat, trace, call, S=3, batch=batch)
            out[mode, call] = ours
        assert (lib.L.ngbBatchArrayBytes(batch.h, b"b4.wscr") > 0) == (mode == "1")
    for (mode, call), ours in out.items():
        if mode == "1":
            for k in ("Ax", "x", "b4_state", "b4_op", "noncon"):
                assert np.array_equal(np.asarray(ours[k]), np.asarray(out["0", call][k]), equal_nan=True), (call, k)


def test_lu_first_packing_hostsim_bit_exact(hostsim_lib, monkeypatch):
    monkeypatch.setenv("NGB_LU_V1", "1")
    test_lu_parity.test_lu_hostsim_bit_exact_when_matrix_is(hostsim_lib)


@pytest.mark.gpu
@pytest.mark.parametrize("name,S", [("ro17", 67), ("mix", 33)])
def test_split_load_gpu(cuda_lib, monkeypatch, name, S):
    monkeypatch.setenv("NGB_B4_SPLIT", "1")
    test_load_parity._check(cuda_lib, name, tol_state=1e-9, tol_mat=1.0 if name in test_load_parity.DUAL else 1e-9, S=S, tol_scaled=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("S", [64, 130])
def test_lu_gpu_warp_per_sample(cuda_lib, S):
    test_lu_parity._check(cuda_lib, "ro17", tol=1e-9, S=S)
