"""The alternative code paths of the hot path give the same bits as the default ones:
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


def test_lu_first_packing_hostsim_bit_exact(hostsim_lib, monkeypatch):
    monkeypatch.setenv("NGB_LU_V1", "1")
    test_lu_parity.test_lu_hostsim_bit_exact_when_matrix_is(hostsim_lib)


@pytest.mark.gpu
@pytest.mark.parametrize("S", [64, 130])
def test_lu_gpu_warp_per_sample(cuda_lib, S):
    test_lu_parity._check(cuda_lib, "ro17", tol=1e-9, S=S)
