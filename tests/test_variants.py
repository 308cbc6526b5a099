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


def _overlay_run(lib, monkeypatch, overlay):
    from parity_util import run_patterns
    if not overlay:
        monkeypatch.setenv("NGB_B4_OVERLAY", "0")
    base = ngt.read(f"{GOLDEN}/ro17k.flat.ngt"); trace = ngt.read(f"{GOLDEN}/ro17k.trace.ngt.gz")
    tab = ngt.read(f"{GOLDEN}/b4temp.tables.ngt.gz")
    raw = {"model": tab["ro17k/b4t/model"], "inst": tab["ro17k/b4t/inst"], "inst_model": tab["ro17k/b4t/inst_model"],
           "temp": tab["ro17k/b4t/temp"][0, 0], "vt0": tab["ro17k/opt/vt0"][0]}
    S = 5
    toxe = 1.4e-9 * (1.0 + 0.03 * np.random.default_rng(7).normal(size=S))
    toxe[3] = toxe[0]                                  # two samples with the same oxide, one of them sample 0
    dv = pkg.mc.draw_delvto(S, 34, seed=11)
    inst, prow_t, mtab, ptab = pkg.mc.bsim4_with_toxe(lib, raw, toxe, dv)
    b = pkg.Batch(pkg.Circuit.from_flat(lib, base, lu_pattern=run_patterns(trace)), S)
    b.put("b4.inst", inst)
    b.set_bsim4_rows(prow_t, mtab, ptab)
    res = b.tran(1024, [18])
    return b, res, res.waves()


def test_per_sample_rows_as_overlay_same_bits_hostsim(hostsim_lib, monkeypatch):
    """ngbBatchSetBsim4Rows: the load reads only the columns that differ between the samples of an instance at the thread's own
    row (oxide thickness: 3 model + 12 bin columns of 221) and the rest at the row of sample 0; same bits as reading every
    column at the own row (NGB_B4_OVERLAY=0), with the specialised and with the generic kernel"""
    b1, r1, (t1, v1) = _overlay_run(hostsim_lib, monkeypatch, True)
    assert b1.bsim4_overlay() == (3, 12) and b1.bsim4_variant()[1]
    b0, r0, (t0, v0) = _overlay_run(hostsim_lib, monkeypatch, False)
    assert b0.bsim4_overlay() is None and b0.bsim4_variant()[1] and b0.bsim4_variant()[0] != b1.bsim4_variant()[0]
    assert np.array_equal(r0.npoints, r1.npoints) and np.array_equal(r0.numiter, r1.numiter)
    assert np.array_equal(t0, t1) and np.array_equal(v0, v1)
    assert len({int(n) for n in r1.numiter}) > 1         # the samples really differ


@pytest.mark.gpu
def test_per_sample_rows_as_overlay_same_bits_gpu(cuda_lib, monkeypatch):
    b1, r1, (t1, v1) = _overlay_run(cuda_lib, monkeypatch, True)
    assert b1.bsim4_overlay() == (3, 12) and b1.bsim4_variant()[1]
    b1.set_bsim4_generic(True)
    rg = b1.tran(1024, [18]); tg, vg = rg.waves()
    b0, r0, (t0, v0) = _overlay_run(cuda_lib, monkeypatch, False)
    assert b0.bsim4_overlay() is None
    for r, t, v in ((r0, t0, v0), (rg, tg, vg)):
        assert np.array_equal(r.npoints, r1.npoints) and np.array_equal(r.numiter, r1.numiter)
        assert np.array_equal(t, t1) and np.array_equal(v, v1)


def test_overlay_rows_in_any_order_hostsim(hostsim_lib, monkeypatch):
    """the overlay addresses a thread's rows by their distance from the rows of the instance's sample 0: the same tables with
    their rows permuted (negative distances, sample 0 no longer at the lowest row) give the same bits"""
    from parity_util import run_patterns
    base = ngt.read(f"{GOLDEN}/ro17k.flat.ngt"); trace = ngt.read(f"{GOLDEN}/ro17k.trace.ngt.gz")
    tab = ngt.read(f"{GOLDEN}/b4temp.tables.ngt.gz")
    raw = {"model": tab["ro17k/b4t/model"], "inst": tab["ro17k/b4t/inst"], "inst_model": tab["ro17k/b4t/inst_model"],
           "temp": tab["ro17k/b4t/temp"][0, 0], "vt0": tab["ro17k/opt/vt0"][0]}
    S = 4
    toxe = 1.4e-9 * (1.0 + 0.03 * np.random.default_rng(3).normal(size=S))
    dv = pkg.mc.draw_delvto(S, 34, seed=12)
    inst, prow_t, mtab, ptab = pkg.mc.bsim4_with_toxe(hostsim_lib, raw, toxe, dv)
    circ = pkg.Circuit.from_flat(hostsim_lib, base, lu_pattern=run_patterns(trace))
    out = []
    perm = np.random.default_rng(4).permutation(mtab.shape[0])
    inv = np.empty_like(perm); inv[perm] = np.arange(len(perm))
    for p_rows, m, p in ((prow_t, mtab, ptab), (inv[prow_t].astype(np.int32), mtab[perm], ptab[perm])):
        b = pkg.Batch(circ, S)
        b.put("b4.inst", inst)
        b.set_bsim4_rows(p_rows, m, p)
        assert b.bsim4_overlay() == (3, 12)
        res = b.tran(512, [18])
        out.append((res.npoints.copy(), res.numiter.copy()) + res.waves())
    assert (inv[prow_t].reshape(34, S)[:, 1:] < inv[prow_t].reshape(34, S)[:, :1]).any()       # some distances are negative
    for a, c in zip(out[0], out[1]):
        assert np.array_equal(a, c)
