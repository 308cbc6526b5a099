"""The parity check bench.py runs on its own output (bench.parity_check): fresh mismatch draws -- not the committed
fixtures -- simulated by the reference binary (oracle/_ref/ngspice, full `.tran .1ns 150ns uic`) and by the batch
path, compared per accepted time point by SURVEY.md section 8(d)'s rule: identical point count, 1e-9 * max(|ref|, vntol)."""
import importlib
import os
import numpy as np
import pytest
from parity_util import GOLDEN, ROOT, ngt, pkg, run_patterns

REF = os.path.join(ROOT, "oracle", "_ref", "ngspice")


def _draws_vs_reference(lib, S, seed):
    bench = importlib.import_module("bench")
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/ngspice not built")
    flat = ngt.read(f"{GOLDEN}/ro17k.flat.ngt")
    flat["tran/tstop"] = np.array([pkg.mc.spice_number("150ns")])
    trace = ngt.read(f"{GOLDEN}/ro17k.trace.ngt.gz")
    wave = ngt.read(f"{GOLDEN}/ro17k.wave.ngt")
    tables = ngt.read(f"{GOLDEN}/ro17tox.tables.ngt")
    ninst = int(flat["b4/ninst"][0])
    dv_raw = pkg.mc.draw_delvto(S, ninst, sigma=0.015, seed=seed)
    dv = pkg.mc.delvto_as_parsed(dv_raw)
    level = np.random.default_rng(seed + 1).integers(0, len(tables["levels"]), size=S)
    inst, prow_t, mtab, ptab = pkg.mc.bsim4_with_tox_levels(lib, flat, tables, level, dv)
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=run_patterns(trace))
    b = pkg.Batch(circ, S)
    b.put("b4.inst", inst)
    b.set_bsim4_rows(prow_t, mtab, ptab)
    clauses = [(int(wave["save_eq"][0]), 0, 10, 0.5, 0.0), (int(wave["save_eq"][0]), 0, 20, 0.5, 0.0)]     # ro_17_4.cir:54
    b.set_measures(clauses)
    res = b.tran(6144, wave["save_eq"][:1])
    assert not res.err.any()
    t, v = res.waves()
    ms = res.measures()
    cpu = bench.cpu_reference_run("mc_ro17", min(S, os.cpu_count() or 1), (S + (os.cpu_count() or 1) - 1) // (os.cpu_count() or 1) if S > (os.cpu_count() or 1) else 1,
                                  draws=[(dv_raw[s], float(tables["levels"][level[s]])) for s in range(S)], keep_raw=True,
                                  inst_names=[n.lower() for n in pkg.mc.instance_names(flat)])
    raws = cpu[4][:S]
    try:
        r = bench.parity_check(raws, t, v, res.npoints, "v(18)")
        r.update(bench.meas_check(raws, ms[1] - ms[0], "v(18)", clauses))
        return r
    finally:
        for f in raws:
            if os.path.exists(f):
                os.remove(f)


def test_bench_parity_hostsim(hostsim_lib):
    r = _draws_vs_reference(hostsim_lib, 2, seed=777)
    assert r["accepted_identical"] and r["bit_identical"] and r["meas_identical"], r


@pytest.mark.gpu
def test_bench_parity_gpu_distinct_draws(cuda_lib):
    """32 distinct (delvto, toxe level) draws in one batch, each against its own reference run"""
    n = 32
    r = _draws_vs_reference(cuda_lib, n, seed=4242)
    assert r["ok"] and r["accepted_identical"] and r["meas_max_rel_err"] <= 1e-9, r
