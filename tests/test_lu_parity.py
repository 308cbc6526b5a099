"""SMPluFac + SMPsolve + NIconvTest(node part) parity against the reference's KLU
(klu_refactor / klu_solve) on recorded matrices: L, U, Udiag, Offx, Rs and the solution."""
import numpy as np
import pytest
from parity_util import GOLDEN, ngt, pkg, relerr, replay_load, trace_calls, first_pattern


def _check(lib, name, tol, S=1):
    flat = ngt.read(f"{GOLDEN}/{name}.flat.ngt")
    trace = ngt.read(f"{GOLDEN}/{name}.trace.ngt.gz")
    pat = first_pattern(trace)
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=pat)
    info = circ.lu_info()
    n = int(pat["n"][0])
    assert info["lnz"] == len(pat["Li"]) and info["unz"] == len(pat["Ui"])
    batch = None
    checked = 0
    for call in trace_calls(trace):
        c = f"c{call}/"
        if c + "lu/Lx" not in trace:
            continue
        batch, ours, ref, maps = replay_load(lib, circ, flat, trace, call, S=S, batch=batch)
        batch.lufac_solve()
        V = batch.get("lu.V", (S, -1))
        Rs = batch.get("lu.Rs", (S, -1))
        x = batch.get("x", (2, circ.neq + 1, S))
        u, l = info["unz"], info["lnz"]
        for s in sorted({0, S - 1}):
            assert relerr(V[s, :u], trace[c + "lu/Ux"], 1e-300).max() <= tol
            assert relerr(V[s, u:u + l], trace[c + "lu/Lx"], 1e-300).max() <= tol
            assert relerr(V[s, u + l:u + l + n], trace[c + "lu/Udiag"], 1e-300).max() <= tol
            assert relerr(V[s, u + l + n:], trace[c + "lu/Offx"], 1e-300).max() <= tol
            assert relerr(Rs[s][pat["Pnum"]], trace[c + "lu/Rs"], 1e-300).max() <= tol
            assert relerr(x[1, 1:, s], trace[c + "sol"][1:circ.neq + 1], 1e-30).max() <= tol * 1e3
        assert (batch.get("lu.singular") < 0).all()
        checked += 1
    assert checked >= 2


def test_lu_hostsim_ro17(hostsim_lib):
    _check(hostsim_lib, "ro17", tol=1e-12)


def test_lu_hostsim_ro101(hostsim_lib):
    _check(hostsim_lib, "ro101", tol=1e-12)


def test_lu_hostsim_bit_exact_when_matrix_is(hostsim_lib):
    """feeding the reference's own Ax/rhs must reproduce KLU's values exactly (same operation
    order, no FMA): this pins the task schedule against klu_refactor.c / klu_solve.c"""
    lib = hostsim_lib
    flat = ngt.read(f"{GOLDEN}/ro17.flat.ngt")
    trace = ngt.read(f"{GOLDEN}/ro17.trace.ngt.gz")
    pat = first_pattern(trace)
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=pat)
    info = circ.lu_info()
    n = int(pat["n"][0]); u, l = info["unz"], info["lnz"]
    b = pkg.Batch(circ, 1)
    for call in trace_calls(trace):
        c = f"c{call}/"
        if c + "lu/Lx" not in trace:
            continue
        b.put("ctl.xsel", np.zeros(1, np.int32))
        b.put("Ax", trace[c + "Ax_fact"])
        x = np.zeros((2, circ.neq + 1, 1)); x[1, :, 0] = trace[c + "rhs"][:circ.neq + 1]; x[1, 0, 0] = 0
        x[0, :, 0] = trace[c + "rhsOld"][:circ.neq + 1]
        b.put("x", x)
        b.lufac_solve()
        V = b.get("lu.V", (1, -1))[0]
        assert np.array_equal(V[:u], trace[c + "lu/Ux"]) and np.array_equal(V[u:u + l], trace[c + "lu/Lx"])
        assert np.array_equal(V[u + l:u + l + n], trace[c + "lu/Udiag"])
        assert np.array_equal(V[u + l + n:], trace[c + "lu/Offx"])
        sol = b.get("x", (2, circ.neq + 1, 1))[1, 1:, 0]
        assert np.array_equal(sol, trace[c + "sol"][1:circ.neq + 1])


@pytest.mark.gpu
@pytest.mark.parametrize("name,S", [("ro17", 1), ("ro17", 37), ("ro101", 1), ("ro101", 3)])
def test_lu_gpu(cuda_lib, name, S):
    _check(cuda_lib, name, tol=1e-9, S=S)
