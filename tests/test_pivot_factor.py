"""The library's own pivoting factor (csrc/ngb_pivot.c, the role klu_factor plays behind SMPreorder) against every
pivoting factor the reference recorded: same symbolic analysis (klu_analyze's P, Q, R) and the matrix values of that
call in, and the row order Pnum, the column patterns of L and U IN THEIR STORED ORDER (it fixes the order of the
subtractions of every later refactor) and the off-diagonal block pattern out -- all integer, all identical."""
import ctypes
import glob
import os
import numpy as np
import pytest
from parity_util import GOLDEN, ngt, pkg

NAMES = sorted(os.path.basename(p)[:-len(".trace.ngt.gz")] for p in glob.glob(f"{GOLDEN}/*.trace.ngt.gz"))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


def _factor(lib, circ, pat, Ax, tol):
    L = lib.L
    n = int(pat["n"][0]); nb = int(pat["nblocks"][0])
    P = np.ascontiguousarray(pat["P"], np.int32); Q = np.ascontiguousarray(pat["Q"], np.int32); R = np.ascontiguousarray(pat["R"], np.int32)
    lib.check(L.ngbCircuitSetSymbolic(circ.h, n, nb, _ip(P), _ip(Q), _ip(R)), "ngbCircuitSetSymbolic")
    Ax = np.ascontiguousarray(Ax, np.float64)
    L.ngbCircuitFactor.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.c_double]
    rc = L.ngbCircuitFactor(circ.h, Ax.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.c_double(tol))
    if rc:
        return rc, None
    out = {"Pnum": np.zeros(n, np.int32), "Lp": np.zeros(n + 1, np.int32), "Up": np.zeros(n + 1, np.int32), "Offp": np.zeros(n + 1, np.int32)}
    lib.check(L.ngbCircuitGetLuPattern(circ.h, _ip(out["Pnum"]), _ip(out["Lp"]), None, _ip(out["Up"]), None, _ip(out["Offp"]), None), "get")
    out["Li"] = np.zeros(max(int(out["Lp"][n]), 1), np.int32); out["Ui"] = np.zeros(max(int(out["Up"][n]), 1), np.int32)
    out["Offi"] = np.zeros(max(int(out["Offp"][n]), 1), np.int32)
    lib.check(L.ngbCircuitGetLuPattern(circ.h, None, None, _ip(out["Li"]), None, _ip(out["Ui"]), None, _ip(out["Offi"])), "get")
    return 0, out


@pytest.mark.parametrize("name", NAMES)
def test_own_pivoting_factor_equals_klu(hostsim_lib, name):
    lib = hostsim_lib
    flat = ngt.read(f"{GOLDEN}/{name}.flat.ngt")
    trace = ngt.read(f"{GOLDEN}/{name}.trace.ngt.gz")
    calls = sorted({int(k.split("/")[0][1:]) for k in trace if k.endswith("/pat/n")})
    if not calls:
        pytest.skip("no pivoting factor recorded")
    circ = pkg.Circuit.from_flat(lib, flat)
    tol = float(flat["opt/pivreltol"][0])
    checked = 0
    for k in calls:
        pre = f"c{k}/"
        if pre + "Ax_fact" not in trace:
            continue
        pat = {kk[len(pre + "pat/"):]: v for kk, v in trace.items() if kk.startswith(pre + "pat/")}
        rc, ours = _factor(lib, circ, pat, trace[pre + "Ax_fact"], tol)
        assert rc == 0, (name, k, lib.last_error() if hasattr(lib, "last_error") else rc)
        n = int(pat["n"][0])
        for key in ("Pnum", "Lp", "Up", "Offp"):
            assert np.array_equal(ours[key], pat[key]), (name, k, key)
        assert np.array_equal(ours["Li"][:pat["Lp"][n]], pat["Li"]), (name, k, "Li")
        assert np.array_equal(ours["Ui"][:pat["Up"][n]], pat["Ui"]), (name, k, "Ui")
        assert np.array_equal(ours["Offi"][:pat["Offp"][n]], pat["Offi"]), (name, k, "Offi")
        checked += 1
    assert checked > 0


def test_singular_matrix_is_reported(hostsim_lib):
    lib = hostsim_lib
    flat = ngt.read(f"{GOLDEN}/inv.flat.ngt"); trace = ngt.read(f"{GOLDEN}/inv.trace.ngt.gz")
    k = sorted({int(kk.split("/")[0][1:]) for kk in trace if kk.endswith("/pat/n")})[0]
    pre = f"c{k}/"
    pat = {kk[len(pre + "pat/"):]: v for kk, v in trace.items() if kk.startswith(pre + "pat/")}
    circ = pkg.Circuit.from_flat(lib, flat)
    rc, _ = _factor(lib, circ, pat, np.zeros_like(trace[pre + "Ax_fact"]), 1e-3)
    assert rc == 102          # E_SINGULAR


def test_own_analysis_gives_a_usable_factor(hostsim_lib):
    """ngbCircuitAnalyze (one block, minimum degree) + ngbCircuitFactor: a complete transient without any imported KLU
    object; the pivot order differs from KLU's, so the waveform agrees to rounding (1e-9 of its range), same step count"""
    lib = hostsim_lib
    flat = ngt.read(f"{GOLDEN}/inv.flat.ngt"); trace = ngt.read(f"{GOLDEN}/inv.trace.ngt.gz"); wave = ngt.read(f"{GOLDEN}/inv.wave.ngt")
    k = sorted({int(kk.split("/")[0][1:]) for kk in trace if kk.endswith("/pat/n")})[0]
    circ = pkg.Circuit.from_flat(lib, flat)
    lib.check(lib.L.ngbCircuitAnalyze(circ.h), "ngbCircuitAnalyze")
    Ax = np.ascontiguousarray(trace[f"c{k}/Ax_fact"], np.float64)
    lib.L.ngbCircuitFactor.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.c_double]
    lib.check(lib.L.ngbCircuitFactor(circ.h, Ax.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.c_double(1e-3)), "ngbCircuitFactor")
    b = pkg.Batch(circ, 1)
    res = b.tran(8192, wave["save_eq"])
    t, v = res.waves()
    n = int(res.npoints[0])
    assert int(res.err[0]) == 0 and n == len(wave["time"])
    rng = np.max(np.abs(wave["values"]), axis=0)
    assert (np.max(np.abs(v[0, :n, :] - wave["values"]), axis=0) / rng <= 1e-9).all()


def test_pivot_events_verified_without_the_host(hostsim_lib, monkeypatch):
    """At the reference's pivoting events the LU launch checks the event's recorded pivot order against KLU's rule on the
    sample's own matrix (NgbLuSched.vchk: lpivot, klu_kernel.c:370-470) and only a sample that fails goes to the host's
    pivoting factor.  The recorded run itself never does; eight sweep points of the mixed cell need it once per point
    instead of four times -- same iteration counts, same bits as with every event factored on the host"""
    import hashlib
    from parity_util import run_patterns
    flat = ngt.read(f"{GOLDEN}/mix.flat.ngt"); trace = ngt.read(f"{GOLDEN}/mix.trace.ngt.gz"); wave = ngt.read(f"{GOLDEN}/mix.wave.ngt")
    pts = [("2.0", "1k"), ("1.6", "200"), ("1.6", "5k"), ("2.4", "200"), ("2.4", "5k"), ("1.69098", "1348.2"), ("1.69412", "2364.7"), ("1.69412", "2383.5")]

    def run(points):
        circ = pkg.Circuit.from_flat(hostsim_lib, flat, lu_pattern=run_patterns(trace))
        b = pkg.Batch(circ, len(points))
        pkg.sweep.apply(b, flat, dc={"vdd": [p[0] for p in points]}, res={"r1": [p[1] for p in points]})
        res = b.tran(8192, wave["save_eq"])
        return res, hashlib.md5(res.waves()[1].tobytes()).hexdigest()

    res, _ = run(pts[:1])
    assert res.repivots == 0 and int(res.numiter[0]) == int(wave["stats"][2])
    dev, h_dev = run(pts)
    monkeypatch.setenv("NGB_HOST_PIVOT", "1")
    host, h_host = run(pts)
    assert h_dev == h_host and np.array_equal(dev.numiter, host.numiter)
    assert 0 < dev.repivots < host.repivots and host.repivots >= 4 * len(pts)
