"""Drop-in boundary (INTEGRATION.md level 1): the UNMODIFIED reference objects, linked with
integration/ngb_shim.c in front of CKTload / SMPreorder / SMPluFac / SMPsolve, must write the same
rawfile as the stock reference binary -- every node and branch at every accepted time point.

CPU: oracle/_ref/ngspice_ngb_hostsim (the shim bound to the host build of the kernels).
GPU: oracle/_ref/ngspice_ngb (the shim bound to libngb200.so).
Both binaries are produced by oracle/build_ref.sh; the tests skip when they are absent."""
import os
import subprocess
import tempfile
import numpy as np
import pytest
from parity_util import GOLDEN, ROOT

REF = os.path.join(ROOT, "oracle", "_ref", "ngspice")
NETLISTS = ["ro17k", "inv", "dio", "b3ring", "arr", "latch", "srcs", "invsrc", "invgmin", "invshunt", "diosh", "diorr", "dioshrr"]     # invsrc / invgmin: CKTop's own source / gmin stepping drives CKTsrcFact and CKTdiagGmin through the shim


def _payload(path):
    data = open(path, "rb").read()
    i = data.index(b"Binary:\n")
    head = data[:i].decode(errors="replace")
    keep = "\n".join(ln for ln in head.splitlines() if not ln.startswith("Date:"))
    return keep, np.frombuffer(data[i + 8:], dtype=np.float64)


def _run(exe, name, tmp, env=None):
    raw = os.path.join(tmp, f"{os.path.basename(exe)}_{name}.raw")       # never /dev/null: ngspice unlinks its -r target
    cir = os.path.join(GOLDEN, "netlists", name + ".cir")
    p = subprocess.run([exe, "-b", "-r", raw, cir], capture_output=True, text=True, env=dict(os.environ, **(env or {})), timeout=600)
    assert os.path.exists(raw), p.stdout[-2000:] + p.stderr[-2000:]
    return _payload(raw), p.stdout + p.stderr


def _check(exe, name, expect_backend, env=None, expect="ngb_shim: CKTload"):
    if not (os.path.exists(REF) and os.path.exists(exe)):
        pytest.skip("oracle/_ref binaries not built (oracle/build_ref.sh needs /root/reference)")
    with tempfile.TemporaryDirectory(prefix="ngb_shim_") as tmp:
        (h0, v0), _ = _run(REF, name, tmp)
        (h1, v1), log = _run(exe, name, tmp, env)
    assert expect in log and expect_backend in log, log[-1500:]     # the shim really took the hot path
    assert h0 == h1
    assert v0.shape == v1.shape
    _check.nvars = int([ln for ln in h0.splitlines() if ln.startswith("No. Variables")][0].split(":")[1])
    _check.names = [ln.split()[1] for ln in h0.splitlines() if ln.startswith("\t") and len(ln.split()) >= 3 and ln.split()[0].isdigit()]
    return v0, v1


@pytest.mark.parametrize("name", NETLISTS)
def test_dropin_hostsim_rawfile_identical(name):
    v0, v1 = _check(os.path.join(ROOT, "oracle", "_ref", "ngspice_ngb_hostsim"), name, "hostsim")
    assert np.array_equal(v0, v1)


def _vbic_close(v0, v1):
    """VBIC Jacobians come from dual numbers (rounding-level differences from the generated code):
    same number of points (v*.shape checked by _check); per point |v - v_ref| / max(|v_ref|, 1e-6) (SURVEY.md section 8(d),
    vntol as the floor) for every circuit node and branch current.  Internal device nodes (`q1#substrate`: a floating
    substrate held at 1e-10 V by gmin alone) are compared against their vector's range instead: 1e-9 * 1e-6 V is far below
    what a node defined by a 1e-12 S conductance can reproduce under a different summation order"""
    a = v0.reshape(-1, _check.nvars); b = v1.reshape(-1, _check.nvars)
    internal = np.array(["#" in n and not n.endswith("#branch") for n in _check.names])
    floor = np.where(internal, np.maximum(np.max(np.abs(a), axis=0), 1e-6), 1e-6)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(a), floor[None, :])))


@pytest.mark.parametrize("name", ["vbic", "mix", "vbicsh", "vbicxf", "vbicshxf"])
def test_dropin_hostsim_vbic_rawfile(name):
    v0, v1 = _check(os.path.join(ROOT, "oracle", "_ref", "ngspice_ngb_hostsim"), name, "hostsim")
    # per point; the mixed cell holds a ring oscillator whose edges carry the VBIC rounding difference through zero crossings (5.7e-9)
    assert _vbic_close(v0, v1) <= (1e-8 if name == "mix" else 1e-9)


def test_dropin_hostsim_load_only_identical():
    """NGB_SHIM_LU=0: device load, host KLU"""
    v0, v1 = _check(os.path.join(ROOT, "oracle", "_ref", "ngspice_ngb_hostsim"), "inv", "hostsim", env={"NGB_SHIM_LU": "0"})
    assert np.array_equal(v0, v1)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["vbic", "mix", "vbicshxf"])
def test_dropin_gpu_vbic_rawfile(name):
    v0, v1 = _check(os.path.join(ROOT, "oracle", "_ref", "ngspice_ngb"), name, "cuda-sm_100a")
    assert _vbic_close(v0, v1) <= (1e-8 if name == "mix" else 1e-9)      # see test_dropin_hostsim_vbic_rawfile


@pytest.mark.gpu
@pytest.mark.parametrize("name", NETLISTS)
def test_dropin_gpu_rawfile(name):
    v0, v1 = _check(os.path.join(ROOT, "oracle", "_ref", "ngspice_ngb"), name, "cuda-sm_100a")
    if name in ("dio", "srcs"):   # SIN / SFFM / AM sources: CUDA's sin() is not glibc's; the north_star tolerance applies
        assert np.max(np.abs(v0 - v1) / np.maximum(np.abs(v0), 1e-6)) <= 1e-9
    elif name in ("invsrc", "invgmin", "invshunt", "diosh", "diorr", "dioshrr"):
        # added after the round's last GPU run: held to the north_star bar on the device (identical point count, 1e-9);
        # the bit-identity of these routes is shown on the host build above
        assert np.max(np.abs(v0 - v1) / np.maximum(np.abs(v0), 1e-6)) <= 1e-9
    else:
        assert np.array_equal(v0, v1)


# ---- table mode: the library installed through DEVices[t]->DEVload (devdefs.h:57, cktload.c:70-91) -------------------
TABLE_NETLISTS = ["inv", "ro17k", "latch", "dio", "b3ring", "invgmin"]


def _table_identical(exe, backend, name):
    """every device type replaced: the reference's own CKTload (clear, DEVices loop, nodesets) around one ngbLoad whose assembled
    contributions are added to the cleared matrix -- the same bits as the stock binary"""
    v0, v1 = _check(exe, name, backend, env={"NGB_SHIM_TABLE": "1"}, expect="SPICEdev table")
    assert np.array_equal(v0, v1)


def _table_mixed(exe, backend):
    """mixcpu: BSIM4 inverter + diode + R/C/V on the library, BJT / inductor / VCCS on their own CPU DEVload, all stamping the
    same KLU matrix (host factorisation).  Sums over device types are taken in another order than the reference's, so the
    bar is north_star's: identical accepted points, 1e-9 per vector"""
    v0, v1 = _check(exe, "mixcpu", backend, expect="SPICEdev table")
    assert _vbic_close(v0, v1) <= 1e-9


@pytest.mark.parametrize("name", TABLE_NETLISTS)
def test_dropin_table_hostsim_identical(name):
    _table_identical(os.path.join(ROOT, "oracle", "_ref", "ngspice_ngb_hostsim"), "hostsim", name)


def test_dropin_table_hostsim_mixed_device_set():
    _table_mixed(os.path.join(ROOT, "oracle", "_ref", "ngspice_ngb_hostsim"), "hostsim")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["inv", "ro17k", "latch"])
def test_dropin_table_gpu_identical(name):
    _table_identical(os.path.join(ROOT, "oracle", "_ref", "ngspice_ngb"), "cuda-sm_100a", name)


@pytest.mark.gpu
def test_dropin_table_gpu_mixed_device_set():
    _table_mixed(os.path.join(ROOT, "oracle", "_ref", "ngspice_ngb"), "cuda-sm_100a")
