"""Error behaviour of the C ABI: what this path does not implement is refused loudly, with the reference's own error
number where it has one (sperror.h: E_SINGULAR 102, E_ITERLIM 103, E_ORDER 104, E_METHOD 105, E_TIMESTEP 106) or
E_UNSUPP (10), when the circuit is described -- never at run time, never by a silent fallback.
Host build of the library; the checks are host logic (ngb_host.c), the same object code the product links."""
import copy
import numpy as np
import pytest
from parity_util import GOLDEN, ngt, pkg, run_patterns


def _flat(name):
    return copy.deepcopy(ngt.read(f"{GOLDEN}/{name}.flat.ngt"))


def _fails(lib, flat, code, text):
    with pytest.raises(pkg.NgbError) as e:
        pkg.Circuit.from_flat(lib, flat)
    assert f"code {code}:" in str(e.value) and text in str(e.value), str(e.value)


def test_unsupported_source_waveforms_refused(hostsim_lib):
    """TRNOISE (7) / TRRANDOM (8) / EXTERNAL sources keep host-side state (vsrcload.c:369-439): E_UNSUPP"""
    for key in ("vsrc/fn", "isrc/fn"):
        flat = _flat("srcs")
        flat[key][0, 0] = 7
        _fails(hostsim_lib, flat, 10, "waveform type 7")


def test_pwl_source_needs_its_corner_list(hostsim_lib):
    flat = _flat("srcs")
    for k in ("isrc/pwl_ptr", "isrc/pwl"):
        del flat[k]
    _fails(hostsim_lib, flat, 1, "has no corner list")


def test_unknown_method_and_high_order_refused(hostsim_lib):
    """integration methods other than TRAPEZOIDAL / GEAR do not exist (niinteg.c:76); orders above 2 are never taken by
    DCtran (dctran.c:794-826) and maxord > 2 would only lengthen the state ring"""
    flat = _flat("inv")
    flat["opt/method"][...] = 3
    _fails(hostsim_lib, flat, 105, "method")
    flat = _flat("inv")
    flat["opt/maxorder"][...] = 3
    _fails(hostsim_lib, flat, 104, "maxord")


def test_bsim4_nqs_refused(hostsim_lib):
    """trnqsMod / acnqsMod add the charge-deficit node and its stamps (b4ld.c:4650-4716): E_UNSUPP"""
    flat = _flat("inv")
    flat["b4/flags"][0] |= 0x100
    _fails(hostsim_lib, flat, 10, "trnqsMod")


def test_bypass_refused(hostsim_lib):
    flat = _flat("inv")
    flat["opt/bypass"][...] = 1
    with pytest.raises(pkg.NgbError):
        pkg.Circuit.from_flat(hostsim_lib, flat)


def test_transient_needs_a_pivot_order(hostsim_lib):
    """klu_analyze / klu_factor stay in the reference: without their result there is no schedule"""
    circ = pkg.Circuit.from_flat(hostsim_lib, _flat("inv"))
    b = pkg.Batch(circ, 1)
    with pytest.raises(pkg.NgbError) as e:
        b.tran(16, np.array([1], np.int32))
    assert "no LU pattern" in str(e.value)


def test_too_few_output_points_is_not_an_overrun(hostsim_lib):
    """max_points smaller than the run: the waveform buffers are not written past their end and the counts go on"""
    flat = _flat("inv")
    trace = ngt.read(f"{GOLDEN}/inv.trace.ngt.gz")
    wave = ngt.read(f"{GOLDEN}/inv.wave.ngt")
    circ = pkg.Circuit.from_flat(hostsim_lib, flat, lu_pattern=run_patterns(trace))
    b = pkg.Batch(circ, 2)
    res = b.tran(100, wave["save_eq"])
    t, v = res.waves()
    assert t.shape == (2, 100)
    for s in range(2):
        assert int(res.accepted[s]) == int(wave["stats"][0]) and int(res.npoints[s]) == len(wave["time"])
        assert np.array_equal(t[s], wave["time"][:100]) and np.array_equal(v[s], wave["values"][:100])


def test_diode_thermal_flag_needs_its_node(hostsim_lib):
    """the thermal node exists exactly for the instances with self-heating, the qp node for those with soft recovery
    (diosetup.c:419-430): flags and node table of ngbCircuitAddDiodes must say the same"""
    flat = _flat("diosh")
    flat["dio/nodes"][2, :] = 0              # the dump's row 2 is DIOtempNode
    _fails(hostsim_lib, flat, 1, "thermal / qp nodes")
    flat = _flat("diorr")
    flat["dio/nodes"][5, :] = 0              # DIOqpNode
    _fails(hostsim_lib, flat, 1, "thermal / qp nodes")
