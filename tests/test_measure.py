"""Device-side `.meas tran` (the output path, src/frontend/outitf.c:633 + com_measure2.c:378-663): the
measurement clauses are evaluated while the accepted points are produced, so a Monte-Carlo run does not have to keep
or copy its waveforms.  Checked (a) against a numpy restatement of com_measure_when over the stored waveform of the
same run -- bit-identical -- and (b) against the stock reference binary's `.meas` results for the same netlist
(tests/golden/ro17k.meas.json, written by make_golden.py ro17kmeas; the reference keeps 7 digits of a measurement)."""
import json
import numpy as np
import pytest
from parity_util import run_patterns, GOLDEN, ngt, pkg


def measure_when(t, v, kind, count, val, td):
    """com_measure_when for one real vector against a constant (com_measure2.c:455-660)"""
    first = 0; section = -1; rise = fall = 0
    pv = pt = 0.0
    for scale, value in zip(t, v):
        if scale < td:
            continue
        if first == 1:
            rise = fall = 0
            if value < val:
                section = 0
                if pv >= val:
                    fall = 1
            else:
                section = 1
                if pv < val:
                    rise = 1
        if first > 1:
            if section == 0 and value >= val:
                section = 1; rise += 1
            elif section == 1 and value <= val:
                section = 0; fall += 1
            have = rise if kind == 0 else (fall if kind == 1 else rise + fall)
            if have == count:
                return pt + (val - pv) * (scale - pt) / (value - pv)
        first += 1
        pv, pt = value, scale
    return float("nan")


def _run(lib, S, max_points):
    flat = ngt.read(f"{GOLDEN}/ro17k.flat.ngt")
    trace = ngt.read(f"{GOLDEN}/ro17k.trace.ngt.gz")
    gold = json.load(open(f"{GOLDEN}/ro17k.meas.json"))
    nn = bytes(flat["node/names_bytes"].astype(np.uint8)).decode().split("\n")
    eq_of = {ln.split(" ", 1)[1].lower(): int(ln.split(" ", 1)[0]) for ln in nn if ln.strip()}
    clauses, owner = [], []
    for nm, g in gold.items():
        for c in g["clauses"]:
            clauses.append((eq_of[c["node"]], c["kind"], c["count"], c["val"], c["td"]))
            owner.append(nm)
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=run_patterns(trace))
    b = pkg.Batch(circ, S)
    b.set_measures(clauses)
    save = sorted({c[0] for c in clauses})
    res = b.tran(max_points, save if max_points else [])
    return gold, clauses, owner, save, res


def _check(lib, S, exact):
    gold, clauses, owner, save, res = _run(lib, S, 4096)
    ms = res.measures()
    t, v = res.waves()
    for s in range(S):
        n = int(res.npoints[s])
        for k, c in enumerate(clauses):
            want = measure_when(t[s, :n], v[s, :n, save.index(c[0])], c[1], c[2], c[3], c[4])
            assert (np.isnan(want) and np.isnan(ms[k, s])) or want == ms[k, s], (k, want, ms[k, s])
        # the reference's own results
        for nm, g in gold.items():
            ks = [k for k, o in enumerate(owner) if o == nm]
            got = ms[ks[-1], s] - ms[ks[0], s] if len(ks) == 2 else ms[ks[0], s]
            if g["value"] is None:
                assert np.isnan(got)
            else:
                assert abs(got - g["value"]) <= (1e-6 if exact else 2e-6) * abs(g["value"]), (nm, got, g["value"])
    return ms


def test_measure_hostsim(hostsim_lib):
    _check(hostsim_lib, 2, True)


def test_measure_without_waveforms_hostsim(hostsim_lib):
    """max_points = 0, nothing saved: the measurements are all that leaves the run"""
    a = _check(hostsim_lib, 1, True)
    gold, clauses, owner, save, res = _run(hostsim_lib, 1, 0)
    b = res.measures()
    assert np.array_equal(a, b, equal_nan=True) and int(res.err[0]) == 0 and int(res.accepted[0]) > 100


@pytest.mark.gpu
def test_measure_device(cuda_lib):
    _check(cuda_lib, 64, False)
    gold, clauses, owner, save, res = _run(cuda_lib, 64, 0)
    ms = res.measures()
    assert (res.err == 0).all() and np.array_equal(ms[:, :1].repeat(64, 1), ms, equal_nan=True)


def test_spice_number_follows_the_reference_parser():
    """pkg.mc.spice_number restates INPevaluate (inpeval.c:65-201): digits accumulate in a double as `10 * mantis + c - '0'`
    -- the character code is added before '0' is subtracted, which matters once the mantissa passes 2^53 -- and the result is
    mantis * pow(10, exponent), not the nearest double.  Values recorded from the reference's own parse of `delvto=<text>`
    (oracle/_ref/ngspice_dump, b4t/inst) for two 17-digit tokens the plain 10 m + digit recurrence gets wrong by one ulp"""
    from parity_util import pkg
    assert pkg.mc.spice_number("0.0095290870168629038") == 0.009529087016862904
    assert pkg.mc.spice_number("0.0087862920898565608") == 0.00878629208985656
    assert pkg.mc.spice_number("1.2905820754597296e-09") == 1.2905820754597296e-09
    assert pkg.mc.spice_number("150ns") == 150 * 1e-9 and pkg.mc.spice_number("2") == 2.0
