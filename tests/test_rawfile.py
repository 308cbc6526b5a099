"""Batch-aware binary rawfile (the output path): ngbTranWriteRaw writes the stored waveforms of a batch
as consecutive `Transient Analysis` plots of one file.  Checked (a) against the header `ngspice -b -r` writes for the same
fixture (src/frontend/outitf.c:881-923, :997-1029; recorded from oracle/_ref/ngspice on tests/golden/netlists/inv.cir: the
lines below, with the reference's own Date / Command lines and its 14 saved vectors replaced by the four the fixture stores) and
(b) by reading the file back with the reader bench.py uses for the reference's rawfiles: every plot returns the bits of
ngbTranWaves, and sample 0 the reference's waveform of the fixture."""
import os
import re
import numpy as np
import pytest
from parity_util import run_patterns, GOLDEN, ngt, pkg, relerr

NAMES = ["v(out)", "v(in)", "i(vdd)", "i(vin)"]              # the fixture's saved equations 3, 2, 13, 12 under the names the reference gives them (outitf.c:1014-1020)


def read_plots(path):
    """every plot of a binary rawfile: [(header dict, variable names, array [points][variables])]"""
    raw = open(path, "rb").read()
    plots, pos = [], 0
    while pos < len(raw):
        end = raw.index(b"Binary:\n", pos) + 8
        lines = raw[pos:end].decode().split("\n")
        hdr = {ln.split(":", 1)[0]: ln.split(":", 1)[1].strip() for ln in lines if ":" in ln and not ln.startswith("\t")}
        names = [ln.split("\t")[2] for ln in lines if ln.startswith("\t")]
        nv, npnt = int(hdr["No. Variables"]), int(hdr["No. Points"])
        assert nv == len(names)
        data = np.frombuffer(raw, dtype="<f8", count=nv * npnt, offset=end).reshape(npnt, nv)
        plots.append((hdr, names, data, lines))
        pos = end + 8 * nv * npnt
    return plots


def _check(lib, tmp_path, S, exact):
    flat = ngt.read(f"{GOLDEN}/inv.flat.ngt"); trace = ngt.read(f"{GOLDEN}/inv.trace.ngt.gz"); wave = ngt.read(f"{GOLDEN}/inv.wave.ngt")
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=run_patterns(trace))
    b = pkg.Batch(circ, S)
    save = [int(e) for e in wave["save_eq"]]
    res = b.tran(2048, save)
    assert (res.err == 0).all()
    t, v = res.waves()
    path = str(tmp_path / "batch.raw")
    res.write_raw(path, NAMES, title="* cmos inverter on the tests/bsim4 qa model cards", date="today")
    plots = read_plots(path)
    assert len(plots) == S
    for s, (hdr, names, data, lines) in enumerate(plots):
        n = int(res.npoints[s])
        # the header, line for line as fileInit / fileInit_pass2 write it
        assert lines[0] == "Title: * cmos inverter on the tests/bsim4 qa model cards" and lines[1] == "Date: today"
        assert lines[2].startswith("Command: ") and lines[3] == "Plotname: Transient Analysis" and lines[4] == "Flags: real"
        assert lines[5] == "No. Variables: 5" and lines[6] == "No. Points: %-8d" % n and lines[7] == "Variables:"
        assert lines[8] == "\t0\ttime\ttime"
        assert lines[9:13] == ["\t1\tv(out)\tvoltage", "\t2\tv(in)\tvoltage", "\t3\ti(vdd)\tcurrent", "\t4\ti(vin)\tcurrent"]
        assert lines[13] == "Binary:"
        assert data.shape == (n, 5)
        assert np.array_equal(data[:, 0], t[s, :n]) and np.array_equal(data[:, 1:], v[s, :n, :])
    # sample 0 is the fixture itself: the reference's accepted points
    d0 = plots[0][2]
    assert d0.shape[0] == wave["time"].shape[0]
    if exact:
        assert np.array_equal(d0[:, 0], wave["time"]) and np.array_equal(d0[:, 1:], wave["values"])
    else:
        assert relerr(d0[:, 0], wave["time"]).max() < 1e-9 and relerr(d0[:, 1], wave["values"][:, 0], floor=1e-6).max() < 1e-9
    # a slice of the batch, and the default date in the reference's datestring format (misc_time.c:59-76)
    if S > 1:
        res.write_raw(path, NAMES, first_sample=1, nsamples=1)
        (hdr, names, data, lines), = read_plots(path)
        assert np.array_equal(data[:, 0], t[1, :int(res.npoints[1])])
        assert re.fullmatch(r"[A-Z][a-z]{2} [A-Z][a-z]{2} [ \d]\d \d\d:\d\d:\d\d  \d{4}", hdr["Date"]), hdr["Date"]


def test_rawfile_hostsim(hostsim_lib, tmp_path):
    _check(hostsim_lib, tmp_path, 2, True)


def test_rawfile_needs_stored_waveforms(hostsim_lib, tmp_path):
    flat = ngt.read(f"{GOLDEN}/inv.flat.ngt"); trace = ngt.read(f"{GOLDEN}/inv.trace.ngt.gz")
    b = pkg.Batch(pkg.Circuit.from_flat(hostsim_lib, flat, lu_pattern=run_patterns(trace)), 1)
    b.set_measures([(1, 2, 1, 0.5, 0.0)])
    res = b.tran(0, [])
    with pytest.raises(pkg.NgbError):
        res.write_raw(str(tmp_path / "x.raw"), [])
    b.set_measures([])
    res = b.tran(2048, [1])
    with pytest.raises(pkg.NgbError):
        res.write_raw(str(tmp_path / "nodir" / "x.raw"), ["v(1)"])
    with pytest.raises(pkg.NgbError):
        res.write_raw(str(tmp_path / "x.raw"), ["v(1)"], first_sample=1, nsamples=1)


def test_rawfile_is_read_by_the_reference(hostsim_lib, tmp_path):
    """the reference's own `load` (frontend/rawfile.c raw_read) takes the file as plots tran1 .. tranS"""
    import subprocess
    from parity_util import ROOT
    ref = os.path.join(ROOT, "oracle", "_ref", "ngspice")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/ngspice not built")
    flat = ngt.read(f"{GOLDEN}/inv.flat.ngt"); trace = ngt.read(f"{GOLDEN}/inv.trace.ngt.gz"); wave = ngt.read(f"{GOLDEN}/inv.wave.ngt")
    b = pkg.Batch(pkg.Circuit.from_flat(hostsim_lib, flat, lu_pattern=run_patterns(trace)), 3)
    res = b.tran(2048, [int(e) for e in wave["save_eq"]])
    raw = str(tmp_path / "batch.raw")
    res.write_raw(raw, NAMES, title="batch of three")
    deck = tmp_path / "load.sp"
    deck.write_text(f"* load\n.control\nload {raw}\nprint length(tran2.time) tran3.time[1065] tran1.v(out)[500]\n.endc\n.end\n")
    out = subprocess.run([ref, "-b", str(deck)], capture_output=True, text=True, timeout=120).stdout
    t, v = res.waves()
    assert "length(tran2.time) = 1.066000e+03" in out, out
    assert "tran3.time[1065] = %e" % t[2, 1065] in out and "tran1.v(out)[500] = %e" % v[0, 500, 0] in out, out


@pytest.mark.gpu
def test_rawfile_device(cuda_lib, tmp_path):
    _check(cuda_lib, tmp_path, 8, False)
