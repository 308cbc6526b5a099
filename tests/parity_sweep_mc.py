"""Wide parity sweep of BASELINE config 3 (not a pytest module; TEST INFRASTRUCTURE: it executes oracle/_ref/ngspice).

   python tests/parity_sweep_mc.py LO HI [hostsim|cuda]

runs draws LO..HI-1 of bench.py's rank-0 Monte-Carlo batch (per-instance delvto, per-sample continuous toxe, laid out by
toxe exactly like the bench) through the library (host build of the kernel bodies by default) as ONE batch and through the
stock reference, one process per draw, and compares every accepted point of v(out) bit for bit.
Round 2: draws 0..4095 as one batch on the B200 (`0 4096 cuda`, 273 s): all 4 096 bit-identical; draws 16..255 on the host build
likewise; after the overlay reads of the per-sample rows (end of round 2): draws 2000..2031 on the host build and draws 100..1155 on the B200 (three calls, 1 056 draws), all bit-identical;
bench.py itself checks draws 0..cores-1 in every run."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench                                                  # noqa: E402  (cpu_reference_run, read_rawfile)
from parity_util import GOLDEN, HOSTSIM, ngt, pkg, run_patterns   # noqa: E402


def main():
    lo, hi = int(sys.argv[1]), int(sys.argv[2])
    which = sys.argv[3] if len(sys.argv) > 3 else "hostsim"
    lib = pkg.Library(HOSTSIM) if which == "hostsim" else pkg.Library()
    S, rank = 4096, 0
    flat = ngt.read(f"{GOLDEN}/ro17k.flat.ngt"); flat["tran/tstop"] = np.array([pkg.mc.spice_number("150ns")])
    trace = ngt.read(f"{GOLDEN}/ro17k.trace.ngt.gz"); wave = ngt.read(f"{GOLDEN}/ro17k.wave.ngt")
    ninst = int(flat["b4/ninst"][0])
    dv_raw = pkg.mc.draw_delvto(S, ninst, sigma=0.015, seed=1000 + rank)
    tox_raw = 1.4e-9 * (1.0 + 0.03 * np.random.default_rng(5000 + rank).normal(size=S))
    order = np.argsort(tox_raw, kind="stable"); tox_raw, dv_raw = tox_raw[order], dv_raw[order]
    dv = pkg.mc.delvto_as_parsed(dv_raw)
    tox = np.array([pkg.mc.spice_number(f"{x:.17g}") for x in tox_raw[lo:hi]])
    b4t = ngt.read(f"{GOLDEN}/b4temp.tables.ngt.gz")
    raw = {"model": b4t["ro17k/b4t/model"], "inst": b4t["ro17k/b4t/inst"], "inst_model": b4t["ro17k/b4t/inst_model"],
           "temp": b4t["ro17k/b4t/temp"][0, 0], "vt0": b4t["ro17k/opt/vt0"][0]}
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=run_patterns(trace))
    inst_host, prow_t, mtab, ptab = pkg.mc.bsim4_with_toxe(lib, raw, tox, dv[lo:hi])
    b = pkg.Batch(circ, hi - lo)
    b.put("b4.inst", inst_host); b.set_bsim4_rows(prow_t, mtab, ptab)
    res = b.tran(6144, wave["save_eq"][:1])
    t, v = res.waves()
    draws = [(dv_raw[p], float(tox_raw[p])) for p in range(lo, hi)]
    nproc = min(os.cpu_count() or 1, hi - lo)
    per = (hi - lo + nproc - 1) // nproc
    draws += [draws[-1]] * (nproc * per - len(draws))
    cpu = bench.cpu_reference_run("mc_ro17", nproc, per, draws=draws, keep_raw=True, inst_names=[n.lower() for n in pkg.mc.instance_names(flat)])
    bad = []
    for i in range(hi - lo):
        tt, vv = bench.read_rawfile(cpu[4][i], ["v(18)"])
        n = len(tt)
        if not (int(res.npoints[i]) == n and np.array_equal(t[i][:n], tt) and np.array_equal(v[i][:n, 0], vv["v(18)"])):
            bad.append(lo + i)
    for f in cpu[4]:
        try:
            os.remove(f)
        except OSError:
            pass
    print(f"draws {lo}..{hi - 1} on {which}: {hi - lo - len(bad)} of {hi - lo} bit-identical to the reference; not identical: {bad}; host re-pivots {res.repivots}")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
