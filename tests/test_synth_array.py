"""The synthetic inverter-array generator (BASELINE config 4, ngspice-sf-mirror_b200/synth.py) against
the reference's own CKTsetup of the same netlist: node numbering differs, so Ax and rhs of one load
are compared entry by entry through the node NAMES."""
import importlib
import numpy as np
import pytest
from parity_util import GOLDEN, ngt, pkg

synth = importlib.import_module("ngspice-sf-mirror_b200.synth")
MODE_TRANOP_FLOAT = 0x20 | 0x100          # MODETRANOP | MODEINITFLOAT


def _load_by_name(lib, flat, names, xval, seed_state=None):
    circ = pkg.Circuit.from_flat(lib, flat)
    b = pkg.Batch(circ, 1)
    neq1 = circ.neq + 1
    b.put("ctl.mode", np.array([MODE_TRANOP_FLOAT], np.int32))
    b.put("ctl.active", np.array([1], np.int32))
    b.put("ctl.order", np.array([1], np.int32))
    b.put("ctl.gmin", np.array([1e-12]))
    b.put("ctl.srcfact", np.array([1.0]))
    x = np.zeros((2, neq1, 1))
    for k in range(1, len(names)):
        x[0, k, 0] = xval[names[k]]
    b.put("x", x)
    b.load()
    pat = circ.pattern()
    assert pat["n"] == len(names) - 1                # no structurally empty column: column k is equation k+1
    Ax = b.get("Ax", (1, -1))[0]
    rhs = b.get("x", (2, neq1, 1))[1, :, 0]
    A = {}
    for col in range(pat["n"]):
        for p in range(pat["Ap"][col], pat["Ap"][col + 1]):
            A[(names[pat["Ai"][p] + 1], names[col + 1])] = Ax[p]
    return A, {names[k]: rhs[k] for k in range(1, len(names))}


def test_synthetic_array_matches_reference_setup(hostsim_lib):
    ref = ngt.read(f"{GOLDEN}/arr.flat.ngt")
    base = ngt.read(f"{GOLDEN}/ro17k.flat.ngt")
    ours = synth.inverter_array(base, 4, 4)
    ref_names = bytes(ref["node/names_bytes"].astype(np.uint8)).decode().split("\n")
    ref_names = [ln.split(" ", 1)[1].lower() if " " in ln else "" for ln in ref_names if ln.strip()]
    our_names = ours.pop("node/names")
    assert sorted(ref_names[1:]) == sorted(our_names[1:])
    rng = np.random.default_rng(3)
    xval = {n: float(rng.uniform(0.0, 2.0)) for n in our_names[1:]}
    A1, r1 = _load_by_name(hostsim_lib, ref, ref_names, xval)
    A2, r2 = _load_by_name(hostsim_lib, ours, our_names, xval)
    assert A1.keys() == A2.keys()
    # instance order differs between the two flattenings, so sums may round differently in the last place
    for k in A1:
        assert abs(A1[k] - A2[k]) <= 1e-12 * max(abs(A1[k]), 1e-30), k
    for k in r1:
        assert abs(r1[k] - r2[k]) <= 1e-12 * max(abs(r1[k]), 1e-18), k


def test_synthetic_array_scales(hostsim_lib):
    base = ngt.read(f"{GOLDEN}/ro17k.flat.ngt")
    fl = synth.inverter_array(base, 30, 20)
    circ = pkg.Circuit.from_flat(hostsim_lib, fl)
    pat = circ.pattern()
    assert pat["n"] == 10 * 600 + 4 and int(fl["b4/ninst"][0]) == 1200


@pytest.mark.gpu
def test_long_rail_assembly_tree_gpu(cuda_lib, hostsim_lib):
    """a 48 x 48 array: the supply rail's matrix entries and right-hand side collect 2 304+ contributions per transistor
    type -- more than NGB_ASM_LONG = 4 096 in total --, summed on the device by the chunk tree (ngb_k_assemble_long1 / 2).
    Against the host build's strictly sequential sums: every Ax / rhs entry within 1e-12 relative (other summation order
    in those few rows, identical everywhere else)"""
    base = ngt.read(f"{GOLDEN}/ro17k.flat.ngt")
    fl = synth.inverter_array(base, 48, 48)
    names = fl.pop("node/names")
    rng = np.random.default_rng(5)
    xval = {n: float(rng.uniform(0.0, 2.0)) for n in names[1:]}
    A1, r1 = _load_by_name(hostsim_lib, fl, names, xval)
    A2, r2 = _load_by_name(cuda_lib, fl, names, xval)
    assert A1.keys() == A2.keys()
    same = 0
    for k in A1:
        assert abs(A1[k] - A2[k]) <= 1e-12 * max(abs(A1[k]), 1e-30), (k, A1[k], A2[k])
        same += A1[k] == A2[k]
    for k in r1:
        assert abs(r1[k] - r2[k]) <= 1e-12 * max(abs(r1[k]), 1e-18), k
    assert same >= len(A1) - 64          # only the rail rows may differ in the last place


def _arr16(lib):
    from parity_util import run_patterns
    flat = ngt.read(f"{GOLDEN}/arr16.flat.ngt.gz"); pats = ngt.read(f"{GOLDEN}/arr16.pat.ngt.gz"); wave = ngt.read(f"{GOLDEN}/arr16.wave.ngt")
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=run_patterns(pats))
    b = pkg.Batch(circ, 1)
    res = b.tran(1024, wave["save_eq"])
    t, v = res.waves()
    n = int(res.npoints[0])
    acc, rej, nit = (int(x) for x in wave["stats"][:3])
    assert int(res.err[0]) == 0 and n == len(wave["time"])
    assert (int(res.accepted[0]), int(res.rejected[0]), int(res.numiter[0])) == (acc, rej, nit)
    return circ, t[0, :n], v[0, :n], wave


def test_array16_dc_op_and_transient_hostsim(hostsim_lib):
    """BASELINE config 4 at 16 x 16 (512 BSIM4, 2 564 unknowns, 20 188 LU values): DC operating point (84 iterations) and
    `.tran 10p 1n` through the library's own NIiter / DCtran with KLU's pivot orders -- bit-identical to the reference"""
    circ, t, v, wave = _arr16(hostsim_lib)
    assert circ.lu_info()["nV"] > 20000
    assert np.array_equal(t, wave["time"]) and np.array_equal(v, wave["values"])


@pytest.mark.gpu
def test_array16_dc_op_and_transient_gpu(cuda_lib):
    """the same on the device: this matrix does not fit one CTA's shared memory, so refactor + solve run in the grid-wide LU
    kernel (ngb_k_lu_grid: values in global memory, a grid barrier per level).  Identical step / iteration counts, 1e-9 per
    point"""
    circ, t, v, wave = _arr16(cuda_lib)
    assert np.max(np.abs(t - wave["time"]) / np.maximum(wave["time"], 1e-300)) <= 1e-9
    flat = ngt.read(f"{GOLDEN}/arr16.flat.ngt.gz")
    scale = np.where(np.asarray(flat["node/type"])[np.asarray(wave["save_eq"])] == 3, 1e-6, 1e-12)     # vntol / abstol
    err = np.max(np.abs(v - wave["values"]) / np.maximum(np.abs(wave["values"]), scale[None, :]), axis=0)
    assert (err <= 1e-9).all(), err
    print("arr16 gpu bit-identical:", bool(np.array_equal(v, wave["values"])))


def _oracle_array(n, tmp):
    """the reference itself (oracle/_ref/ngspice_dump, prebuilt) run on an n x n array netlist: circuit dump, pivoting
    factors and waveforms, written under tmp (tests/golden/make_golden.py's own routine)"""
    import os, sys
    sys.path.insert(0, os.path.join(GOLDEN))
    import make_golden as M
    if not os.path.exists(M.DUMP):
        pytest.skip("oracle/_ref/ngspice_dump not built")
    src = open(os.path.join(GOLDEN, "netlists", "ro17.cir")).read()
    cards = src[src.index(".model"):src.rindex(".end")]
    M.HERE = tmp; M.TMP = os.path.join(tmp, "work")
    name = f"arr{n}"
    M.run(name, synth.inverter_array_netlist(n, n, cards), "0-1", ["out_0_0", f"out_{n - 1}_{n - 1}", "in_2_1", "vdd#branch"])
    return (ngt.read(os.path.join(tmp, name + ".flat.ngt")), ngt.read(os.path.join(tmp, name + ".trace.ngt.gz")),
            ngt.read(os.path.join(tmp, name + ".wave.ngt")))


@pytest.mark.gpu
def test_array48_gpu_against_the_oracle(cuda_lib, tmp_path):
    """BASELINE config 4 at 48 x 48 (4 608 BSIM4, 23 044 unknowns, supply-rail rows past NGB_ASM_LONG): the reference runs the
    netlist here, the device repeats DC operating point + `.tran 10p 1n` with the grid-wide LU on KLU's pivot orders.
    Identical accepted / rejected / iteration counts; 1e-9 per point (the rail rows are summed by a chunk tree on the device,
    so their last place may differ)"""
    from parity_util import run_patterns
    flat, trace, wave = _oracle_array(48, str(tmp_path))
    circ = pkg.Circuit.from_flat(cuda_lib, flat, lu_pattern=run_patterns(trace))
    b = pkg.Batch(circ, 1)
    res = b.tran(1024, wave["save_eq"])
    t, v = res.waves()
    n = int(res.npoints[0])
    acc, rej, nit = (int(x) for x in wave["stats"][:3])
    assert int(res.err[0]) == 0 and n == len(wave["time"])
    assert (int(res.accepted[0]), int(res.rejected[0]), int(res.numiter[0])) == (acc, rej, nit)
    assert np.max(np.abs(t[0, :n] - wave["time"]) / np.maximum(wave["time"], 1e-300)) <= 1e-9
    scale = np.where(np.asarray(flat["node/type"])[np.asarray(wave["save_eq"])] == 3, 1e-6, 1e-12)
    err = np.max(np.abs(v[0, :n] - wave["values"]) / np.maximum(np.abs(wave["values"]), scale[None, :]), axis=0)
    assert (err <= 1e-9).all(), err
    print("arr48 gpu: unknowns", circ.neq, "LU values", circ.lu_info()["nV"], "max rel err", float(err.max()),
          "bit-identical", bool(np.array_equal(v[0, :n], wave["values"])))
