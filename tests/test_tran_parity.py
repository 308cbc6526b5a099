"""DCtran parity: the device-resident transient driver against the reference's accepted time
points, waveforms (v(out), two internal stage nodes, i(vdd)) and run statistics.

hostsim (CPU): same libm as the reference => everything must be IDENTICAL, bit for bit, even
for the stock ring oscillator whose start-up grows out of rounding noise.
GPU: CUDA exp/log differ from glibc in the last place, so the noise-started oscillator cannot
be compared point by point; the deterministic-start variants (ro17k, Monte-Carlo samples) must
agree within 1e-9 of the waveform range with identical accepted/rejected/iteration counts."""
import numpy as np
import pytest
from parity_util import run_patterns, pattern_at, GOLDEN, ngt, pkg, first_pattern


def _run(lib, name, S=1, inst=None, max_points=8192):
    flat = ngt.read(f"{GOLDEN}/{name}.flat.ngt")
    trace = ngt.read(f"{GOLDEN}/{name}.trace.ngt.gz")
    wave = ngt.read(f"{GOLDEN}/{name}.wave.ngt")
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=run_patterns(trace))
    b = pkg.Batch(circ, S)
    if inst is not None:
        b.put("b4.inst", inst)
    res = b.tran(max_points, wave["save_eq"])
    t, v = res.waves()
    wave["_scale"] = _scales(flat, wave)
    return res, t, v, wave


def _scales(flat, wave):
    """|v - v_ref| <= tol * max(|v_ref|, scale) at every accepted point, scale = vntol (1e-6 V) for node
    voltages and abstol (1e-12 A) for branch currents"""
    nt = np.asarray(flat["node/type"])[np.asarray(wave["save_eq"])]
    return np.where(nt == 3, 1e-6, 1e-12)


def _compare(res, t, v, wave, s, exact, tol=1e-9, same_route=True):
    acc, rej, nit = (int(x) for x in wave["stats"][:3])
    assert int(res.accepted[s]) == acc and int(res.rejected[s]) == rej
    assert int(res.numiter[s]) == nit or not same_route
    n = int(res.npoints[s])
    assert n == len(wave["time"]) and int(res.err[s]) == 0
    if exact:
        assert np.array_equal(t[s, :n], wave["time"])
        assert np.array_equal(v[s, :n, :], wave["values"])
    else:
        assert np.max(np.abs(t[s, :n] - wave["time"]) / np.maximum(wave["time"], 1e-300)) <= tol     # t = 0 is the first point without UIC
        # per point, not per vector range
        scale = wave.get("_scale", np.full(wave["values"].shape[1], 1e-6))
        err = np.max(np.abs(v[s, :n, :] - wave["values"]) / np.maximum(np.abs(wave["values"]), scale[None, :]), axis=0)
        assert (err <= tol).all(), err


# invsrc / invgmin: the inverter with `.option noopiter` (+ `gminsteps=0`): CKTop goes straight to gillespie_src /
# dynamic_gmin (cktop.c:42-96), which run per sample inside the device controller
@pytest.mark.parametrize("name", ["ro17", "ro17k", "inv", "dio", "b3ring", "latch", "latchns", "srcs", "invsrc", "invgmin", "invshunt"])   # latchns: .nodeset on the operating point (the ipass rule of niiter.c:307-331)       # invshunt: gshunt=1e-9 ends the ladder and stays on the diagonal
def test_tran_hostsim_bit_identical(hostsim_lib, name):
    res, t, v, wave = _run(hostsim_lib, name)
    _compare(res, t, v, wave, 0, exact=True)


@pytest.mark.parametrize("name", ["ro17kg", "invg", "diog", "b3ringg"])
def test_tran_hostsim_gear_bit_identical(hostsim_lib, name):
    """`.option method=gear`: NIcomCof's GEAR system, NIintegrate's accumulation and CKTterr's GEAR coefficients
    (nicomcof.c:52-121, niinteg.c:42-71, cktterr.c:23-62) for BSIM4, BSIM3, diodes and capacitors"""
    res, t, v, wave = _run(hostsim_lib, name)
    _compare(res, t, v, wave, 0, exact=True)


def test_tran_hostsim_gear_mix_cell(hostsim_lib):
    """GEAR on the cell with every model family (VBIC: 1e-9, see test_tran_hostsim_vbic)"""
    res, t, v, wave = _run(hostsim_lib, "mixg")
    _compare(res, t, v, wave, 0, exact=False, tol=1e-8)      # per point; 4.6e-9 where v(y4) crosses zero (VBIC Jacobian rounding)


def test_bsim4_variant_kernels_same_bits(hostsim_lib):
    """the load specialised on the model selectors (csrc/bsim4_variants.h) and the generic one: same accepted points, same bits.
    The GEAR run of the same card has no specialised instantiation and reports so"""
    outs = []
    for generic in (False, True):
        flat = ngt.read(f"{GOLDEN}/ro17k.flat.ngt"); trace = ngt.read(f"{GOLDEN}/ro17k.trace.ngt.gz"); wave = ngt.read(f"{GOLDEN}/ro17k.wave.ngt")
        circ = pkg.Circuit.from_flat(hostsim_lib, flat, lu_pattern=run_patterns(trace))
        b = pkg.Batch(circ, 2)
        b.set_bsim4_generic(generic)
        key, special = b.bsim4_variant()
        assert key != 0xffffffff and special == (not generic)
        res = b.tran(8192, wave["save_eq"])
        outs.append((res.accepted.copy(), res.numiter.copy()) + res.waves())
        _compare(res, outs[-1][2], outs[-1][3], wave, 1, exact=True)
    assert all(np.array_equal(a, b) for a, b in zip(outs[0], outs[1]))
    flat = ngt.read(f"{GOLDEN}/ro17kg.flat.ngt"); trace = ngt.read(f"{GOLDEN}/ro17kg.trace.ngt.gz")
    b = pkg.Batch(pkg.Circuit.from_flat(hostsim_lib, flat, lu_pattern=run_patterns(trace)), 1)
    assert b.bsim4_variant()[1] is False


@pytest.mark.parametrize("name", ["invsrc", "invgmin"])
def test_tran_hostsim_fallback_batch(hostsim_lib, name):
    """the `.option noopiter` start state is set up for every sample of a batch, not only the first"""
    res, t, v, wave = _run(hostsim_lib, name, S=3)
    for s in range(3):
        _compare(res, t, v, wave, s, exact=True)


@pytest.mark.parametrize("name", ["vbicsh", "vbicxf", "vbicshxf"])
def test_tran_hostsim_vbic_selfheating_excess_phase(hostsim_lib, name):
    """VBIC electro-thermal (RTH / CTH, thermal node `dt`, DEVlimitlog) and excess phase (TD, the xf1 / xf2 filter nodes), alone
    and together (vbicload.c:668-672, 703-748, 1268-1478): identical accepted / rejected / iteration counts, node voltages,
    branch current AND the temperature rise of three devices within 1e-9 per point.  The d/dVrth partials are the tenth
    component of the same dual number the other partials come from"""
    res, t, v, wave = _run(hostsim_lib, name)
    _compare(res, t, v, wave, 0, exact=False)


@pytest.mark.parametrize("name", ["diosh", "diorr", "dioshrr"])
def test_tran_hostsim_diode_selfheating_soft_recovery(hostsim_lib, name):
    """diodes with the thermal terminal (rth0 / cth0: DIOtempUpdate at DIOtemp + delTemp inside every load, DEVlimitlog on
    the temperature rise, the d/dT stamps) and with the soft reverse-recovery charge node qp (vp, tt), alone and together
    (dioload.c:80-81, 317-322, 565-582, 736-862): every accepted point bit-identical, temperatures included"""
    res, t, v, wave = _run(hostsim_lib, name)
    _compare(res, t, v, wave, 0, exact=True)


def test_tran_hostsim_vbic(hostsim_lib):
    """VBIC stages (DC operating point + PULSE transient): identical accepted / rejected / iteration
    counts and 1e-9 on the waveforms.  Not bit-identical by construction: the Jacobian entries come from
    forward-mode dual numbers, equal to the reference's generated derivative code up to rounding."""
    res, t, v, wave = _run(hostsim_lib, "vbic")
    _compare(res, t, v, wave, 0, exact=False)


MIX_POINTS = [("2.0", "1k"), ("1.6", "200"), ("1.6", "5k"), ("2.4", "200"), ("2.4", "5k"),     # make_golden.py
              ("1.69098", "1348.2"),                          # the reference needs dynamic gmin stepping here (cktop.c:162)
              ("1.69412", "2364.7"), ("1.69412", "2383.5")]   # marginal: bound to the centre's pivot orders the batch needs it too
MIX_SAME_ROUTE = 8          # with every sample pivoting on its own matrix (csrc/ngb_pivot.c) all eight points take the reference's route: identical iteration counts (bound to the centre's pivot orders, round 1, the last two did not)


def _mix_sweep(lib, reps=1):
    """BASELINE config 5: the mixed BSIM4 + BSIM3 + VBIC + diode + R/C cell, every sample with its own supply
    voltage and interconnect resistor (centre + four corners of the sweep grid, plus three points whose
    operating point needs gmin stepping), in ONE batch"""
    flat = ngt.read(f"{GOLDEN}/mix.flat.ngt")
    trace = ngt.read(f"{GOLDEN}/mix.trace.ngt.gz")
    wave = ngt.read(f"{GOLDEN}/mix.wave.ngt")
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=run_patterns(trace))
    pts = MIX_POINTS * reps
    b = pkg.Batch(circ, len(pts))
    pkg.sweep.apply(b, flat, dc={"vdd": [p[0] for p in pts]}, res={"r1": [p[1] for p in pts]})
    res = b.tran(8192, wave["save_eq"])
    t, v = res.waves()
    for s in range(len(pts)):
        k = s % len(MIX_POINTS)
        w = ngt.read(f"{GOLDEN}/mix{k if k else ''}.wave.ngt"); w["_scale"] = _scales(flat, w)
        _compare(res, t, v, w, s, exact=False, same_route=k < MIX_SAME_ROUTE)


def test_tran_hostsim_mix_cell(hostsim_lib):
    """all model families of the path in one circuit (1e-9, identical step / iteration counts; not
    bit-identical because of the VBIC Jacobian, see test_tran_hostsim_vbic)"""
    res, t, v, wave = _run(hostsim_lib, "mix")
    _compare(res, t, v, wave, 0, exact=False)


def test_tran_hostsim_mix_sweep(hostsim_lib):
    _mix_sweep(hostsim_lib)


def _mix_source_stepping(lib, reps=1, exact_count=True):
    """CKTop with gmin stepping switched off (`.option gminsteps=0`) in a batch: the sweep point whose plain Newton
    iteration fails goes into gillespie_src (cktop.c:481-660) per sample inside the device controller while its
    neighbours, which converge directly, run their transients undisturbed.  In the reference source stepping FAILS for
    this point too (tests/golden/make_golden.py, "mixsrc"; it is OPtran that rescues it there, and OPtran is not on this
    path).  Every sample's own matrix is factored with pivoting at the reference's pivoting events (csrc/ngb_pivot.c, the
    default whenever klu_analyze's symbolic analysis travels with the pattern), so inside the batch the hard point takes the
    reference's route step for step: source stepping fails with E_ITERLIM after exactly the reference's 759 CKTop
    iterations (`exact_count`; the GPU's VBIC Jacobian differs in rounding, there the count may move).  Bound to the batch's
    recorded pivot orders instead (round 1) the zero-source matrix had an exact zero pivot and the sample ended with
    E_SINGULAR."""
    flat = ngt.read(f"{GOLDEN}/mix.flat.ngt")
    trace = ngt.read(f"{GOLDEN}/mix.trace.ngt.gz")
    wave = ngt.read(f"{GOLDEN}/mix.wave.ngt")
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=run_patterns(trace))
    circ.set_op_fallbacks(gminsteps=0, srcsteps=1)
    pts = [MIX_POINTS[0], MIX_POINTS[5], MIX_POINTS[1]] * reps
    b = pkg.Batch(circ, len(pts))
    pkg.sweep.apply(b, flat, dc={"vdd": [p[0] for p in pts]}, res={"r1": [p[1] for p in pts]})
    res = b.tran(8192, wave["save_eq"])
    t, v = res.waves()
    for s in range(len(pts)):
        if s % 3 != 1:
            w = ngt.read(f"{GOLDEN}/{('mix', '', 'mix1')[s % 3]}.wave.ngt"); w["_scale"] = _scales(flat, w)
            _compare(res, t, v, w, s, exact=False, same_route=True)
        else:
            ref = ngt.read(f"{GOLDEN}/mixsrc.wave.ngt")
            assert int(res.accepted[s]) == 0 and int(res.err[s]) == 103, (s, int(res.err[s]))
            if exact_count:
                assert int(res.numiter[s]) == int(ref["stats"][5]) == 759, int(res.numiter[s])
    return res


def _mix_source_stepping_route(lib, S=1, count_tol=0.0):
    """the same point as one circuit on the pivot orders the reference computed for it (the zero-source solve's for the
    operating point, the transient's for the transient): gillespie_src takes the reference's route step for step -- the
    first solve with the sources at zero, the adaptive raising of CKTsrcFact with its saved / restored solutions and
    states, the failure at 9.4 % of the supplies -- in exactly the reference's 759 CKTop iterations (`op_loads`)."""
    flat = ngt.read(f"{GOLDEN}/mixsrc.flat.ngt")
    trace = ngt.read(f"{GOLDEN}/mixsrc.trace.ngt.gz")
    wave = ngt.read(f"{GOLDEN}/mixsrc.wave.ngt")
    ks = sorted({int(k.split("/")[0][1:]) for k in trace if k.endswith("/pat/n")})
    assert ks == [0, 1, 101, 102, 759, 760]
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=[pattern_at(trace, k) for k in (101, 102, 759, 760)])
    b = pkg.Batch(circ, S)
    res = b.tran(8192, wave["save_eq"])
    op_loads = int(wave["stats"][5])
    assert op_loads == 759
    for s in range(S):
        assert int(res.err[s]) == 103 and int(res.accepted[s]) == 0           # "source stepping failed"
        assert abs(int(res.numiter[s]) - op_loads) <= count_tol * op_loads, int(res.numiter[s])
        assert int(res.numiter[s]) == int(res.numiter[0])


def test_tran_hostsim_mix_source_stepping_route(hostsim_lib):
    _mix_source_stepping_route(hostsim_lib)


@pytest.mark.gpu
def test_tran_gpu_mix_source_stepping_route(cuda_lib):
    _mix_source_stepping_route(cuda_lib, S=5, count_tol=0.25)     # VBIC Jacobian / CUDA rounding: the count is the host build's test


def test_tran_hostsim_mix_source_stepping(hostsim_lib):
    _mix_source_stepping(hostsim_lib)


def _op_chain_fails(lib, S=1, count_tol=0.0):
    """tolerances no Newton iteration can meet (reltol 1e-15): in the reference the plain NIiter, dynamic_gmin, new_gmin
    and gillespie_src fail one after the other (cktop.c:62-96; log lines kept in make_golden.py).  The sample must fail
    with E_ITERLIM after exactly the reference's 3355 CKTop iterations -- that count is the sum of the lengths of all
    three ladders (every failed NIiter is 101 iterations, the zero-source solve of gillespie_src converges)."""
    flat = ngt.read(f"{GOLDEN}/invfail.flat.ngt")
    trace = ngt.read(f"{GOLDEN}/invfail.trace.ngt.gz")
    stats = ngt.read(f"{GOLDEN}/invfail.wave.ngt")["stats"]
    assert int(stats[0]) == 0 and int(stats[6]) != 0              # the reference gave up too
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=run_patterns(trace))
    b = pkg.Batch(circ, S)
    res = b.tran(64, np.array([1], np.int32))
    err = res.err
    for s in range(S):
        assert int(res.accepted[s]) == 0 and int(res.npoints[s]) == 0
        assert int(err[s]) == 103                                  # E_ITERLIM, "source stepping failed"
        assert abs(int(res.numiter[s]) - int(stats[5])) <= count_tol * int(stats[5]), (int(res.numiter[s]), int(stats[5]))
        assert int(res.numiter[s]) == int(res.numiter[0])
    assert np.all(b.get("ctl.srcfact") == 1.0)                    # "no path out of this code allows CKTsrcFact to be anything but 1"


def test_op_fallback_chain_fails_like_reference(hostsim_lib):
    _op_chain_fails(hostsim_lib)


@pytest.mark.gpu
def test_op_fallback_chain_fails_like_reference_gpu(cuda_lib):
    # whether an iteration "converges" under reltol 1e-15 hangs on the last bit: the exact count is the host build's test,
    # on the device the chain must fail the same way, every sample alike, with a count in the reference's range
    _op_chain_fails(cuda_lib, S=33, count_tol=0.1)


def _timestep_too_small(lib, S=1, count_tol=0.0):
    """tolerances the transient cannot hold (reltol 9e-13): the reference abandons the run with E_TIMESTEP ("timestep too
    small", dctran.c:901-913) after 1044 accepted and 284 rejected points and 43 314 iterations; every sample of the batch
    must stop at the same point with the same code"""
    flat = ngt.read(f"{GOLDEN}/invtstep.flat.ngt")
    trace = ngt.read(f"{GOLDEN}/invtstep.trace.ngt.gz")
    acc, rej, nit, _, _, _, ret = (int(x) for x in ngt.read(f"{GOLDEN}/invtstep.wave.ngt")["stats"])
    assert ret == 106
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=run_patterns(trace))
    b = pkg.Batch(circ, S)
    res = b.tran(8192, np.array([1], np.int32))
    err = res.err
    for s in range(S):
        assert int(err[s]) == 106
        got = (int(res.accepted[s]), int(res.rejected[s]), int(res.numiter[s]))
        assert all(abs(g - r) <= count_tol * r for g, r in zip(got, (acc, rej, nit))), (got, (acc, rej, nit))
        assert got == (int(res.accepted[0]), int(res.rejected[0]), int(res.numiter[0]))


def test_tran_hostsim_timestep_too_small(hostsim_lib):
    _timestep_too_small(hostsim_lib)


@pytest.mark.gpu
def test_tran_gpu_timestep_too_small(cuda_lib):
    _timestep_too_small(cuda_lib, S=3, count_tol=0.1)       # the point of failure hangs on rounding noise; exact on the host build


def test_op_fallback_options_refused(hostsim_lib):
    """spice3_gmin / spice3_src (counts > 1) are not on this path: refused when the option is set, not at run time"""
    flat = ngt.read(f"{GOLDEN}/inv.flat.ngt")
    circ = pkg.Circuit.from_flat(hostsim_lib, flat)
    with pytest.raises(pkg.NgbError):
        circ.set_op_fallbacks(gminsteps=10)
    with pytest.raises(pkg.NgbError):
        circ.set_op_fallbacks(srcsteps=5)
    circ.set_op_fallbacks(gminsteps=0, srcsteps=0)


@pytest.mark.gpu
def test_tran_gpu_mix_source_stepping(cuda_lib):
    _mix_source_stepping(cuda_lib, reps=12, exact_count=False)     # 36 samples: the stepping sample shares its warps with direct ones


@pytest.mark.gpu
def test_tran_gpu_mix_sweep(cuda_lib):
    _mix_sweep(cuda_lib, reps=9)        # 72 samples: more than two warps, every point on its own time axis


def _mc_inst(lib):
    base = ngt.read(f"{GOLDEN}/ro17k.flat.ngt")
    dv_netlist = np.load(f"{GOLDEN}/ro17mc.delvto.npy")       # columns: mp1 mn1 mp2 mn2 ... (netlist order)
    col = {}
    for k in range(1, 18):
        col[f"mp{k}"] = 2 * (k - 1); col[f"mn{k}"] = 2 * (k - 1) + 1
    order = [col[n.lower()] for n in pkg.mc.instance_names(base)]   # flat tables are in reference list order
    dv = pkg.mc.delvto_as_parsed(dv_netlist[:, order])      # what INPevaluate makes of the netlist text
    return base, dv, pkg.mc.bsim4_inst_with_delvto(lib, base, dv)


def test_mc_mismatch_parameters_match_bsim4temp(hostsim_lib):
    """the host-side delvto applicator reproduces what BSIM4temp wrote into the instances"""
    base, dv, inst = _mc_inst(hostsim_lib)
    for i in range(dv.shape[0]):
        ref = ngt.read(f"{GOLDEN}/ro17mc{i}.flat.ngt")["b4/inst"]
        err = np.abs(inst[:, :, i] - ref) / np.maximum(np.abs(ref), 1e-300)
        assert np.array_equal(inst[:, :, i], ref), (i, err.max())


def test_tran_hostsim_mc_batch(hostsim_lib):
    """four mismatch samples advanced in ONE batch, each on its own time axis, each equal to
    its own sequential reference run"""
    base, dv, inst = _mc_inst(hostsim_lib)
    # use the reference's own instance tables so the comparison is exact
    for i in range(dv.shape[0]):
        inst[:, :, i] = ngt.read(f"{GOLDEN}/ro17mc{i}.flat.ngt")["b4/inst"]
    res, t, v, _ = _run(hostsim_lib, "ro17k", S=dv.shape[0], inst=inst)
    for i in range(dv.shape[0]):
        wave = ngt.read(f"{GOLDEN}/ro17mc{i}.wave.ngt")
        _compare(res, t, v, wave, i, exact=True)


# GPU: with -fmad=false, IEEE division/sqrt and the libm-compatible exp/log of csrc/ngb_math.cuh the
# device follows the reference bit for bit; the assertions below allow 1e-9 (the north_star
# tolerance) but identical accepted / rejected / iteration counts are required.
@pytest.mark.gpu
@pytest.mark.parametrize("name,exact", [("ro17k", True), ("ro17", True), ("inv", True), ("dio", False), ("b3ring", True), ("vbic", False), ("latch", True), ("srcs", False),
                                        ("invsrc", False), ("invgmin", False), ("invshunt", False),
                                        ("diosh", False), ("diorr", False), ("dioshrr", False)])    # diode self-heating / soft recovery: north_star bar on the device, bit-identity shown on the host build
def test_tran_gpu_matches_reference(cuda_lib, name, exact):
    """north_star bar: 1e-9 relative and identical accepted-step count; the device arithmetic
    (no FMA contraction, glibc-compatible exp/log) in fact reproduces the reference bit for bit.
    `dio` is driven by a SIN source: CUDA's sin() is not glibc's, so only the 1e-9 bar applies."""
    res, t, v, wave = _run(cuda_lib, name)
    _compare(res, t, v, wave, 0, exact=False)
    if exact:
        _compare(res, t, v, wave, 0, exact=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["vbicsh", "vbicxf", "vbicshxf"])
def test_tran_gpu_vbic_selfheating_excess_phase(cuda_lib, name):
    res, t, v, wave = _run(cuda_lib, name, S=33)
    for s in (0, 32):
        _compare(res, t, v, wave, s, exact=False)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ro17kg", "invg", "diog", "b3ringg", "mixg"])
def test_tran_gpu_gear_matches_reference(cuda_lib, name):
    """`.option method=gear` on the device (1e-9, identical step and iteration counts); these runs also take the GENERIC
    BSIM4 load kernel (no specialised instantiation carries gear = 1)"""
    res, t, v, wave = _run(cuda_lib, name)
    _compare(res, t, v, wave, 0, exact=False, tol=1e-8 if name == "mixg" else 1e-9)     # mixg: see test_tran_hostsim_gear_mix_cell


@pytest.mark.gpu
def test_bsim4_variant_kernels_same_bits_gpu(cuda_lib):
    """specialised against generic load kernel on the device: bit-identical waveforms for 64 mismatch samples"""
    flat = ngt.read(f"{GOLDEN}/ro17k.flat.ngt"); trace = ngt.read(f"{GOLDEN}/ro17k.trace.ngt.gz"); wave = ngt.read(f"{GOLDEN}/ro17k.wave.ngt")
    circ = pkg.Circuit.from_flat(cuda_lib, flat, lu_pattern=run_patterns(trace))
    dv = pkg.mc.draw_delvto(64, 34, seed=11)
    outs = []
    for generic in (False, True):
        b = pkg.Batch(circ, 64)
        b.put("b4.inst", pkg.mc.bsim4_inst_with_delvto(cuda_lib, flat, dv))
        b.set_bsim4_generic(generic)
        assert b.bsim4_variant()[1] == (not generic)
        res = b.tran(2048, wave["save_eq"])
        assert (res.err == 0).all()
        outs.append((res.accepted.copy(), res.rejected.copy(), res.numiter.copy()) + res.waves())
    assert all(np.array_equal(a, b) for a, b in zip(outs[0], outs[1]))


@pytest.mark.gpu
def test_tran_gpu_ro101_single_circuit(cuda_lib):
    """BASELINE config 2: 101-stage oscillator, one circuit (CTA-per-sample LU)"""
    res, t, v, wave = _run(cuda_lib, "ro101")
    _compare(res, t, v, wave, 0, exact=False)


@pytest.mark.gpu
def test_tran_gpu_mc_batch(cuda_lib):
    """Monte-Carlo batch: 32 samples (4 distinct mismatch draws x 8), every lane of a warp on its
    own time axis, each compared with its own sequential reference run"""
    base, dv, inst = _mc_inst(cuda_lib)
    for i in range(dv.shape[0]):
        inst[:, :, i] = ngt.read(f"{GOLDEN}/ro17mc{i}.flat.ngt")["b4/inst"]
    reps = 8
    inst = np.tile(inst, (1, 1, reps))
    res, t, v, _ = _run(cuda_lib, "ro17k", S=dv.shape[0] * reps, inst=inst)
    for s in range(dv.shape[0] * reps):
        wave = ngt.read(f"{GOLDEN}/ro17mc{s % dv.shape[0]}.wave.ngt")
        _compare(res, t, v, wave, s, exact=False)


@pytest.mark.gpu
def test_tran_gpu_mc_host_side_mismatch(cuda_lib):
    """the same batch with the mismatch applied by the host-side helper (mc.py) instead of
    BSIM4temp: identical parameters, hence identical waveforms and step counts"""
    base, dv, inst = _mc_inst(cuda_lib)
    res, t, v, _ = _run(cuda_lib, "ro17k", S=dv.shape[0], inst=inst)
    for s in range(dv.shape[0]):
        wave = ngt.read(f"{GOLDEN}/ro17mc{s}.wave.ngt")
        _compare(res, t, v, wave, s, exact=True)


def _tox_batch(lib):
    """two samples with different oxide-thickness levels AND per-instance delvto, in one batch"""
    base = ngt.read(f"{GOLDEN}/ro17k.flat.ngt")
    trace = ngt.read(f"{GOLDEN}/ro17k.trace.ngt.gz")
    tables = ngt.read(f"{GOLDEN}/ro17tox.tables.ngt")
    dv_netlist = np.load(f"{GOLDEN}/ro17tox.delvto.npy"); level = np.load(f"{GOLDEN}/ro17tox.level.npy")
    col = {}
    for k in range(1, 18):
        col[f"mp{k}"] = 2 * (k - 1); col[f"mn{k}"] = 2 * (k - 1) + 1
    order = [col[n.lower()] for n in pkg.mc.instance_names(base)]
    dv = pkg.mc.delvto_as_parsed(dv_netlist[:, order])
    inst, prow_t, mtab, ptab = pkg.mc.bsim4_with_tox_levels(lib, base, tables, level, dv)
    for i in range(2):                       # the assembled tables equal what BSIM4temp wrote for each sample
        ref = ngt.read(f"{GOLDEN}/ro17tox{i}.flat.ngt")
        assert np.array_equal(inst[:, :, i], ref["b4/inst"])
        assert np.array_equal(mtab[prow_t.reshape(34, 2)[:, i]], ref["b4/mtab"][ref["b4/prow"]])
        assert np.array_equal(ptab[prow_t.reshape(34, 2)[:, i]], ref["b4/ptab"][ref["b4/prow"]])
    circ = pkg.Circuit.from_flat(lib, base, lu_pattern=run_patterns(trace))
    b = pkg.Batch(circ, 2)
    b.put("b4.inst", inst)
    b.set_bsim4_rows(prow_t, mtab, ptab)
    wave0 = ngt.read(f"{GOLDEN}/ro17tox0.wave.ngt")
    res = b.tran(1024, wave0["save_eq"])
    t, v = res.waves()
    return res, t, v


def test_tran_hostsim_tox_and_vth_mismatch(hostsim_lib):
    """model-parameter (toxe) plus instance (delvto) mismatch: each sample equals its own reference run"""
    res, t, v = _tox_batch(hostsim_lib)
    for s in range(2):
        _compare(res, t, v, ngt.read(f"{GOLDEN}/ro17tox{s}.wave.ngt"), s, exact=True)


@pytest.mark.gpu
def test_tran_gpu_tox_and_vth_mismatch(cuda_lib):
    res, t, v = _tox_batch(cuda_lib)
    for s in range(2):
        _compare(res, t, v, ngt.read(f"{GOLDEN}/ro17tox{s}.wave.ngt"), s, exact=True)


def _tox_batch_own_temp(lib, copies=1):
    """the same two samples, but their model / bin / instance rows come from the library's own BSIM4temp (csrc/ngb_b4temp.c)
    applied to the NOMINAL card with each sample's toxe and delvto -- nothing recorded per oxide thickness -- one set of rows per sample"""
    base = ngt.read(f"{GOLDEN}/ro17k.flat.ngt")
    trace = ngt.read(f"{GOLDEN}/ro17k.trace.ngt.gz")
    tab = ngt.read(f"{GOLDEN}/b4temp.tables.ngt.gz")
    raw = {"model": tab["ro17k/b4t/model"], "inst": tab["ro17k/b4t/inst"], "inst_model": tab["ro17k/b4t/inst_model"],
           "temp": tab["ro17k/b4t/temp"][0, 0], "vt0": tab["ro17k/opt/vt0"][0]}
    levels = ngt.read(f"{GOLDEN}/ro17tox.tables.ngt")["levels"]
    dv_netlist = np.load(f"{GOLDEN}/ro17tox.delvto.npy"); level = np.load(f"{GOLDEN}/ro17tox.level.npy")
    col = {}
    for k in range(1, 18):
        col[f"mp{k}"] = 2 * (k - 1); col[f"mn{k}"] = 2 * (k - 1) + 1
    order = [col[n.lower()] for n in pkg.mc.instance_names(base)]
    dv = np.tile(pkg.mc.delvto_as_parsed(dv_netlist[:, order]), (copies, 1))
    toxe = np.tile(np.array([pkg.mc.spice_number(f"{levels[k]:.17g}") for k in level]), copies)
    inst, prow_t, mtab, ptab = pkg.mc.bsim4_with_toxe(lib, raw, toxe, dv)
    circ = pkg.Circuit.from_flat(lib, base, lu_pattern=run_patterns(trace))
    b = pkg.Batch(circ, 2 * copies)
    b.put("b4.inst", inst)
    b.set_bsim4_rows(prow_t, mtab, ptab)
    assert b.bsim4_variant()[1]            # the specialised kernel serves per-sample rows too
    wave0 = ngt.read(f"{GOLDEN}/ro17tox0.wave.ngt")
    res = b.tran(1024, wave0["save_eq"])
    t, v = res.waves()
    return res, t, v


def test_tran_hostsim_continuous_tox_own_bsim4temp(hostsim_lib):
    res, t, v = _tox_batch_own_temp(hostsim_lib)
    for s in range(2):
        _compare(res, t, v, ngt.read(f"{GOLDEN}/ro17tox{s}.wave.ngt"), s, exact=True)


@pytest.mark.gpu
def test_tran_gpu_continuous_tox_own_bsim4temp(cuda_lib):
    res, t, v = _tox_batch_own_temp(cuda_lib, copies=48)       # 96 samples: three warps per instance
    for s in range(96):
        _compare(res, t, v, ngt.read(f"{GOLDEN}/ro17tox{s % 2}.wave.ngt"), s, exact=True)


B3_CAP_CASES = [f"b3c{cm}x{tag}" for cm in (0, 1, 2, 3) for tag in ("0", "5", "1")]


@pytest.mark.parametrize("name", B3_CAP_CASES)
def test_tran_hostsim_bsim3_capmod_xpart(hostsim_lib, name):
    """every capMod (0-3) x charge partition (xpart 0, 0.5, 1) of BSIM3v3.3.0"""
    res, t, v, wave = _run(hostsim_lib, name)
    _compare(res, t, v, wave, 0, exact=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", B3_CAP_CASES)
def test_tran_gpu_bsim3_capmod_xpart(cuda_lib, name):
    res, t, v, wave = _run(cuda_lib, name)
    _compare(res, t, v, wave, 0, exact=True)
