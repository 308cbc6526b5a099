"""CKTload parity: replay recorded calls of the reference (BSIM4load + CAPload + VSRCload on a
17- and a 101-stage BSIM4 ring oscillator) through the C ABI and compare Ax, rhs, state0, the
operating point and CKTnoncon with what the reference produced.

CPU (`not gpu`): kernel bodies compiled for the host (tests/hostsim) -- same libm as the
reference, so state/op-point must agree bit for bit and Ax/rhs to summation-order rounding.
GPU: the CUDA library, same checks (its exp/log replicate glibc's, csrc/ngb_math.cuh); the
tolerances passed below are the north_star's 1e-9 per element and 1e-12 of the column scale."""
import numpy as np
import pytest
from parity_util import GOLDEN, ngt, pkg, relerr, replay_load, trace_calls, first_pattern

DUAL = ("vbic", "mix", "vbicsh", "vbicxf", "vbicshxf")      # fixtures holding VBIC devices: derivatives by dual numbers, see vbic_eval.cuh
CASES = ["ro17", "ro101", "inv", "dio", "b3ring", "vbic", "mix", "latch", "srcs", "vbicsh", "vbicxf", "vbicshxf", "diosh", "diorr", "dioshrr"]   # latch: .nodeset / .ic row overrides; srcs: PWL / EXP / SFFM / AM sources      # inv: DC operating point (INITJCT/INITFIX/INITFLOAT) + PULSE transient


def _load_case(name):
    flat = ngt.read(f"{GOLDEN}/{name}.flat.ngt")
    trace = ngt.read(f"{GOLDEN}/{name}.trace.ngt.gz")
    return flat, trace


def _scaled_err(a, ref, Ap=None):
    """error relative to the largest magnitude in the same matrix column (or in the vector)"""
    a = np.asarray(a); ref = np.asarray(ref)
    if Ap is None:
        return np.abs(a - ref).max() / max(np.abs(ref).max(), 1e-300)
    worst = 0.0
    for j in range(len(Ap) - 1):
        lo, hi = Ap[j], Ap[j + 1]
        if hi > lo:
            worst = max(worst, np.abs(a[lo:hi] - ref[lo:hi]).max() / max(np.abs(ref[lo:hi]).max(), 1e-300))
    return worst


def _check(lib, name, tol_state, tol_mat, S=1, tol_scaled=None):
    flat, trace = _load_case(name)
    circ = pkg.Circuit.from_flat(lib, flat)
    pat = circ.pattern()
    # SMPconvertCOOtoCSC + BSIM4bindCSC parity: identical CSC pattern and slot map
    assert pat["n"] == int(flat["klu/n"][0]) and pat["nnz"] == int(flat["klu/nz"][0])
    assert np.array_equal(pat["Ap"], flat["klu/Ap"]) and np.array_equal(pat["Ai"], flat["klu/Ai"])
    assert np.array_equal(pat["diag"], flat["klu/diag"])
    if ngt.scalar(flat, "b4/ninst", 0):
        assert np.array_equal(circ.bsim4_slots(int(flat["b4/ninst"][0])), flat["b4/slots"])
    calls = trace_calls(trace)
    assert calls
    batch = None
    for call in calls:
        batch, ours, ref, maps = replay_load(lib, circ, flat, trace, call, S=S, batch=batch)
        for s in sorted({0, S - 1}):
            assert relerr(ours["Ax"][s], ref["Ax"], 1e-300).max() <= tol_mat, (name, call, "Ax")
            if tol_scaled is not None:
                assert _scaled_err(ours["Ax"][s], ref["Ax"], pat["Ap"]) <= tol_scaled, (name, call, "Ax scaled")
                assert _scaled_err(ours["x"][1, 1:, s], ref["rhs"][1:]) <= tol_scaled, (name, call, "rhs scaled")
            assert relerr(ours["x"][1, 1:, s], ref["rhs"][1:], 1e-300).max() <= tol_mat, (name, call, "rhs")
            if "cap" in maps:
                assert relerr(ours["cap_state"][0, :, :, s], ref["state0"][maps["cap"]], 1e-300).max() <= tol_state
            if "b3" in maps:
                used = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 16]        # the NQS states 12-15 stay untouched
                assert relerr(ours["b3_state"][0, :, :, s][used], ref["state0"][maps["b3"]][used], 1e-300).max() <= tol_state, (name, call, "b3 state0")
                if ref["mode"] & 0x1000 and ref["state1"] is not None:
                    qrows = [4, 5, 6, 7, 8, 9]
                    assert relerr(ours["b3_state"][1, :, :, s][qrows], ref["state1"][maps["b3"]][qrows], 1e-300).max() <= tol_state
            if "dio" in maps:
                assert relerr(ours["dio_state"][0, :, :, s], ref["state0"][maps["dio"]], 1e-300).max() <= tol_state, (name, call, "dio state0")
                if ref["mode"] & 0x1000 and ref["state1"] is not None:
                    qrows = [6, 7, 10, 11, 15, 16]         # capCharge, capCurrent, qth, cqth, srcapCharge, srcapCurrent
                    assert relerr(ours["dio_state"][1, :, :, s][qrows], ref["state1"][maps["dio"]][qrows], 1e-300).max() <= tol_state
            if "vbic" in maps:
                # currents, charges, capacitor currents and limited voltages are bit-exact; the partial
                # derivatives come from forward-mode duals instead of the reference's generated
                # expressions and agree to rounding.  d/dVrth states (18, 70, 79) are dead without self-heating
                exact = [0, 1, 2, 3, 4, 5, 6, 7, 8, 11, 13, 15, 20, 23, 25, 29, 33, 37, 38, 40, 41, 42, 43, 44, 45, 46, 47,
                         49, 50, 52, 53, 55, 57, 61, 62]
                # (computed since self-heating is on the path; in fixtures without it the reference leaves 70 untouched)
                live = [k for k in range(maps["vbic"].shape[0]) if k not in ((18, 70, 79) if name in ("vbic", "mix") else ())]
                st = ours["vbic_state"][0, :, :, s]; rs = ref["state0"][maps["vbic"]]
                assert np.array_equal(st[exact], rs[exact]), (name, call, "vbic value states")
                # (floor: conductances 1e6 below gmin are differences of cancelling terms, e.g. d(avalanche)/dVbei ~ 1e-35 S)
                assert relerr(st[live], rs[live], 1e-18).max() <= 1e-12, (name, call, "vbic derivative states")
            assert (ours["noncon"][s] != 0) == (ref["noncon"] != 0), (name, call, "noncon")
            if "b4" not in maps:
                continue
            st0 = ours["b4_state"][0, :, :, s]
            assert relerr(st0, ref["state0"][maps["b4"]], 1e-300).max() <= tol_state, (name, call, "state0")
            assert relerr(ours["b4_op"][:, :, s], ref["b4_op"], 1e-300).max() <= tol_state, (name, call, "op")
            if ref["mode"] & 0x1000:   # MODEINITTRAN copies q0 -> state1
                st1 = ours["b4_state"][1, :, :, s]
                qrows = [11, 13, 15, 19, 21]
                assert relerr(st1[qrows], ref["b4_state1"][maps["b4"]][qrows], 1e-300).max() <= tol_state


@pytest.mark.parametrize("name", CASES)
def test_load_hostsim_matches_reference(hostsim_lib, name):
    if name in DUAL:        # Jacobian entries from dual numbers: rounding-level differences, scaled per column
        _check(hostsim_lib, name, tol_state=0.0, tol_mat=1.0, tol_scaled=1e-12)     # element-wise relative error is meaningless where contributions cancel
    else:
        _check(hostsim_lib, name, tol_state=0.0, tol_mat=1e-14)


def test_load_hostsim_batched_samples_identical(hostsim_lib):
    _check(hostsim_lib, "ro17", tol_state=0.0, tol_mat=1e-14, S=5)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_load_gpu_matches_reference(cuda_lib, name):
    _check(cuda_lib, name, tol_state=1e-9, tol_mat=1.0 if name in DUAL else 1e-9, S=1, tol_scaled=1e-12)


@pytest.mark.gpu
def test_load_gpu_batched(cuda_lib):
    _check(cuda_lib, "ro17", tol_state=1e-9, tol_mat=1e-9, S=67, tol_scaled=1e-12)
