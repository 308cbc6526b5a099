"""Shared helpers for the parity tests: replay one recorded CKTload call of the reference
(oracle/ref_hooks.c trace) through the C ABI and return what the reference produced beside
what we produced."""
import importlib
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
pkg = importlib.import_module("ngspice-sf-mirror_b200")
ngt = pkg.ngt

HOSTSIM = os.path.join(ROOT, "tests", "hostsim", "libngb200_hostsim.so")
GOLDEN = os.path.join(ROOT, "tests", "golden")


def trace_calls(trace):
    return sorted({int(k.split("/")[0][1:]) for k in trace if k.startswith("c") and k.endswith("/mode")})


# the pivoting factors a recorded run holds: helpers of the package (the bench uses them too)
first_pattern, run_patterns, pattern_at = ngt.first_pattern, ngt.run_patterns, ngt.pattern_at


def state_maps(flat, lib):
    """index arrays mapping reference CKTstate offsets to our [k][inst] layouts"""
    out = {}
    n = ngt.scalar(flat, "b4/ninst", 0)
    if n:
        nst = lib.layout[6]
        out["b4"] = flat["b4/state_base"][None, :] + np.arange(nst)[:, None]     # [29][n]
    n = ngt.scalar(flat, "cap/n", 0)
    if n:
        out["cap"] = flat["cap/state_base"][None, :] + np.arange(2)[:, None]
    n = ngt.scalar(flat, "b3/ninst", 0)
    if n:
        out["b3"] = flat["b3/state_base"][None, :] + np.arange(lib.b3_layout[5])[:, None]
    n = ngt.scalar(flat, "dio/n", 0)
    if n:
        out["dio"] = flat["dio/state_base"][None, :] + np.arange(lib.dio_layout[1])[:, None]
    n = ngt.scalar(flat, "vbic/n", 0)
    if n:
        out["vbic"] = flat["vbic/state_base"][None, :] + np.arange(lib.vbic_layout[3])[:, None]
    return out


def replay_load(lib, circ, flat, trace, call, S=1, batch=None):
    """Set up a batch exactly as the reference circuit stood before CKTload call `call`,
    run ngbLoad, and return (batch, ours, ref) dictionaries."""
    c = f"c{call}/"
    b = batch or pkg.Batch(circ, S)
    S = b.S
    rep = lambda v, dt: np.full(S, v, dtype=dt)
    mode = int(trace[c + "mode"][0])
    b.put("ctl.mode", rep(mode, np.int32))
    b.put("ctl.active", rep(1, np.int32))
    b.put("ctl.head", rep(0, np.int32))
    b.put("ctl.order", rep(int(trace[c + "order"][0]), np.int32))
    b.put("ctl.xsel", rep(0, np.int32))
    ag = trace[c + "ag"]
    b.put("ctl.ag0", rep(ag[0], np.float64)); b.put("ctl.ag1", rep(ag[1], np.float64))
    b.put("ctl.delta", rep(trace[c + "delta"][0], np.float64))
    b.put("ctl.delta_old", np.repeat(trace[c + "deltaOld"], S))
    b.put("ctl.time", rep(trace[c + "time"][0], np.float64))
    b.put("ctl.gmin", rep(trace[c + "gmin"][0], np.float64))
    b.put("ctl.diag_gmin", rep(0.0, np.float64))
    b.put("ctl.srcfact", rep(trace[c + "srcfact"][0], np.float64))
    neq1 = circ.neq + 1
    x = np.zeros((2, neq1, S))
    x[0] = trace[c + "rhsOld"][:neq1, None]
    b.put("x", x)
    maps = state_maps(flat, lib)
    hist = [trace[c + "state0_in"], trace[c + "state1_in"], trace.get(c + "state2_in")]
    for dev, key in (("b4", "b4.state"), ("cap", "cap.state"), ("dio", "dio.state"), ("b3", "b3.state"), ("vbic", "vbic.state")):
        if dev not in maps:
            continue
        m = maps[dev]                                  # [k][n]
        st = np.zeros((4,) + m.shape + (S,))
        for h in range(3):
            if hist[h] is not None:
                st[h] = hist[h][m][:, :, None]
        b.put(key, st)
    if "b4" in maps:
        nop = lib.layout[7]
        n = maps["b4"].shape[1]
        op = np.zeros((nop, n, S))
        op_in = trace[c + "b4_op_in"]
        op[0] = op_in[0][:, None]                      # von is the only field the load reads back
        b.put("b4.op", op)
        b.set_op_full(True)
    if "b3" in maps:
        b.put("b3.von", np.repeat(trace[c + "b3_von_in"][:, None], S, axis=1))
    b.load()
    ours = dict(Ax=b.get("Ax", (S, -1)), x=b.get("x", (2, neq1, S)), noncon=b.get("ctl.noncon"))
    if "b4" in maps:
        ours["b4_state"] = b.get("b4.state", (4,) + maps["b4"].shape + (S,))
        ours["b4_op"] = b.get("b4.op", (lib.layout[7], maps["b4"].shape[1], S))
    if "cap" in maps:
        ours["cap_state"] = b.get("cap.state", (4,) + maps["cap"].shape + (S,))
    if "b3" in maps:
        ours["b3_state"] = b.get("b3.state", (4,) + maps["b3"].shape + (S,))
    if "dio" in maps:
        ours["dio_state"] = b.get("dio.state", (4,) + maps["dio"].shape + (S,))
    if "vbic" in maps:
        ours["vbic_state"] = b.get("vbic.state", (4,) + maps["vbic"].shape + (S,))
    ref = dict(Ax=trace[c + "Ax"], rhs=trace[c + "rhs"][:neq1], noncon=int(trace[c + "noncon"][0]),
               state0=trace[c + "state0_out"], state1=trace.get(c + "state1_out"), mode=mode)
    if c + "b4_op_out" in trace:
        ref["b4_op"] = trace[c + "b4_op_out"]
        ref["b4_state1"] = trace[c + "b4_state1"]
    return b, ours, ref, maps


def relerr(a, b, floor=0.0):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.where(den > 0, np.abs(a - b) / den, 0.0)
    return r
