"""Prints the error statistics of the GPU kernels against the recorded reference calls and the
timing of a few runs (development aid; run on the GPU box)."""
import sys, time
import numpy as np
from parity_util import GOLDEN, ngt, pkg, relerr, replay_load, trace_calls, first_pattern

lib = pkg.library()
print("backend", lib.backend)
for name in ("ro17", "ro101"):
    flat = ngt.read(f"{GOLDEN}/{name}.flat.ngt"); trace = ngt.read(f"{GOLDEN}/{name}.trace.ngt.gz")
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=first_pattern(trace))
    b = None
    for call in trace_calls(trace):
        b, ours, ref, maps = replay_load(lib, circ, flat, trace, call, S=4, batch=b)
        c = f"c{call}/"
        eA = relerr(ours["Ax"][3], ref["Ax"], 1e-300).max()
        er = relerr(ours["x"][1, 1:, 3], ref["rhs"][1:], 1e-300).max()
        es = relerr(ours["b4_state"][0, :, :, 3], ref["state0"][maps["b4"]], 1e-300).max()
        eo = relerr(ours["b4_op"][:, :, 3], ref["b4_op"], 1e-300).max(axis=1)
        line = f"{name} call {call}: Ax {eA:.1e} rhs {er:.1e} state {es:.1e} op {eo.max():.1e} ({lib.fields['op'][int(eo.argmax())]})"
        if c + "sol" in trace:
            b.lufac_solve()
            x = b.get("x", (2, circ.neq + 1, 4))
            line += f" sol {relerr(x[1, 1:, 3], trace[c + 'sol'][1:circ.neq + 1], 1e-30).max():.1e}"
        print(line)
for name, S in (("ro17k", 1), ("ro17k", 256), ("ro17", 1), ("ro17", 1024), ("ro101", 1)):
    flat = ngt.read(f"{GOLDEN}/{name}.flat.ngt"); trace = ngt.read(f"{GOLDEN}/{name}.trace.ngt.gz"); wave = ngt.read(f"{GOLDEN}/{name}.wave.ngt")
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=first_pattern(trace))
    b = pkg.Batch(circ, S)
    t0 = time.time(); res = b.tran(8192, wave["save_eq"]); dt = time.time() - t0
    t, v = res.waves(); n = int(res.npoints[0]); m = min(n, len(wave["time"]))
    dv = np.abs(v[0, :m, :] - wave["values"][:m]).max(axis=0) / np.abs(wave["values"]).max(axis=0)
    print(f"tran {name} S={S}: {dt:.3f}s ticks {res.ticks} acc {res.accepted[0]} rej {res.rejected[0]} iter {res.numiter[0]} "
          f"(ref {wave['stats'][:3]}) npts {n}/{len(wave['time'])} wave err {dv} us/tick {dt / res.ticks * 1e6:.1f}")
