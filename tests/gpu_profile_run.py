"""short Monte-Carlo transient used as the ncu target (kicked 17-stage oscillator, 20 ns):
   python gpu_profile_run.py [samples]"""
import sys, time
import numpy as np
from parity_util import GOLDEN, ngt, pkg, first_pattern
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lib = pkg.library()
flat = ngt.read(f"{GOLDEN}/ro17k.flat.ngt"); trace = ngt.read(f"{GOLDEN}/ro17k.trace.ngt.gz"); wave = ngt.read(f"{GOLDEN}/ro17k.wave.ngt")
circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=first_pattern(trace))
b = pkg.Batch(circ, S)
import os
dv = pkg.mc.draw_delvto(S, 34, seed=5)
if os.environ.get("NGB_DV0"):
    dv = dv * 0.0
b.put("b4.inst", pkg.mc.bsim4_inst_with_delvto(lib, flat, dv))
t0 = time.time(); res = b.tran(1024, wave["save_eq"][:1]); dt = time.time() - t0
print(f"S={S} ticks {res.ticks} time {dt:.3f}s us/tick {dt / res.ticks * 1e6:.1f} iters {int(res.numiter.astype(np.int64).sum())} "
      f"evals/s {34 * int(res.numiter.astype(np.int64).sum()) / dt:.3e}")
if len(sys.argv) > 2 and sys.argv[2] == "stages":
    # in-situ breakdown: every 16th Newton step goes kernel by kernel on one stream with an event after each stage
    import ctypes
    b2 = pkg.Batch(circ, S)
    b2.put("b4.inst", pkg.mc.bsim4_inst_with_delvto(lib, flat, dv))
    lib.L.ngbProfile(1, 16)
    res = b2.tran(1024, wave["save_eq"][:1])
    ms = (ctypes.c_double * 8)()
    n = lib.L.ngbProfileStages(ms)
    names = ["", "small loads", "bsim4_load", "assemble", "lu", "bsim4_lte", "control", ""]
    print("stages (us per sampled step, %d steps): " % n + "  ".join(f"{names[k]} {ms[k] / max(n, 1) * 1e3:.1f}" for k in range(1, 7))
          + f"  sum {sum(ms) / max(n, 1) * 1e3:.1f}")
    lib.L.ngbProfile(0, 1)
