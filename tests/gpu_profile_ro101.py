"""single-circuit latency (BASELINE config 2): the 101-stage oscillator, S = 1, first 20 ns; per-stage breakdown
   python gpu_profile_ro101.py"""
import ctypes, sys, time
import numpy as np
from parity_util import GOLDEN, ngt, pkg, run_patterns
lib = pkg.library()
flat = ngt.read(f"{GOLDEN}/ro101.flat.ngt"); trace = ngt.read(f"{GOLDEN}/ro101.trace.ngt.gz"); wave = ngt.read(f"{GOLDEN}/ro101.wave.ngt")
flat["tran/tstop"] = np.array([pkg.mc.spice_number("20ns")])
circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=run_patterns(trace))
for rep in range(2):
    b = pkg.Batch(circ, 1)
    t0 = time.time(); res = b.tran(4096, wave["save_eq"][:1]); dt = time.time() - t0
print(f"ro101 S=1 ticks {res.ticks} time {dt:.3f}s us/tick {dt / res.ticks * 1e6:.1f}")
b = pkg.Batch(circ, 1)
lib.L.ngbProfile(1, 8)
res = b.tran(4096, wave["save_eq"][:1])
ms = (ctypes.c_double * 8)()
n = lib.L.ngbProfileStages(ms)
names = ["", "small loads", "bsim4_load", "assemble", "lu", "bsim4_lte", "control", ""]
print("stages (us per sampled step, %d steps): " % n + "  ".join(f"{names[k]} {ms[k] / max(n, 1) * 1e3:.1f}" for k in range(1, 7)) + f"  sum {sum(ms) / max(n, 1) * 1e3:.1f}")
