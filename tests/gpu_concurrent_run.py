"""The 4096-sample Monte-Carlo transient of gpu_profile_run.py as K batches driven from K host threads
(each thread has its own launch stream, so the batches' kernels overlap on the device):
   python gpu_concurrent_run.py [samples] [threads]"""
import sys, time, threading
import numpy as np
from parity_util import GOLDEN, ngt, pkg, first_pattern
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lib = pkg.library()
flat = ngt.read(f"{GOLDEN}/ro17k.flat.ngt"); trace = ngt.read(f"{GOLDEN}/ro17k.trace.ngt.gz"); wave = ngt.read(f"{GOLDEN}/ro17k.wave.ngt")
dv = pkg.mc.draw_delvto(S, 34, seed=5)
inst = pkg.mc.bsim4_inst_with_delvto(lib, flat, dv)          # [NI][ninst][S]
parts = np.array_split(np.arange(S), K)
res = [None] * K
batches = [None] * K
def setup(k):
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=first_pattern(trace))
    b = pkg.Batch(circ, len(parts[k]))
    b.put("b4.inst", np.ascontiguousarray(inst[:, :, parts[k]]))
    batches[k] = b
def work(k):
    res[k] = batches[k].tran(1024, wave["save_eq"][:1])
for phase in (setup, work):
    t0 = time.time()
    th = [threading.Thread(target=phase, args=(k,)) for k in range(K)]
    for t in th: t.start()
    for t in th: t.join()
    dt = time.time() - t0
iters = sum(int(r.numiter.astype(np.int64).sum()) for r in res)
ticks = max(r.ticks for r in res)
print(f"S={S} threads {K} ticks(max) {ticks} time {dt:.3f}s us/tick-equivalent {dt / ticks * 1e6:.1f} iters {iters} evals/s {34 * iters / dt:.3e}")
