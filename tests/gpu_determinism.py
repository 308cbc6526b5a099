"""development aid: repeatability of the GPU transient (run on the GPU box)"""
import sys, hashlib
import numpy as np
from parity_util import GOLDEN, ngt, pkg, first_pattern
lib = pkg.library()
name = sys.argv[1] if len(sys.argv) > 1 else "ro17k"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
flat = ngt.read(f"{GOLDEN}/{name}.flat.ngt"); trace = ngt.read(f"{GOLDEN}/{name}.trace.ngt.gz"); wave = ngt.read(f"{GOLDEN}/{name}.wave.ngt")
circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=first_pattern(trace))
for r in range(reps):
    b = pkg.Batch(circ, 3)
    res = b.tran(2048, wave["save_eq"])
    t, v = res.waves(); n = int(res.npoints[0])
    m = min(n, len(wave["time"]))
    err = np.abs(v[0, :m, :] - wave["values"][:m]).max(axis=0) / np.abs(wave["values"]).max(axis=0)
    first = np.where(np.abs(v[0, :m, 0] - wave["values"][:m, 0]) > 0)[0]
    print(r, hashlib.md5(v[:, :n].tobytes()).hexdigest()[:10], hashlib.md5(t[:, :n].tobytes()).hexdigest()[:10], n, err, "first diff idx", first[:3],
          "s0==s2", np.array_equal(v[0, :n], v[2, :n]))
