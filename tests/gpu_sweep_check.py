"""GPU helper (not a test): run one rank's share of the config-5 sweep and list the points that did not
complete, with their error codes, into gpurun_out/sweep_check.json"""
import json
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from parity_util import ngt, pkg, run_patterns, GOLDEN   # noqa: E402
import bench                                              # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
lib = pkg.library()
flat = ngt.read(f"{GOLDEN}/mix.flat.ngt"); trace = ngt.read(f"{GOLDEN}/mix.trace.ngt.gz"); wave = ngt.read(f"{GOLDEN}/mix.wave.ngt")
circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=run_patterns(trace))
b = pkg.Batch(circ, S)
vt, rt = bench.sweep_points(0, 1, S)
pkg.sweep.apply(b, flat, dc={"vdd": vt}, res={"r1": rt})
res = b.tran(320, wave["save_eq"][:5])
err = b.get("ctl.err")
bad = [int(i) for i in np.nonzero((res.accepted < 100) | (err != 0))[0]]
out = {"S": S, "bad": [{"i": i, "vdd": vt[i], "r1": rt[i], "err": int(err[i]), "accepted": int(res.accepted[i]),
                        "rejected": int(res.rejected[i]), "numiter": int(res.numiter[i])} for i in bad],
       "accepted_hist": np.bincount(res.accepted.astype(np.int64)).nonzero()[0].tolist()}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/sweep_check.json", "w"), indent=1)
print(json.dumps(out)[:2000])
