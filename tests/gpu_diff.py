"""dev aid: where does a GPU transient first differ from the reference fixture?  python gpu_diff.py NAME"""
import sys
import numpy as np
from parity_util import GOLDEN, ngt, pkg
from test_tran_parity import _run
name = sys.argv[1]
res, t, v, wave = _run(pkg.library(), name, S=1)
print(name, "ours acc/rej/iter", int(res.accepted[0]), int(res.rejected[0]), int(res.numiter[0]), "ref", wave["stats"][:3])
n = min(int(res.npoints[0]), len(wave["time"]))
dt = np.nonzero(t[0, :n] != wave["time"][:n])[0]
dv = np.nonzero((v[0, :n, :] != wave["values"][:n]).any(axis=1))[0]
print("first time diff", dt[:3], "first value diff", dv[:3])
if len(dv):
    k = dv[0]
    print("t", wave["time"][k], "ours", v[0, k], "ref", wave["values"][k], "diff", v[0, k] - wave["values"][k])
rng = np.max(np.abs(wave["values"]), axis=0)
print("max rel err (of range)", np.max(np.abs(v[0, :n, :] - wave["values"][:n]), axis=0) / rng)
