mkdir -p gpurun_out
( python tools/dbg/mcpar.py hostsim 32 host 2>&1 | tail -1 ) &
python tools/dbg/mcpar.py cuda 4096 gpu_default 2>&1 | tail -1
NGB_HOST_PIVOT=1 python tools/dbg/mcpar.py cuda 4096 gpu_hostpivot 2>&1 | tail -1
NGB_ASM_TILED=0 python tools/dbg/mcpar.py cuda 4096 gpu_notiled 2>&1 | tail -1
NGB_LTE_FLAT=0 python tools/dbg/mcpar.py cuda 4096 gpu_noflat 2>&1 | tail -1
NGB_L2_PERSIST=0 python tools/dbg/mcpar.py cuda 4096 gpu_nol2 2>&1 | tail -1
wait
python - <<'PY'
import numpy as np
ref = np.load("gpurun_out/mc_host.npy"); tref = np.load("gpurun_out/mct_host.npy")
for tag in ("gpu_default", "gpu_hostpivot", "gpu_notiled", "gpu_noflat", "gpu_nol2"):
    v = np.load(f"gpurun_out/mc_{tag}.npy"); t = np.load(f"gpurun_out/mct_{tag}.npy")
    d = np.abs(v - ref).max(axis=(1, 2)); dt = np.abs(t - tref).max(axis=1)
    print(tag, "samples differing:", np.nonzero(d)[0].tolist(), "max", d.max(), "t", dt.max())
PY
