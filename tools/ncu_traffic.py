#!/usr/bin/env python3
"""profiles/r02_traffic.json from the round's `ncu --set full` captures (gpurun_out/r02_*.ncu-rep): DRAM bytes and duration
per launch of the three kernels of a Newton step; bench.py scales the BSIM4 load's figure into `roofline.traffic`.
   python tools/ncu_traffic.py [dir with r02_b4load.ncu-rep r02_lu.ncu-rep r02_asm.ncu-rep]"""
import csv, json, os, subprocess, sys
d = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u, v = rows[0], rows[1], rows[-1]
    return {k: (x, un) for k, un, x in zip(h, u, v)}


def num(m, key):
    x, un = m[key]
    x = float(x.replace(",", ""))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3}.get(un, 1)
    return x * scale


out = {"source": "ncu --set full --clock-control none, one launch each; ngb_k_bsim4_load: launch 15 000 of the default bench's transient "
                 "(tools/gpu_r2_final.sh: 4096 samples, per-sample parameter rows, every sample at its own phase); LU and assembly: launch 30 of the "
                 "lock-step batch (tests/gpu_profile_run.py 4096, tools/gpu_r2_profiles.sh); summaries in profiles/r02_*_ncu_details.txt"}
for key, rep, units in (("ngb_k_bsim4_load", "r02_b4load.ncu-rep", 139264), ("ngb_k_lu_packed", "r02_lu.ncu-rep", 4096), ("ngb_k_assemble", "r02_asm.ncu-rep", 4096)):
    p = os.path.join(d, rep)
    if not os.path.exists(p):
        continue
    m = raw(p)
    out[key] = {"dram_bytes_read": num(m, "dram__bytes_read.sum"), "dram_bytes_write": num(m, "dram__bytes_write.sum"),
                "units_per_launch": units, "duration_us": num(m, "gpu__time_duration.sum"),
                "kernel": m.get("Kernel Name", ("", ""))[0]}
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
