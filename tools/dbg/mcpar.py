import sys, os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "tests"))
from parity_util import *
which, K, tag = sys.argv[1], int(sys.argv[2]), sys.argv[3]
lib = pkg.Library(HOSTSIM) if which == "hostsim" else pkg.Library()
S = 4096; rank = 0
flat = ngt.read(f"{GOLDEN}/ro17k.flat.ngt"); flat["tran/tstop"] = np.array([pkg.mc.spice_number("150ns")])
trace = ngt.read(f"{GOLDEN}/ro17k.trace.ngt.gz"); wave = ngt.read(f"{GOLDEN}/ro17k.wave.ngt")
ninst = int(flat["b4/ninst"][0])
circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=run_patterns(trace))
dv_raw = pkg.mc.draw_delvto(S, ninst, sigma=0.015, seed=1000 + rank); dv = pkg.mc.delvto_as_parsed(dv_raw)
b4t = ngt.read(f"{GOLDEN}/b4temp.tables.ngt.gz")
raw = {"model": b4t["ro17k/b4t/model"], "inst": b4t["ro17k/b4t/inst"], "inst_model": b4t["ro17k/b4t/inst_model"], "temp": b4t["ro17k/b4t/temp"][0, 0], "vt0": b4t["ro17k/opt/vt0"][0]}
tox_raw = 1.4e-9 * (1.0 + 0.03 * np.random.default_rng(5000 + rank).normal(size=S))
order = np.argsort(tox_raw, kind="stable"); tox_raw, dv = tox_raw[order], dv[order]
tox = np.array([pkg.mc.spice_number(f"{x:.17g}") for x in tox_raw])
inst_host, prow_t, mtab_all, ptab_all = pkg.mc.bsim4_with_toxe(lib, raw, tox[:K], dv[:K])
b = pkg.Batch(circ, K)
b.put("b4.inst", inst_host); b.set_bsim4_rows(prow_t, mtab_all, ptab_all)
res = b.tran(6144, wave["save_eq"][:1])
t, v = res.waves()
print(tag, "repivots", res.repivots, "ticks", res.ticks)
np.save(f"gpurun_out/mc_{tag}.npy", v[:32]); np.save(f"gpurun_out/mct_{tag}.npy", t[:32])
