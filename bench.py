#!/usr/bin/env python
"""bench.py -- throughput of the Newton hot path on B200.

Default workload (BASELINE.json configs[2], the configuration the multi-GPU metric is quoted on,
and the largest that shards): Monte-Carlo transient of the 17-stage BSIM4 ring oscillator
(examples/mos/ro_17_4.cir cards, `.tran .1ns 150ns uic`, KLU), 4096 samples PER GPU with
per-instance Vth mismatch (delvto ~ N(0, 15 mV), seeded), every sample on its own adaptive time
axis.  One "step" = one complete transient of the batch.  `--workload ro101` runs configs[1]
(single 101-stage circuit).

  value  = BSIM4 instance-evals/s over the whole job (34 instances x Newton iterations of every
           sample / time), inputs resident in HBM, timed with CUDA events on the launch stream
  e2e    = the same job through the public API with HOST buffers: per-sample parameter table
           copied from pinned memory and result waveforms copied back inside the timed region
  roofline = dominant kernel (bsim4_load): algorithmic bytes per launch / CUDA-event duration
  cpu_baseline = oracle/_ref/ngspice (the reference compiled here) on the host cores, bounded sample

`--impl reference` times the reference CPU implementation of the same workload instead.
"""
import argparse
import ctypes
import importlib
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")

B4_BYTES_PER_EVAL = 2000.0          # algorithmic bytes per BSIM4 instance-eval
B4_FLOP_PER_EVAL = 2957.0           # pinned with gcov on the reference (oracle/flops_gcov.sh, tools/flops_gcov.py)
# oxide-thickness levels of the Monte-Carlo workload: 8 equal-probability bins of N(1.4 nm, 3 %)
# (tests/golden/make_golden.py: tox_levels; the BSIM4temp results per level are in ro17tox.tables.ngt)
TOX_Z = [-1.5341205443525463, -0.8871465590188759, -0.4887764111146695, -0.15731068461017067,
         0.15731068461017067, 0.4887764111146695, 0.8871465590188759, 1.5341205443525463]
TOX_LEVELS = [1.4e-9 * (1.0 + 0.03 * z) for z in TOX_Z]           # provisional source-level FP op count per eval


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    def __init__(self, dev):
        self.dev, self.rows, self.stop = dev, [], False
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        """NVML in-process (one cheap query per 0.5 s); spawning nvidia-smi every 200 ms was measured to
        slow the timed region by ~10 %, so the subprocess form is only the fallback, at 2 s intervals"""
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.dev)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            bits = [("hw_slowdown", pynvml.nvmlClocksThrottleReasonHwSlowdown),
                    ("hw_thermal_slowdown", pynvml.nvmlClocksThrottleReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", pynvml.nvmlClocksThrottleReasonSwThermalSlowdown),
                    ("sw_power_cap", pynvml.nvmlClocksThrottleReasonSwPowerCap)]
            while not self.stop:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if (r & b) else "Not Active" for _, b in bits])
                time.sleep(0.5)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.dev}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(2.0)

    def __enter__(self):
        self.th.start(); return self

    def __exit__(self, *a):
        self.stop = True; self.th.join(timeout=3)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons}


def traffic_from_profiles(units):
    """DRAM bytes per ngb_k_bsim4_load launch from the committed `ncu --set full` capture of this build
    (profiles/r02_traffic.json, written by tools/ncu_traffic.py), scaled to this run's evaluations per launch"""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))["ngb_k_bsim4_load"]
        return (d["dram_bytes_read"] + d["dram_bytes_write"]) * units / d["units_per_launch"]
    except Exception:
        return None


# ------------------------------------------------------------------ reference CPU arm
def read_rawfile(path, names):
    """binary rawfile of the reference (src/frontend/rawfile.c): returns (time [points], {name: values [points]})"""
    data = open(path, "rb").read()
    i = data.index(b"Binary:\n")
    head = data[:i].decode(errors="replace").splitlines()
    nvar = int([ln for ln in head if ln.startswith("No. Variables")][0].split(":")[1])
    npts = int([ln for ln in head if ln.startswith("No. Points")][0].split(":")[1])
    k = head.index("Variables:")
    vars_ = [ln.split()[1].lower() for ln in head[k + 1:k + 1 + nvar]]
    a = np.frombuffer(data[i + 8:], dtype=np.float64)[:npts * nvar].reshape(npts, nvar)
    return a[:, 0].copy(), {n: a[:, vars_.index(n)].copy() for n in names}


def cpu_reference_run(workload, nproc, samples_per_proc, seed0=1000, draws=None, keep_raw=False, inst_names=None):
    """Runs the REFERENCE ngspice (oracle/_ref/ngspice, stock code path) on the host cores:
    `nproc` processes in parallel, each simulating `samples_per_proc` mismatch samples
    sequentially.  `draws` = [(delvto [ninst], toxe)] gives the samples explicitly (the bench passes the draws of
    its own GPU samples 0..k-1 so that the two arms simulate the same circuits; delvto then is in the order of
    `inst_names`, the instance order of the flattened tables, and is matched to the netlist lines by name); keep_raw
    leaves the rawfiles in place.  Returns (evals/s, samples/s, wall seconds, description, rawfile paths)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ngspice")
    if not os.path.exists(exe):
        return None
    # Monte-Carlo workload: the kicked oscillator (`.ic` alternates the stage outputs, deterministic start) over the
    # example's full 150 ns -- the circuit whose BSIM4temp tables per oxide-thickness level are in ro17tox.tables.ngt
    base = open(os.path.join(GOLDEN, "netlists", ("ro101" if workload == "ro101" else "ro17k") + ".cir")).read()
    base = base.replace(".tran .1ns 20ns uic", ".tran .1ns 150ns uic")
    ninst = 202 if workload == "ro101" else 34
    tmp = tempfile.mkdtemp(prefix="ngb_cpu_")
    jobs = []
    for p in range(nproc):
        files = []
        for k in range(samples_per_proc):
            if draws is not None:
                dv, tox = draws[p * samples_per_proc + k]
            else:
                rng = np.random.default_rng(seed0 + p * samples_per_proc + k)
                dv = rng.normal(0.0, 0.015, size=ninst) if workload != "ro101" else np.zeros(ninst)
                tox = 1.4e-9 * (1.0 + 0.03 * float(rng.normal())) if workload != "ro101" else None      # continuous, like the GPU arm's default
            lines, i = [], 0
            for ln in base.splitlines():
                if ln[:2].lower() in ("mp", "mn") and " l=" in ln:
                    j = inst_names.index(ln.split()[0].lower()) if inst_names is not None else i
                    ln = ln + f" delvto={dv[j]:.17g}"; i += 1
                lines.append(ln)
            f = os.path.join(tmp, f"s{p}_{k}.cir")
            text = "\n".join(lines).replace(".option xmu=0.49 klu", ".option xmu=0.49 klu acct")
            if tox is not None:
                text = re.sub(r"toxe\s*=\s*1\.4e-0*9", f"toxe    = {tox:.17g}", text)
            open(f, "w").write(text + "\n")
            files.append(f)
        jobs.append(files)
    t0 = time.time()
    procs = []
    for files in jobs:
        # the rawfile goes to a scratch file: ngspice unlinks and recreates its -r target, so it must never be /dev/null
        cmd = " ; ".join(f"{exe} -b -r {f}.raw {f} > {f}.log 2>&1" + ("" if keep_raw else f" ; rm -f {f}.raw") for f in files)
        procs.append(subprocess.Popen(["bash", "-c", cmd]))
    for pr in procs:
        pr.wait()
    wall = time.time() - t0
    iters = 0
    for files in jobs:
        for f in files:
            try:
                for ln in open(f + ".log", errors="replace"):
                    if ln.startswith("Total iterations"):
                        iters += int(ln.split("=")[-1].strip().split()[0])
            except Exception:
                pass
    nsamp = nproc * samples_per_proc
    if iters == 0:                       # acct line not found: fall back to the recorded iteration count
        iters = nsamp * (24622 if workload == "ro101" else 24291)
    evals = ninst * iters
    raws = [f + ".raw" for files in jobs for f in files]
    return evals / wall, nsamp / wall, wall, f"{nsamp} full transients ({samples_per_proc} per process x {nproc} processes)", raws


def measure_when(t, v, kind, count, val, td):
    """com_measure_when for one real vector against a constant (src/frontend/com_measure2.c:455-660): the checker's
    restatement, applied to the REFERENCE's waveforms"""
    first = 0; section = -1; rise = fall = 0
    pv = pt = 0.0
    for scale, value in zip(t, v):
        if scale < td:
            continue
        if first == 1:
            rise = fall = 0
            if value < val:
                section = 0
                if pv >= val:
                    fall = 1
            else:
                section = 1
                if pv < val:
                    rise = 1
        if first > 1:
            if section == 0 and value >= val:
                section = 1; rise += 1
            elif section == 1 and value <= val:
                section = 0; fall += 1
            have = rise if kind == 0 else (fall if kind == 1 else rise + fall)
            if have == count:
                return pt + (val - pv) * (scale - pt) / (value - pv)
        first += 1
        pv, pt = value, scale
    return float("nan")


def meas_check(raws, tdiff_gpu, out_name, clauses):
    worst = 0.0; same = True
    for k, raw in enumerate(raws):
        t, vecs = read_rawfile(raw, [out_name])
        m = [measure_when(t, vecs[out_name], c[1], c[2], c[3], c[4]) for c in clauses]
        ref = m[1] - m[0]
        got = float(tdiff_gpu[k])
        if not (ref == got or (np.isnan(ref) and np.isnan(got))):
            same = False
            worst = max(worst, abs(got - ref) / abs(ref) if ref == ref and ref != 0 else float("inf"))
    return {"meas": "tdiff = TRIG v(out) VAL=0.5 RISE=10 TARG RISE=20 (ro_17_4.cir:54), device-side", "meas_identical": same,
            "meas_max_rel_err": worst}


def parity_check(raws, t_gpu, v_gpu, npoints, out_name):
    """the reference's rawfiles against the GPU waveforms of the same draws: identical number of accepted time points
    and, per point, |t - t_ref| and |v - v_ref| within 1e-9 * max(|ref|, vntol-scale);
    vntol = 1e-6 V, and 1e-18 s stands in for it on the time axis"""
    worst_v = worst_t = 0.0
    same = True
    bit = True
    for s, raw in enumerate(raws):
        t_ref, vals = read_rawfile(raw, [out_name])
        v_ref = vals[out_name]
        n = int(npoints[s])
        if n != len(t_ref):
            same = False
            continue
        tg, vg = t_gpu[s, :n], v_gpu[s, :n, 0]
        worst_t = max(worst_t, float(np.max(np.abs(tg - t_ref) / np.maximum(np.abs(t_ref), 1e-18))))
        worst_v = max(worst_v, float(np.max(np.abs(vg - v_ref) / np.maximum(np.abs(v_ref), 1e-6))))
        bit = bit and np.array_equal(tg, t_ref) and np.array_equal(vg, v_ref)
    return {"samples": len(raws), "accepted_identical": same, "max_rel_err": worst_v, "max_rel_err_time": worst_t,
            "bit_identical": bool(same and bit), "tolerance": 1e-9, "vector": out_name,
            "ok": bool(same and worst_v <= 1e-9 and worst_t <= 1e-9)}


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    if args.workload == "array":
        r = array_cpu_baseline(cores)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ngspice not built"}))
            return
        print(json.dumps({
            "impl": "reference", "metric": "BSIM4 instance-evals/s", "value": r[0], "unit": "evals/s", "n_gpus": args.gpus,
            "steps": 1, "warmup": 0, "ms_per_step": r[1] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "flat BSIM4 inverter array with RC links (config 4), CKTload only"},
            "cpu_baseline": {"value": r[0], "unit": "evals/s", "cores": cores, "kind": "reference", "sample": r[2]},
            "e2e": {"value": r[0], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    vals = []
    steps = max(1, min(args.steps, 3))
    t_all = []
    for _ in range(steps):
        r = cpu_reference_run(args.workload, cores, 1)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ngspice not built"}))
            return
        vals.append(r); t_all.append(r[2])
    ev = float(np.mean([v[0] for v in vals])); sps = float(np.mean([v[1] for v in vals]))
    line = {
        "impl": "reference", "metric": "BSIM4 instance-evals/s", "value": ev, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": 0, "ms_per_step": float(np.mean(t_all)) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args), "mc_samples_per_s": sps,
        "cpu_baseline": {"value": ev, "unit": "evals/s", "cores": cores, "kind": "reference", "sample": vals[0][3]},
        "e2e": {"value": ev, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def workload_config(args):
    if args.workload == "ro101":
        return {"workload": "101-stage BSIM4 ring oscillator transient (.tran .1ns 150ns uic), single circuit, KLU pivot order",
                "samples_per_gpu": 1, "bsim4_instances": 202, "unknowns": 911,
                "l2": "working set smaller than L2 by nature (single circuit); no flush"}
    return {"workload": "Monte Carlo transient, 17-stage BSIM4 ring oscillator (ro_17_4.cir cards, version 4.8.3), "
                        ".tran .1ns 150ns uic from alternating `.ic` stage voltages, per-instance delvto mismatch sigma 15 mV and per-sample toxe "
                        + ("~ N(1.4 nm, 3 %), continuous (rows by the library's BSIM4temp)" if getattr(args, "tox", "continuous") == "continuous" else "from 8 levels of N(1.4 nm, 3 %)"),
            "samples_per_gpu": args.samples, "bsim4_instances": 34, "unknowns": 155,
            "layout": ("per-sample parameter rows" + ("" if os.environ.get("NGB_B4_OVERLAY") == "0" else " read as an overlay (only the columns that differ between samples per lane)") + ", " + ("draws unsorted" if os.environ.get("NGB_BENCH_UNSORTED") else "samples laid out by toxe (same draws)")) if getattr(args, "tox", "continuous") == "continuous"
                      else ("draws unsorted" if os.environ.get("NGB_BENCH_UNSORTED") else "samples laid out level by level (same draws)"),
            "l2": "inputs larger than L2: per-step working set (parameters+states+stamps+matrices) ~%.0f MB" %
                  (args.samples * 34 * (51 + 4 * 29 + 38 + 52) * 8 / 1e6 + args.samples * 904 * 8 / 1e6)}


# ------------------------------------------------------------------ our arm
def bench_ours(args):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("ngspice-sf-mirror_b200"); ngt = pkg.ngt
    first_pattern, run_patterns = ngt.first_pattern, ngt.run_patterns

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    lib = pkg.library()
    assert lib.backend == "cuda-sm_100a"
    lib.check(lib.L.ngbInit(local), "ngbInit")
    stream = torch.cuda.Stream(device=local)
    lib.check(lib.L.ngbSetStream(ctypes.c_void_p(stream.cuda_stream)), "ngbSetStream")

    name = "ro101" if args.workload == "ro101" else "ro17k"
    flat = ngt.read(f"{GOLDEN}/{name}.flat.ngt")
    if name == "ro17k":
        flat["tran/tstop"] = np.array([pkg.mc.spice_number("150ns")])        # the fixture was recorded over 20 ns; same step limits (tmax = tstep)
    trace = ngt.read(f"{GOLDEN}/{name}.trace.ngt.gz")
    wave = ngt.read(f"{GOLDEN}/{name}.wave.ngt")
    ninst = int(flat["b4/ninst"][0])
    S = 1 if args.workload == "ro101" else args.samples
    if args.scaling == "strong" and args.workload != "ro101":
        S = max(1, args.samples // world)              # BASELINE config 3 as written: 4096 samples in total, 4096 / N per GPU
    # the pivoting factors the reference computed in this run (two for a UIC transient; identical for this circuit,
    # checked here so that the benchmarked path is the one tests/test_tran_parity.py pins)
    pats = run_patterns(trace)
    assert all(np.array_equal(p[k], first_pattern(trace)[k]) for p in pats for k in p)
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=pats)
    batch = pkg.Batch(circ, S, device=local)
    save_eq = wave["save_eq"][:1]                     # v(out)
    max_points = 6144

    # per-sample mismatch parameters, built on the host like the reference's MC front end would
    if args.workload == "ro101":
        inst_host = np.repeat(np.asarray(flat["b4/inst"])[:, :, None], S, axis=2)
    else:
        dv_raw = pkg.mc.draw_delvto(S, ninst, sigma=0.015, seed=1000 + rank)
        dv = pkg.mc.delvto_as_parsed(dv_raw)          # what the reference's number parser makes of the netlist text
        if args.tox == "continuous":
            # every sample its own oxide thickness ~ N(1.4 nm, 3 %): the model / bin / instance rows of each sample come from
            # the library's own BSIM4temp (csrc/ngb_b4temp.c) on the nominal card; per-sample table rows
            b4t = ngt.read(f"{GOLDEN}/b4temp.tables.ngt.gz")
            raw = {"model": b4t["ro17k/b4t/model"], "inst": b4t["ro17k/b4t/inst"], "inst_model": b4t["ro17k/b4t/inst_model"],
                   "temp": b4t["ro17k/b4t/temp"][0, 0], "vt0": b4t["ro17k/opt/vt0"][0]}
            sigma = float(os.environ.get("NGB_BENCH_TOX_SIGMA", "0.03"))       # 1e-9: distinct rows, identical dynamics (measures the row access alone)
            tox_raw = 1.4e-9 * (1.0 + sigma * np.random.default_rng(5000 + rank).normal(size=S))
            if not os.environ.get("NGB_BENCH_UNSORTED"):
                # same draws, laid out by oxide thickness: the 32 samples a warp evaluates then oscillate at nearly the same
                # period and stay in step for longer (the load kernel's cost is the divergence between them)
                order = np.argsort(tox_raw, kind="stable")
                tox_raw, dv, dv_raw = tox_raw[order], dv[order], dv_raw[order]
            tox = np.array([pkg.mc.spice_number(f"{x:.17g}") for x in tox_raw])     # what the reference's parser makes of the card text
            inst_host, prow_t, mtab_all, ptab_all = pkg.mc.bsim4_with_toxe(lib, raw, tox, dv)
            tox_of = lambda p: float(tox_raw[p])
        else:
            tox_tables = ngt.read(f"{GOLDEN}/ro17tox.tables.ngt")
            level = np.random.default_rng(5000 + rank).integers(0, len(tox_tables["levels"]), size=S)
            if not os.environ.get("NGB_BENCH_UNSORTED"):
                # same draws, laid out level by level: a warp's 32 samples then share their parameter rows
                order = pkg.mc.group_by_level(level)
                level, dv, dv_raw = level[order], dv[order], dv_raw[order]
            inst_host, prow_t, mtab_all, ptab_all = pkg.mc.bsim4_with_tox_levels(lib, flat, tox_tables, level, dv)
            tox_of = lambda p: float(tox_tables["levels"][level[p]])
        batch.set_bsim4_rows(prow_t, mtab_all, ptab_all)
    pinned = torch.empty(inst_host.shape, dtype=torch.float64).pin_memory()
    pinned.numpy()[...] = inst_host
    out_t = torch.empty((S, max_points), dtype=torch.float64).pin_memory()
    out_v = torch.empty((S, max_points, 1), dtype=torch.float64).pin_memory()
    h2d_bytes = pinned.numel() * 8
    if args.workload != "ro101":
        h2d_bytes += prow_t.nbytes + mtab_all.nbytes + ptab_all.nbytes
    # what the example asks of every run (examples/mos/ro_17_4.cir:54): `meas tran tdiff TRIG v(18) VAL=0.5 RISE=10 TARG
    # v(18) VAL=0.5 RISE=20`.  The end-to-end pass evaluates it on the device while the points are produced
    # (ngbTranSetMeasures) and reads back two doubles per sample; no waveform is stored or copied in that pass
    meas_clauses = [(int(save_eq[0]), 0, 10, 0.5, 0.0), (int(save_eq[0]), 0, 20, 0.5, 0.0)]
    meas_out = torch.empty((2, S), dtype=torch.float64).pin_memory()
    d2h_bytes = meas_out.numel() * 8
    tdiff = [None]

    def step(e2e):
        if e2e:
            batch.put("b4.inst", pinned.numpy())
            if args.workload != "ro101":
                batch.set_bsim4_rows(prow_t, mtab_all, ptab_all)
            batch.set_measures(meas_clauses)
            res = batch.tran(0, [])
            lib.check(lib.L.ngbTranMeasures(batch.h, ctypes.cast(meas_out.data_ptr(), ctypes.POINTER(ctypes.c_double))), "ngbTranMeasures")
            tdiff[0] = meas_out.numpy()[1] - meas_out.numpy()[0]
            batch.set_measures([])
            return res
        return batch.tran(max_points, save_eq)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batch.put("b4.inst", pinned.numpy())
    for _ in range(args.warmup):
        step(False)

    def timed(e2e, profile):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        iters = 0; ticks = 0; failed = 0; cut = 0
        if profile:
            lib.L.ngbProfile(1, 16)
        barrier()
        n0 = lib.launch_count()
        t0 = time.time()
        for k in range(args.steps):
            evs[k][0].record(stream)
            res = step(e2e)
            evs[k][1].record(stream)
            iters += int(res.numiter.astype(np.int64).sum()); ticks += res.ticks
            failed += int((res.err != 0).sum()); cut += int((res.npoints > max_points).sum())
        barrier()
        wall = time.time() - t0
        ms = sum(a.elapsed_time(b) for a, b in evs)
        launches = lib.launch_count() - n0
        prof = None
        if profile:
            msum, cnt = ctypes.c_double(), ctypes.c_long()
            lib.L.ngbProfileRead(ctypes.byref(msum), ctypes.byref(cnt))
            lib.L.ngbProfile(0, 1)
            prof = (msum.value, cnt.value)
        return ms, wall, iters, ticks, launches, prof, res, failed, cut

    with ClockSampler(local) as clk:
        ms, wall, iters, ticks, launches, prof, res, failed, cut = timed(False, True)
    clocks = clk.summary()
    # the waveforms of the device-timed pass are what the parity check compares with the reference's rawfiles (read
    # back here, outside both timed regions)
    lib.check(lib.L.ngbTranWaves(batch.h, ctypes.cast(out_t.data_ptr(), ctypes.POINTER(ctypes.c_double)),
                                 ctypes.cast(out_v.data_ptr(), ctypes.POINTER(ctypes.c_double))), "ngbTranWaves")
    npoints_wave = res.npoints.copy()
    ms_e2e, wall_e2e, iters_e2e, _, _, _, res_e2e, failed_e2e, _ = timed(True, False)

    # max over ranks of the device time; totals over ranks
    tt = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    # samples that ended with an error code are not counted as simulated (bench.py:283)
    ww = torch.tensor([float(iters), float(iters_e2e), float(S * args.steps - failed), float(failed), float(cut),
                       float(S * args.steps - failed_e2e)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ww, op=dist.ReduceOp.SUM)
        # the measurement results gathered over NCCL/NVLink (the only collective on this path, off the timed region)
        gathered = None
        try:
            wv = torch.from_numpy(np.ascontiguousarray(tdiff[0])).to("cuda")
            lst = [torch.empty_like(wv) for _ in range(world)] if rank == 0 else None
            dist.gather(wv, lst, dst=0)
            gathered = True if rank != 0 else int(sum(int(torch.isfinite(x).sum()) for x in lst))   # samples with a period measured
        except Exception as e:                         # report but do not fail the bench
            gathered = str(e)
    ms_max, ms_e2e_max = tt.tolist()
    iters_tot, iters_e2e_tot, samples_tot, failed_tot, cut_tot, samples_e2e_tot = ww.tolist()
    evals = ninst * iters_tot
    value = evals / (ms_max * 1e-3)
    e2e_val = ninst * iters_e2e_tot / (ms_e2e_max * 1e-3)

    if rank == 0:
        hbm_peak, which = peaks()
        units = ninst * S                                  # evals one bsim4_load launch processes (all samples active)
        k_ms = prof[0] / max(prof[1], 1)
        achieved = units * B4_BYTES_PER_EVAL / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        # CPU arm: the reference simulates the SAME draws as this rank's samples 0..cores-1 and its rawfiles are the
        # parity check of the benchmarked run (e2e pass: the waveforms already sit in the pinned host buffers)
        cores = os.cpu_count() or 1
        k = min(cores, S)
        if args.workload == "ro101":
            draws = [(np.zeros(ninst), None)] * k
        else:
            draws = [(dv_raw[p], tox_of(p)) for p in range(k)]
        cpu = cpu_reference_run(args.workload, k, 1, draws=draws, keep_raw=True,
                                inst_names=[n.lower() for n in pkg.mc.instance_names(flat)])
        parity = None
        if cpu is not None:
            names = bytes(np.asarray(flat["node/names_bytes"]).astype(np.uint8)).decode().split("\n")
            out_name = "v(%s)" % [ln.split()[1] for ln in names if ln and int(ln.split()[0]) == int(save_eq[0])][0].lower()
            parity = parity_check(cpu[4], out_t.numpy(), out_v.numpy(), npoints_wave, out_name)
            # the measurement of the end-to-end pass against com_measure_when applied to the reference's own waveforms
            parity.update(meas_check(cpu[4], tdiff[0], out_name, meas_clauses))
            parity["ok"] = bool(parity["ok"] and parity["meas_identical"])
            for f in cpu[4]:
                try:
                    os.remove(f)
                except OSError:
                    pass
        fp = (ctypes.c_double * 3)()
        fp64_peak = None
        if lib.L.ngbMeasureFp64Peak(fp) == 0:
            fp64_peak = {"dfma_tflops": fp[0] / 1e12, "dadd_dmul_tflops": fp[1] / 1e12}
        line = {
            "metric": "BSIM4 instance-evals/s", "value": value, "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args), samples_per_gpu=S),
            "mc_samples_per_s": samples_tot / (ms_max * 1e-3),
            "samples_failed": int(failed_tot), "samples_truncated": int(cut_tot),
            "parity_check": parity,
            "newton_steps_per_transient": ticks / args.steps,
            "e2e": {"value": e2e_val, "unit": "evals/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                    "mc_samples_per_s": samples_e2e_tot / (ms_e2e_max * 1e-3)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak if hbm_peak else None, "traffic": traffic_from_profiles(units),
                         "kernel": "ngb_k_bsim4_load", "avg_launch_ms": k_ms, "timed_launches": prof[1],
                         "units_per_launch": units, "bytes_per_unit": B4_BYTES_PER_EVAL, "peak_source": which,
                         "fp64_tflops_algorithmic": units * B4_FLOP_PER_EVAL / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0,
                         # against the FP64 pipe of this GPU, measured in this run (ngbMeasureFp64Peak): one flop per
                         # DADD/DMUL instruction is what code compiled without contraction can reach, DFMA counts two
                         "fp64_peak": fp64_peak,
                         "fp64_frac": (units * B4_FLOP_PER_EVAL / (k_ms * 1e-3) / 1e12 / fp64_peak["dfma_tflops"]) if (fp64_peak and k_ms > 0) else None,
                         "fp64_frac_unfused": (units * B4_FLOP_PER_EVAL / (k_ms * 1e-3) / 1e12 / fp64_peak["dadd_dmul_tflops"]) if (fp64_peak and k_ms > 0) else None,
                         "kernel_share_of_step": (k_ms * ticks / args.steps) / (ms_max / args.steps) if ms_max else None},
        }
        if cpu is not None:
            line["cpu_baseline"] = {"value": cpu[0], "unit": "evals/s", "cores": os.cpu_count() or 1, "kind": "reference",
                                    "sample": cpu[3], "mc_samples_per_s": cpu[1]}
        if world > 1:
            line["nccl_gather"] = gathered
        print(json.dumps(line))
        if parity is not None and not parity["ok"]:
            raise SystemExit("bench.py: parity check against the reference failed: %s" % json.dumps(parity))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ config 4: one large flat circuit
def array_cpu_baseline(nproc, nx=32, ny=32):
    """reference ngspice on an nx-by-ny instance of the same generator, one process per core; evals/s
    from its own STATloadTime ("Matrix load time" of `.option acct`): load only, like the GPU figure"""
    exe = os.path.join(ROOT, "oracle", "_ref", "ngspice")
    if not os.path.exists(exe):
        return None
    synth = importlib.import_module("ngspice-sf-mirror_b200.synth")
    src = open(os.path.join(GOLDEN, "netlists", "ro17.cir")).read()
    cards = src[src.index(".model"):src.rindex(".end")]
    tmp = tempfile.mkdtemp(prefix="ngb_cpu_arr_")
    cir = os.path.join(tmp, "arr.cir")
    open(cir, "w").write(synth.inverter_array_netlist(nx, ny, cards).replace(".option klu", ".option klu acct"))
    t0 = time.time()
    procs = [subprocess.Popen(["bash", "-c", f"{exe} -b -r {tmp}/raw{p} {cir} > {tmp}/log{p} 2>&1 ; rm -f {tmp}/raw{p}"]) for p in range(nproc)]
    for pr in procs:
        pr.wait()
    wall = time.time() - t0
    rate = 0.0
    for p in range(nproc):
        it, lt = 0, 0.0
        for ln in open(f"{tmp}/log{p}", errors="replace"):
            if ln.startswith("Total iterations"):
                it = int(ln.split("=")[-1].split()[0])
            if ln.startswith("Matrix load time"):
                lt = float(ln.split("=")[-1].split()[0])
        if it and lt:
            rate += 2 * nx * ny * it / lt
    return rate, wall, f"{nproc} processes x one {nx}x{ny} array transient each ({2 * nx * ny} BSIM4 instances); evals / STATloadTime, summed"


def bench_array(args):
    """BASELINE config 4: flat array of 2*cells BSIM4 transistors, one circuit (S = 1), instance-parallel
    CKTload (device loads + assembly).  LU of the ~10*cells unknowns is not part of this figure."""
    import torch
    pkg = importlib.import_module("ngspice-sf-mirror_b200")
    ngt = pkg.ngt
    synth = importlib.import_module("ngspice-sf-mirror_b200.synth")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = pkg.library()
    lib.check(lib.L.ngbInit(local), "ngbInit")
    stream = torch.cuda.Stream(device=local)
    lib.check(lib.L.ngbSetStream(ctypes.c_void_p(stream.cuda_stream)), "ngbSetStream")
    nx = int(round(args.cells ** 0.5)); ny = (args.cells + nx - 1) // nx
    base = ngt.read(f"{GOLDEN}/ro17k.flat.ngt")
    flat = synth.inverter_array(base, nx, ny)
    flat.pop("node/names", None)
    ninst = int(flat["b4/ninst"][0])
    circ = pkg.Circuit.from_flat(lib, flat)
    pat = circ.pattern()
    batch = pkg.Batch(circ, 1, device=local)
    neq1 = circ.neq + 1
    # a transient Newton iteration (MODETRAN | MODEINITFLOAT, order 1, h = 10 ps) on the bias pattern of a switching wave
    # crossing the array: cell (i, j) sits at phase (i + j) of it, its output is the inverse of its input, the internal
    # nodes of a transistor follow the terminal they hang on.  (Round 1 drew every node voltage uniformly at random: every
    # lane of a warp then sat in another operating region, and the divergence cost 2x -- NGB_ARRAY_RANDOM_BIAS=1 restores it.)
    rng = np.random.default_rng(rank)
    xhost = torch.empty((2, neq1, 1), dtype=torch.float64).pin_memory()
    xhost.numpy()[...] = 0.0
    if os.environ.get("NGB_ARRAY_RANDOM_BIAS"):
        xhost.numpy()[0, 1:, 0] = rng.uniform(0.0, 2.0, size=neq1 - 1)
    else:
        N = nx * ny
        cell = np.arange(N); ci, cj = cell // nx, cell % nx
        ph = 2.0 * np.pi * (ci + cj) / 64.0 + 0.01 * rng.normal(size=N)
        vin = 1.0 + np.tanh(4.0 * np.sin(ph)); vout = 2.0 - vin
        xv = xhost.numpy()[0, :, 0]
        xv[1] = 2.0; xv[2] = 0.0                                    # vdd, source
        xv[3 + 2 * cell] = vin; xv[4 + 2 * cell] = vout
        int0 = 3 + 2 * N
        for k, val in enumerate((vin, np.full(N, 2.0), np.full(N, 2.0), np.full(N, 2.0))):     # pmos: gate, dbody, body, sbody
            xv[int0 + 8 * cell + k] = val
        for k, val in enumerate((vin, np.zeros(N), np.zeros(N), np.zeros(N))):                # nmos
            xv[int0 + 8 * cell + 4 + k] = val
    out_A = torch.empty(pat["nnz"], dtype=torch.float64).pin_memory()
    out_x = torch.empty((2, neq1, 1), dtype=torch.float64).pin_memory()
    one = lambda v, dt: np.array([v], dtype=dt)
    batch.put("ctl.mode", one(0x1 | 0x100, np.int32)); batch.put("ctl.active", one(1, np.int32)); batch.put("ctl.order", one(1, np.int32))
    batch.put("ctl.ag0", one(1e11, np.float64)); batch.put("ctl.ag1", one(-1e11, np.float64)); batch.put("ctl.delta", one(1e-11, np.float64))
    batch.put("ctl.delta_old", np.full(7, 1e-11)); batch.put("ctl.time", one(1e-10, np.float64))
    batch.put("ctl.gmin", one(1e-12, np.float64)); batch.put("ctl.srcfact", one(1.0, np.float64))
    batch.put("x", xhost.numpy())

    def step(e2e):
        if e2e:
            batch.put("x", xhost.numpy())
        lib.check(lib.L.ngbLoad(batch.h), "ngbLoad")
        if e2e:
            lib.check(lib.L.ngbBatchDownload(batch.h, b"Ax", ctypes.c_void_p(out_A.data_ptr()), ctypes.c_long(out_A.numel() * 8), ctypes.c_long(0)), "download Ax")
            lib.check(lib.L.ngbBatchDownload(batch.h, b"x", ctypes.c_void_p(out_x.data_ptr()), ctypes.c_long(out_x.numel() * 8), ctypes.c_long(0)), "download rhs")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step(False)

    def timed(e2e, profile):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if profile:
            lib.L.ngbProfile(1, 1)
        barrier(); n0 = lib.launch_count()
        a.record(stream)
        for _ in range(args.steps):
            step(e2e)
        b.record(stream)
        barrier()
        prof = None
        if profile:
            msum, cnt = ctypes.c_double(), ctypes.c_long()
            lib.L.ngbProfileRead(ctypes.byref(msum), ctypes.byref(cnt)); lib.L.ngbProfile(0, 1)
            prof = (msum.value, cnt.value)
        return a.elapsed_time(b), lib.launch_count() - n0, prof

    with ClockSampler(local) as clk:
        ms, launches, prof = timed(False, True)
    clocks = clk.summary()
    ms_e2e, _, _ = timed(True, False)
    tt = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms, ms_e2e = tt.tolist()
    if rank == 0:
        hbm_peak, which = peaks()
        k_ms = prof[0] / max(prof[1], 1)
        achieved = ninst * B4_BYTES_PER_EVAL / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        cpu = array_cpu_baseline(os.cpu_count() or 1)
        line = {
            "metric": "BSIM4 instance-evals/s", "value": world * ninst * args.steps / (ms * 1e-3), "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"flat {nx}x{ny} BSIM4 inverter array with RC links (config 4), one circuit per GPU (replicas), "
                                   "one CKTload (device loads + assembly) per step, LU not included",
                       "bsim4_instances": ninst, "unknowns": pat["n"], "nnz": pat["nnz"],
                       "l2": f"inputs larger than L2: {ninst * 2000 / 1e6:.0f} MB touched per load"},
            "e2e": {"value": world * ninst * args.steps / (ms_e2e * 1e-3), "unit": "evals/s",
                    "h2d_bytes_per_step": int(xhost.numel() * 8), "d2h_bytes_per_step": int((out_A.numel() + out_x.numel()) * 8)},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak if hbm_peak else None, "traffic": traffic_from_profiles(ninst),
                         "kernel": "ngb_k_bsim4_load", "avg_launch_ms": k_ms, "timed_launches": prof[1],
                         "units_per_launch": ninst, "bytes_per_unit": B4_BYTES_PER_EVAL, "peak_source": which,
                         "kernel_share_of_step": k_ms / (ms / args.steps) if ms else None},
        }
        if cpu is not None:
            line["cpu_baseline"] = {"value": cpu[0], "unit": "evals/s", "cores": os.cpu_count() or 1, "kind": "reference", "sample": cpu[2]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def array_tran_reference(nx, ny, save):
    """the reference on the same netlist (one process: a single circuit has no second thread in the serial build): wall time of
    the whole run, its own counters, and the saved waveforms from the rawfile"""
    exe = os.path.join(ROOT, "oracle", "_ref", "ngspice")
    if not os.path.exists(exe):
        return None
    synth = importlib.import_module("ngspice-sf-mirror_b200.synth")
    src = open(os.path.join(GOLDEN, "netlists", "ro17.cir")).read()
    cards = src[src.index(".model"):src.rindex(".end")]
    tmp = tempfile.mkdtemp(prefix="ngb_cpu_arrtran_")
    cir = os.path.join(tmp, "arr.cir"); raw = os.path.join(tmp, "arr.raw")
    open(cir, "w").write(synth.inverter_array_netlist(nx, ny, cards).replace(".option klu", ".option klu acct"))
    t0 = time.time()
    out = subprocess.run([exe, "-b", "-r", raw, cir], capture_output=True, text=True).stdout
    wall = time.time() - t0
    st = {}
    for ln in out.splitlines():
        for key, tag in (("iters", "Total iterations"), ("load", "Matrix load time"), ("lu", "Matrix factor time"), ("reorder", "Matrix reorder time"),
                         ("solve", "Matrix solve time"), ("tran", "Transient analysis time"), ("analysis", "Total analysis time"), ("accepted", "Accepted timepoints"), ("rejected", "Rejected timepoints")):
            if ln.startswith(tag):
                try:
                    st[key] = float(ln.split("=")[-1].split()[0])
                except ValueError:
                    pass
    data = open(raw, "rb").read()
    i = data.index(b"Binary:\n"); head = data[:i].decode(errors="replace")
    nv = int([ln for ln in head.splitlines() if ln.startswith("No. Variables")][0].split(":")[1])
    npts = int([ln for ln in head.splitlines() if ln.startswith("No. Points")][0].split(":")[1])
    names = [ln.split()[1].lower() for ln in head.splitlines() if ln.startswith("\t") and len(ln.split()) >= 3 and ln.split()[0].isdigit()]
    arr = np.frombuffer(data, dtype=np.float64, count=nv * npts, offset=i + 8).reshape(npts, nv)
    cols = [names.index(f"v({n.lower()})") for n in save]
    return {"wall": wall, "stats": st, "time": arr[:, 0].copy(), "values": arr[:, cols].copy()}


def bench_array_tran(args):
    """BASELINE config 4 with the LU: DC operating point + `.tran 10p 1n` of an nx-by-ny array, one circuit (S = 1).  The matrix does
    not fit a CTA's shared memory, so refactor + solve run in the grid-wide LU kernel; ordering and pivoting are the library's own
    (ngbCircuitAnalyze / ngbCircuitFactor: minimum degree, not KLU's AMD), so the result agrees with the reference to rounding along
    the same time grid, not bit for bit (with KLU's orders it is bit-identical: tests/test_synth_array.py)."""
    import torch
    pkg = importlib.import_module("ngspice-sf-mirror_b200")
    ngt = pkg.ngt
    synth = importlib.import_module("ngspice-sf-mirror_b200.synth")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    local = int(os.environ.get("LOCAL_RANK", "0")); rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = pkg.library()
    lib.check(lib.L.ngbInit(local), "ngbInit")
    stream = torch.cuda.Stream(device=local)
    lib.check(lib.L.ngbSetStream(ctypes.c_void_p(stream.cuda_stream)), "ngbSetStream")
    nx = int(round(args.cells ** 0.5)); ny = (args.cells + nx - 1) // nx
    base = ngt.read(f"{GOLDEN}/ro17k.flat.ngt")
    flat = synth.inverter_array(base, nx, ny)
    flat.pop("node/names", None)
    ninst = int(flat["b4/ninst"][0])
    t_setup = time.time()
    circ = pkg.Circuit.from_flat(lib, flat)
    pat = circ.pattern()
    one = lambda v, dt: np.array([v], dtype=dt)
    b0 = pkg.Batch(circ, 1, device=local)            # the matrix of the first iteration (MODETRANOP | MODEINITJCT from zero): what the first pivoting factor sees
    b0.put("ctl.mode", one(0x20 | 0x200, np.int32)); b0.put("ctl.active", one(1, np.int32)); b0.put("ctl.order", one(1, np.int32))
    b0.put("ctl.gmin", one(1e-12, np.float64)); b0.put("ctl.srcfact", one(1.0, np.float64))
    b0.load()
    Ax0 = np.ascontiguousarray(b0.get("Ax", (1, -1))[0])
    del b0
    lib.check(lib.L.ngbCircuitAnalyze(circ.h), "ngbCircuitAnalyze")
    lib.L.ngbCircuitFactor.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.c_double]
    lib.check(lib.L.ngbCircuitFactor(circ.h, Ax0.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), ctypes.c_double(1e-3)), "ngbCircuitFactor")
    info = circ.lu_info()
    batch = pkg.Batch(circ, 1, device=local)
    t_setup = time.time() - t_setup
    save_names = ["out_0_0", f"out_{ny - 1}_{nx - 1}", "in_2_1"]
    cell_of = lambda i, j: i * nx + j                                   # node numbering of synth.inverter_array: in = 3 + 2 cell, out = 4 + 2 cell
    save = np.array([4 + 2 * cell_of(0, 0), 4 + 2 * cell_of(ny - 1, nx - 1), 3 + 2 * cell_of(2, 1)], np.int32)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    last = {}

    def step(e2e):
        res = batch.tran(1024, save)
        if e2e:
            last["t"], last["v"] = res.waves()
        last["res"] = res

    for _ in range(max(args.warmup, 3)):
        step(False)

    def timed(e2e, profile):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if profile:
            lib.L.ngbProfile(1, 16)
        barrier(); n0 = lib.launch_count()
        a.record(stream)
        for _ in range(args.steps):
            step(e2e)
        b.record(stream)
        barrier()
        prof = None
        if profile:
            msum, cnt = ctypes.c_double(), ctypes.c_long()
            lib.L.ngbProfileRead(ctypes.byref(msum), ctypes.byref(cnt)); lib.L.ngbProfile(0, 1)
            prof = (msum.value, cnt.value)
        return a.elapsed_time(b), lib.launch_count() - n0, prof

    with ClockSampler(local) as clk:
        ms, launches, prof = timed(False, True)
    clocks = clk.summary()
    stage_ms = (ctypes.c_double * 8)()
    nstage = lib.L.ngbProfileStages(stage_ms)
    ms_e2e, _, _ = timed(True, False)
    res = last["res"]
    numiter = int(res.numiter[0]); npts = int(res.npoints[0])
    tt = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms, ms_e2e = tt.tolist()
    if rank == 0:
        hbm_peak, which = peaks()
        k_ms = prof[0] / max(prof[1], 1)
        achieved = ninst * B4_BYTES_PER_EVAL / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
        evals = ninst * numiter
        ref = array_tran_reference(nx, ny, save_names)
        line = {
            "metric": "BSIM4 instance-evals/s", "value": world * evals * args.steps / (ms * 1e-3), "unit": "evals/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"flat {nx}x{ny} BSIM4 inverter array with RC links (config 4 at reduced size), one circuit per GPU (replicas): "
                                   "DC operating point + .tran 10p 1n per step, device loads + assembly + grid-wide LU + NIiter / DCtran on the device; "
                                   "ordering and pivoting by the library (minimum degree, own pivoting factor)",
                       "bsim4_instances": ninst, "unknowns": pat["n"], "nnz": pat["nnz"], "lu_values": info["nV"], "lu_levels": info["nlev"] + info["nslev"],
                       "lu_products": info["npairs"], "setup_s_not_timed": t_setup,
                       "l2": "working set (states, stamps, matrix, factors) smaller than L2 below ~64 x 64; no flush"},
            "newton_iterations": numiter, "accepted_points": int(res.accepted[0]), "rejected_points": int(res.rejected[0]),
            "us_per_newton_iteration": ms / args.steps / max(numiter, 1) * 1e3,
            "stage_us_per_sampled_step": {n: stage_ms[k] / max(nstage, 1) * 1e3 for k, n in
                                          ((1, "small loads"), (2, "bsim4_load"), (3, "assemble"), (4, "lu"), (5, "bsim4_lte"), (6, "control"))},
            "e2e": {"value": world * evals * args.steps / (ms_e2e * 1e-3), "unit": "evals/s",
                    "h2d_bytes_per_step": int(save.nbytes), "d2h_bytes_per_step": int(npts * (len(save) + 1) * 8)},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak if hbm_peak else None, "traffic": None,
                         "kernel": "ngb_k_bsim4_load", "avg_launch_ms": k_ms, "timed_launches": prof[1],
                         "units_per_launch": ninst, "bytes_per_unit": B4_BYTES_PER_EVAL, "peak_source": which,
                         "kernel_share_of_step": k_ms * numiter / (ms / args.steps) if ms else None},
        }
        if ref is not None:
            st = ref["stats"]
            it_ref = int(st.get("iters", 0))
            line["cpu_baseline"] = {"value": ninst * it_ref / (st.get("analysis") or ref["wall"]), "unit": "evals/s", "cores": 1, "kind": "reference",
                                    "sample": f"the same {nx}x{ny} netlist, one reference process, its own 'Total analysis time' (DC op + transient, no parse); its counters: "
                                              f"load {st.get('load')} s, LU {st.get('lu')} + {st.get('reorder')} s, solve {st.get('solve')} s, {it_ref} iterations",
                                    "us_per_newton_iteration": (st.get("load", 0) + st.get("lu", 0) + st.get("reorder", 0) + st.get("solve", 0)) / max(it_ref, 1) * 1e6,
                                    "analysis_s": st.get("analysis")}
            t, v = last["t"][0, :npts], last["v"][0, :npts]
            same_grid = npts == len(ref["time"])
            pc = {"accepted_identical": int(res.accepted[0]) == int(st.get("accepted", -1)) and int(res.rejected[0]) == int(st.get("rejected", -1)),
                  "points_identical": bool(same_grid), "iterations": [numiter, it_ref], "tolerance": 1e-9,
                  "note": "own pivot order (minimum degree, not KLU's AMD): same accepted points, values to rounding; the Newton path of the "
                          "operating point, hence the iteration count, may differ"}
            err = None
            if same_grid:                # per point: |v - v_ref| <= tol * max(|v_ref|, vntol)
                pc["max_rel_err_time"] = float(np.max(np.abs(t - ref["time"]) / np.maximum(ref["time"], 1e-300)))
                err = float(np.max(np.abs(v - ref["values"]) / np.maximum(np.abs(ref["values"]), 1e-6)))
            pc["max_rel_err"] = err
            pc["ok"] = bool(pc["accepted_identical"] and same_grid and err <= 1e-9 and pc["max_rel_err_time"] <= 1e-9)
            line["parity_check"] = pc
        print(json.dumps(line))
        if ref is not None and not line["parity_check"]["ok"]:
            raise SystemExit("bench.py: array transient differs from the reference beyond 1e-9")
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ config 5: mixed-model parameter sweep
SWEEP_GRID = 256                      # 256 x 256 = 65 536 points of (vdd, r1); 8 192 points per GPU on 8 GPUs
SWEEP_COUNTS = {"bsim4": 16, "bsim3": 8, "vbic": 2, "diode": 3}


def sweep_points(rank, world, n):
    """the (vdd, r1) netlist tokens of this rank's share of the grid: consecutive rows of the 256 x 256 grid"""
    pkg = importlib.import_module("ngspice-sf-mirror_b200")
    vt = pkg.sweep.grid_tokens(1.6, 2.4, SWEEP_GRID)
    rt = pkg.sweep.grid_tokens(200.0, 5000.0, SWEEP_GRID, fmt="{:.5g}")
    first = (rank * n) % (SWEEP_GRID * SWEEP_GRID)
    idx = (first + np.arange(n)) % (SWEEP_GRID * SWEEP_GRID)
    return [vt[i // SWEEP_GRID] for i in idx], [rt[i % SWEEP_GRID] for i in idx]


def sweep_cpu_baseline(nproc, per_proc=128, own=None):
    """the reference on the same sweep: `nproc` processes, each running `per_proc` grid points one after the other.
    own = (vdd tokens, r tokens) of the calling rank's points: the reference then simulates every (len / n)-th of THOSE points
    and keeps its rawfiles -- the parity check of the benchmarked run; the return value gets a fifth element [(point, rawfile)]"""
    exe = os.path.join(ROOT, "oracle", "_ref", "ngspice")
    if not os.path.exists(exe):
        return None
    base = open(os.path.join(GOLDEN, "netlists", "mix.cir")).read()
    n = nproc * per_proc
    stride = (SWEEP_GRID * SWEEP_GRID) // n
    pkg = importlib.import_module("ngspice-sf-mirror_b200")
    vt = pkg.sweep.grid_tokens(1.6, 2.4, SWEEP_GRID)
    rt = pkg.sweep.grid_tokens(200.0, 5000.0, SWEEP_GRID, fmt="{:.5g}")
    tmp = tempfile.mkdtemp(prefix="ngb_sweep_")
    procs = []
    kept = []
    t0 = time.time()
    for p in range(nproc):
        files = []
        for k in range(per_proc):
            if own is not None:
                q = ((p * per_proc + k) * len(own[0])) // n                      # evenly through this rank's share
                vtok, rtok = own[0][q], own[1][q]
            else:
                i = (p * per_proc + k) * stride + (p * 37 + k * 11) % SWEEP_GRID      # spread over the grid
                q, vtok, rtok = -1, vt[(i // SWEEP_GRID) % SWEEP_GRID], rt[i % SWEEP_GRID]
            text = base.replace("vdd dd 0 dc 2.0", f"vdd dd 0 dc {vtok}").replace("r1 a8 x 1k", f"r1 a8 x {rtok}")
            text = text.replace(".option klu", ".option klu acct")
            f = os.path.join(tmp, f"p{p}_{k}.cir")
            open(f, "w").write(text)
            files.append(f)
            kept.append((q, f + ".raw"))
        cmd = " ; ".join(f"{exe} -b -r {f}.raw {f} > {f}.log 2>&1" + ("" if own is not None else f" ; rm -f {f}.raw") for f in files)
        procs.append((subprocess.Popen(["bash", "-c", cmd]), files))
    iters = 0
    for pr, files in procs:
        pr.wait()
    wall = time.time() - t0
    for _, files in procs:
        for f in files:
            for ln in open(f + ".log", errors="replace"):
                if ln.startswith("Total iterations"):
                    iters += int(ln.split("=")[-1].strip().split()[0])
    return n / wall, iters, wall, f"{n} grid points ({per_proc} per process x {nproc} processes), DC op + 200-step transient each", kept


def sweep_parity(kept, save_names, npts, t_gpu, v_gpu, tol=1e-7):
    """the reference's rawfiles of the rank's own grid points against the waveforms of the end-to-end pass: per accepted point
    |v - v_ref| <= tol * max(|v_ref|, 1e-6 V).  Not bit-identical by construction: the VBIC Jacobian comes
    from dual numbers, and the cell's ring oscillator carries that rounding through zero crossings, where the rule's floor of
    1e-6 V turns 5e-14 V into 5e-8 (worst of 2 048 points; 2 047 are within 1e-8, tests/test_tran_parity.py uses 1e-8 on its
    eight points).  Identical point counts and time points within 1e-9 are required of every point"""
    same, within, within8, worst, worst_t, n = 0, 0, 0, 0.0, 0.0, 0
    for q, raw in kept:
        try:
            tt, vv = read_rawfile(raw, [f"v({s})" for s in save_names])
        except Exception:
            continue
        finally:
            try:
                os.remove(raw)
            except OSError:
                pass
        n += 1
        k = len(tt)
        if int(npts[q]) != k:
            continue
        same += 1
        ref = np.stack([vv[f"v({s})"] for s in save_names], axis=1)
        e = float(np.max(np.abs(v_gpu[q, :k, :] - ref) / np.maximum(np.abs(ref), 1e-6)))
        et = float(np.max(np.abs(t_gpu[q, :k] - tt) / np.maximum(tt, 1e-300)))
        worst = max(worst, e); worst_t = max(worst_t, et)
        within += (e <= tol and et <= 1e-9); within8 += (e <= 1e-8 and et <= 1e-9)
    return {"points": n, "same_point_count": same, "within_tolerance": int(within), "tolerance": tol, "within_1e-8": int(within8), "max_rel_err": worst,
            "max_rel_err_time": worst_t, "ok": bool(n > 0 and same == n and within == n)}


def bench_sweep(args):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("ngspice-sf-mirror_b200"); ngt = pkg.ngt
    run_patterns = ngt.run_patterns

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            cores = os.cpu_count() or 1
            r = sweep_cpu_baseline(cores)
            if r is None:
                print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ngspice not built"}))
            else:
                print(json.dumps({"impl": "reference", "metric": "sweep points/s", "value": r[0], "unit": "points/s", "n_gpus": args.gpus,
                                  "steps": 1, "warmup": 0, "ms_per_step": r[2] * 1e3, "higher_is_better": True, "scaling": "weak",
                                  "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": sweep_config(args),
                                  "cpu_baseline": {"value": r[0], "unit": "points/s", "cores": cores, "kind": "reference", "sample": r[3]},
                                  "e2e": {"value": r[0], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = pkg.library()
    assert lib.backend == "cuda-sm_100a"
    lib.check(lib.L.ngbInit(local), "ngbInit")
    stream = torch.cuda.Stream(device=local)
    lib.check(lib.L.ngbSetStream(ctypes.c_void_p(stream.cuda_stream)), "ngbSetStream")

    flat = ngt.read(f"{GOLDEN}/mix.flat.ngt")
    trace = ngt.read(f"{GOLDEN}/mix.trace.ngt.gz")
    wave = ngt.read(f"{GOLDEN}/mix.wave.ngt")
    S = args.samples
    circ = pkg.Circuit.from_flat(lib, flat, lu_pattern=run_patterns(trace))
    batch = pkg.Batch(circ, S, device=local)
    save_eq = wave["save_eq"][:5]                     # a8, x, y4, cq, e2
    max_points = 320
    vdd_tok, r_tok = sweep_points(rank, world, S)
    vpar = pkg.sweep.vsource_table(flat, S, {"vdd": vdd_tok})
    gtab = pkg.sweep.resistor_table(flat, S, {"r1": r_tok})
    out_t = torch.empty((S, max_points), dtype=torch.float64).pin_memory()
    out_v = torch.empty((S, max_points, len(save_eq)), dtype=torch.float64).pin_memory()
    h2d_bytes = vpar.nbytes + gtab.nbytes
    d2h_bytes = (out_t.numel() + out_v.numel()) * 8

    last_res = [None]

    def step(e2e):
        if e2e:
            batch.put("vsrc.par", vpar)
            batch.set_resistors(gtab)
        res = batch.tran(max_points, save_eq)
        last_res[0] = res
        if e2e:
            lib.check(lib.L.ngbTranWaves(batch.h, ctypes.cast(out_t.data_ptr(), ctypes.POINTER(ctypes.c_double)),
                                         ctypes.cast(out_v.data_ptr(), ctypes.POINTER(ctypes.c_double))), "ngbTranWaves")
        return res

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    batch.put("vsrc.par", vpar)
    batch.set_resistors(gtab)
    res = None
    for _ in range(max(args.warmup, 1)):                         # the completion check below needs one run at least
        res = step(False)
    bad = int((res.accepted.astype(np.int64) < 100).sum())       # points whose transient did not run through

    def timed(e2e):
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        iters = 0; ticks = 0
        barrier()
        n0 = lib.launch_count()
        for k in range(args.steps):
            evs[k][0].record(stream)
            r = step(e2e)
            evs[k][1].record(stream)
            iters += int(r.numiter.astype(np.int64).sum()); ticks += r.ticks
        barrier()
        return sum(a.elapsed_time(b) for a, b in evs), iters, ticks, lib.launch_count() - n0

    with ClockSampler(local) as clk:
        ms, iters, ticks, launches = timed(False)
    clocks = clk.summary()
    ms_e2e, iters_e2e, _, _ = timed(True)
    tt = torch.tensor([ms, ms_e2e], dtype=torch.float64, device="cuda")
    ww = torch.tensor([float(iters), float(S * args.steps), float(bad)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(ww, op=dist.ReduceOp.SUM)
    ms_max, ms_e2e_max = tt.tolist()
    iters_tot, points_tot, bad_tot = ww.tolist()
    if rank == 0:
        hbm_peak, which = peaks()
        ndev = sum(SWEEP_COUNTS.values())
        # algorithmic bytes of one Newton step of one point: the four load kernels' tables, states and stamps
        # (BSIM4 2 000 B, BSIM3 1 400 B, VBIC 3 400 B, diode 700 B per evaluation)
        bytes_step = SWEEP_COUNTS["bsim4"] * 2000 + SWEEP_COUNTS["bsim3"] * 1400 + SWEEP_COUNTS["vbic"] * 3400 + SWEEP_COUNTS["diode"] * 700
        achieved = iters_tot * bytes_step / (ms_max * 1e-3) / 1e9 / world          # per GPU
        cpu = sweep_cpu_baseline(os.cpu_count() or 1, own=(vdd_tok, r_tok))
        parity = None
        if cpu is not None:
            names = bytes(np.asarray(flat["node/names_bytes"]).astype(np.uint8)).decode().split("\n")
            eq_name = {int(ln.split()[0]): ln.split()[1].lower() for ln in names if ln.strip()}
            parity = sweep_parity(cpu[4], [eq_name[int(e)] for e in save_eq], last_res[0].npoints, out_t.numpy(), out_v.numpy())
        line = {
            "metric": "sweep points/s", "value": points_tot / (ms_max * 1e-3), "unit": "points/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": sweep_config(args),
            "device_evals_per_s": ndev * iters_tot / (ms_max * 1e-3), "newton_iterations_per_point": iters_tot / points_tot,
            "newton_steps_per_sweep": ticks / args.steps, "points_not_completed": int(bad_tot),
            "e2e": {"value": points_tot / (ms_e2e_max * 1e-3), "unit": "points/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak if hbm_peak else None,
                         "traffic": None, "kernel": "whole Newton step (all load kernels + assembly + LU), algorithmic device-load bytes only",
                         "bytes_per_unit": bytes_step, "peak_source": which},
        }
        if cpu is not None:
            line["cpu_baseline"] = {"value": cpu[0], "unit": "points/s", "cores": os.cpu_count() or 1, "kind": "reference", "sample": cpu[3]}
            line["parity_check"] = parity
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def sweep_config(args):
    return {"workload": "mixed-model sweep cell (16 BSIM4 + 8 BSIM3 + 2 VBIC + 3 diodes + R/C, 104 unknowns), 256 x 256 grid of "
                        "(vdd 1.6-2.4 V, r1 200-5000 ohm), DC operating point + .tran 25p 5n per point (config 5)",
            "points_per_gpu": args.samples, "l2": "per-step working set larger than L2 at 8 192 points"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="mc_ro17", choices=["mc_ro17", "ro101", "array", "array_tran", "sweep"])
    ap.add_argument("--cells", type=int, default=None, help="inverters of the array workloads (2 transistors each): default 500 000 (array), 4 096 (array_tran)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --samples per GPU (default, the driver's scaling run); strong: --samples in total, samples / N per GPU "
                         "(BASELINE config 3 as written: 4096 samples, batch = 4096 / N)")
    ap.add_argument("--tox", default="continuous", choices=["continuous", "levels"],
                    help="per-sample oxide thickness: continuous N(1.4 nm, 3 %%) through the library's BSIM4temp (default), or round 1's 8 recorded levels")
    ap.add_argument("--samples", type=int, default=None, help="Monte-Carlo samples (default 4096) or sweep points (default 8192) per GPU")
    args = ap.parse_args()
    if args.samples is None:
        args.samples = 8192 if args.workload == "sweep" else 4096
    if args.cells is None:
        args.cells = 4096 if args.workload == "array_tran" else 500000
    if args.workload == "sweep":
        bench_sweep(args)
    elif args.workload == "array_tran":
        bench_array_tran(args)
    elif args.impl == "reference":
        bench_reference(args)
    elif args.workload == "array":
        bench_array(args)
    else:
        bench_ours(args)


if __name__ == "__main__":
    main()
