#!/usr/bin/env python3
"""the handful of numbers of a bench.py JSON line that the experiment logs quote: python tools/bench_brief.py line.json"""
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline", {}); p = d.get("parity_check", {})
    print(f"ms_per_step {d['ms_per_step']:.1f} value {d['value']:.4g} e2e {d['e2e']['value']:.4g} load_us {r.get('avg_launch_ms', 0) * 1e3:.1f} "
          f"frac {r.get('frac', 0):.4f} share {r.get('kernel_share_of_step', 0):.3f} parity_ok {p.get('ok')} bit_identical {p.get('bit_identical')} "
          f"failed {d.get('samples_failed')} clocks {d.get('clocks')}")
except Exception as e:
    print("unreadable bench line:", e, open(sys.argv[1]).read()[-600:])
