"""Parameter sweeps (BASELINE config 5): every sample of a batch is the same topology with its own
source values and resistor values -- what a `.control` loop of `alter vdd = ...` / `alter r1 = ...`
followed by `tran` does in the reference, one run after the other.

Values are given the way the reference gets them, as netlist text: `spice_number` (mc.py) reproduces
INPevaluate's result for the token, and the conductance follows REStemp
(src/spicelib/devices/res/restemp.c: RESconduct = m / (R * (1 + tc1*dT + tc2*dT^2) * scale), which is
m / R at the nominal temperature), so the parameters are the reference's bit for bit."""
import numpy as np
from . import mc


def _names(flat, key):
    return [n.lower() for n in bytes(np.asarray(flat[key]).astype(np.uint8)).decode().split("\n")[:-1]]


def resistor_table(flat, S, values=None):
    """g [nres][S]: the circuit's conductances, with `values` {name: [S] resistance tokens or floats} replaced"""
    g = np.repeat(np.asarray(flat["res/g"], np.float64)[:, None], S, axis=1)
    names = _names(flat, "res/names_bytes")
    for name, vals in (values or {}).items():
        r = np.array([mc.spice_number(v) if isinstance(v, str) else float(v) for v in vals], np.float64)
        g[names.index(name.lower())] = 1.0 / r
    return g


def vsource_table(flat, S, dc=None):
    """par [9][nv][S]: the circuit's source parameters, with the DC value of `dc` {name: [S] tokens or floats} replaced"""
    par = np.repeat(np.asarray(flat["vsrc/par"], np.float64)[:, :, None], S, axis=2)
    names = _names(flat, "vsrc/names_bytes")
    for name, vals in (dc or {}).items():
        par[0, names.index(name.lower())] = [mc.spice_number(v) if isinstance(v, str) else float(v) for v in vals]
    return par


def apply(batch, flat, dc=None, res=None):
    """write the per-sample tables of a sweep into a batch"""
    if dc:
        batch.put("vsrc.par", vsource_table(flat, batch.S, dc))
    if res:
        batch.set_resistors(resistor_table(flat, batch.S, res))


def grid_tokens(lo, hi, n, fmt="{:.6g}"):
    """n evenly spaced values as netlist tokens (the text is the value: both sides parse the same string)"""
    return [fmt.format(v) for v in np.linspace(lo, hi, n)]
