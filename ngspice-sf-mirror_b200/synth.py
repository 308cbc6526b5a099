"""Synthetic flat circuits for the large-array workload (BASELINE config 4).

`inverter_array(base, nx, ny)` tiles one PMOS/NMOS pair of a flattened reference circuit (its
BSIM4temp results: instance rows, model and bin tables) into an nx-by-ny grid of CMOS inverters.
Cell (i, j) drives the inputs of its right and bottom neighbours through 100-ohm resistors, every
input carries 1 fF to ground, cell (0, 0) is driven by a PULSE source, all cells share one supply:
10 unknowns per cell (input, output, 2 x 4 internal BSIM4 nodes) plus two source branches.  The
node numbering is our own (any consistent numbering is a valid CKTsetup result); the same
generator also writes the netlist, so small grids can be run through the reference for parity
(tests/golden/make_golden.py `arr`)."""
import numpy as np
from . import mc


def inverter_array_netlist(nx, ny, cards, vdd=2.0, tran=".tran 10p 1n"):
    lines = [f"* {nx}x{ny} BSIM4 inverter array", f"vdd vdd 0 {vdd}", f"vin src 0 pulse(0 {vdd} 0 50p 50p 0.4n 1n)",
             "rsrc src in_0_0 100"]
    for i in range(ny):
        for j in range(nx):
            lines.append(f"mp_{i}_{j} out_{i}_{j} in_{i}_{j} vdd vdd p1 l=0.1u w=10u ad=5p pd=6u as=5p ps=6u")
            lines.append(f"mn_{i}_{j} out_{i}_{j} in_{i}_{j} 0 0 n1 l=0.1u w=5u ad=5p pd=6u as=5p ps=6u")
            lines.append(f"c_{i}_{j} in_{i}_{j} 0 1f")
            if j + 1 < nx:
                lines.append(f"rr_{i}_{j} out_{i}_{j} in_{i}_{j + 1} 100")
            if i + 1 < ny:
                lines.append(f"rb_{i}_{j} out_{i}_{j} in_{i + 1}_{j} 100")
    lines += [".option klu", tran]
    return "\n".join(lines) + "\n" + cards + "\n.end\n"


def inverter_array(base, nx, ny, vdd=2.0, tstep=1e-11, tstop=1e-9):
    """base: flat dict of a circuit that holds at least one `mp*` and one `mn*` BSIM4 instance with
    rgateMod = 1 and rbodyMod = 1 (the ring-oscillator fixtures).  Returns a flat dict."""
    names = mc.instance_names(base)
    ip = next(k for k, n in enumerate(names) if n.lower().startswith("mp"))
    inn = next(k for k, n in enumerate(names) if n.lower().startswith("mn"))
    N = nx * ny
    cell = np.arange(N, dtype=np.int64)
    ci, cj = cell // nx, cell % nx
    VDD, SRC = 1, 2
    n_in = 3 + 2 * cell
    n_out = 4 + 2 * cell
    int0 = 3 + 2 * N
    pint = int0 + 8 * cell                     # pmos: gate, dbody, body, sbody
    nint = pint + 4
    br_vdd = int0 + 8 * N
    br_vin = br_vdd + 1
    neq = int(br_vin)

    def b4_nodes(d, g, s, b, internal):
        z = np.zeros(N, np.int64)
        # dNode gNodeExt sNode bNode dNodePrime gNodePrime gNodeMid sNodePrime bNodePrime dbNode sbNode qNode
        return np.stack([d, g, s, b, d, internal + 0, g, s, internal + 2, internal + 1, internal + 3, z])

    full = np.full(N, VDD, np.int64)
    zero = np.zeros(N, np.int64)
    nodes = np.concatenate([b4_nodes(n_out, n_in, full, full, pint), b4_nodes(n_out, n_in, zero, zero, nint)], axis=1)
    inst = np.asarray(base["b4/inst"], np.float64)
    flat = {k: v for k, v in base.items() if k.startswith("opt/")}
    flat["meta/neq"] = np.array([neq], np.int32)
    nt = np.full(neq + 1, 3, np.int32); nt[br_vdd] = 4; nt[br_vin] = 4
    flat["node/type"] = nt
    flat["tran/tstep"] = np.array([tstep]); flat["tran/tstop"] = np.array([tstop]); flat["tran/tmax"] = np.array([tstep])
    flat["tran/tstart"] = np.array([0.0]); flat["tran/uic"] = np.array([0], np.int32)
    # the options the base fixture's own netlist set or DCtran derived from ITS .tran line (the oscillator example runs with
    # xmu = 0.49): the array netlist has `.option klu` only
    flat["opt/xmu"] = np.array([0.5])
    flat["opt/delmin"] = np.array([1e-11 * tstep])           # CKTdelmin = 1e-11 * CKTmaxStep (dctran.c:137)
    flat["opt/minbreak"] = np.array([tstep * 5e-5])          # CKTminBreak = CKTmaxStep * 5e-5 (dctran.c:186)
    flat["b4/ninst"] = np.array([2 * N], np.int32)
    flat["b4/nodes"] = nodes.astype(np.int32)
    flat["b4/flags"] = np.concatenate([np.full(N, base["b4/flags"][ip]), np.full(N, base["b4/flags"][inn])]).astype(np.int32)
    flat["b4/prow"] = np.concatenate([np.full(N, base["b4/prow"][ip]), np.full(N, base["b4/prow"][inn])]).astype(np.int32)
    flat["b4/inst"] = np.concatenate([np.repeat(inst[:, ip:ip + 1], N, axis=1), np.repeat(inst[:, inn:inn + 1], N, axis=1)], axis=1)
    flat["b4/mtab"] = base["b4/mtab"]; flat["b4/ptab"] = base["b4/ptab"]
    # resistors: source feed, right and bottom links
    right = cell[cj + 1 < nx]; bottom = cell[ci + 1 < ny]
    rp = np.concatenate([[SRC], n_out[right], n_out[bottom]])
    rq = np.concatenate([[n_in[0]], n_in[right + 1], n_in[bottom + nx]])
    flat["res/n"] = np.array([len(rp)], np.int32)
    flat["res/nodes"] = np.stack([rp, rq]).astype(np.int32)
    flat["res/g"] = np.full(len(rp), 1.0 / 100.0)
    flat["cap/n"] = np.array([N], np.int32)
    flat["cap/nodes"] = np.stack([n_in, zero]).astype(np.int32)
    flat["cap/par"] = np.stack([np.full(N, 1e-15), np.ones(N), np.zeros(N)])
    flat["vsrc/n"] = np.array([2], np.int32)
    flat["vsrc/nodes"] = np.array([[VDD, SRC], [0, 0], [br_vdd, br_vin]], np.int32)
    flat["vsrc/fn"] = np.array([[0, 1], [0, 7], [1, 0]], np.int32)           # DC ; PULSE with 7 coefficients
    par = np.zeros((9, 2)); par[0, 0] = vdd
    par[1:8, 1] = [0.0, vdd, 0.0, 50e-12, 50e-12, 0.4e-9, 1e-9]
    flat["vsrc/par"] = par
    flat["isrc/n"] = np.array([0], np.int32)
    if N <= 4096:                                           # node names (reference spelling) for small grids
        nm = [""] * (neq + 1)
        nm[0], nm[VDD], nm[SRC] = "0", "vdd", "src"
        for c in range(N):
            i, j = int(ci[c]), int(cj[c])
            nm[n_in[c]] = f"in_{i}_{j}"; nm[n_out[c]] = f"out_{i}_{j}"
            for base_eq, dev in ((pint[c], "mp"), (nint[c], "mn")):
                for k, suf in enumerate(("gate", "dbody", "body", "sbody")):
                    nm[base_eq + k] = f"{dev}_{i}_{j}#{suf}"
        nm[br_vdd], nm[br_vin] = "vdd#branch", "vin#branch"
        flat["node/names"] = nm
    return flat
