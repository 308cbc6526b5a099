#!/usr/bin/env python
"""Generates the committed golden fixtures by running the REFERENCE ngspice (compiled by
oracle/build_ref.sh into oracle/_ref/) on netlists assembled from the reference's own example
and test files.  Needs /root/reference; the fixtures it writes do not.

  ro17   examples/mos/ro_17_4.cir (17-stage BSIM4 ring oscillator, `.tran .1ns 150ns uic`),
         model cards switched to `version = 4.8.3` so the BSIM4 4.8.3 code of b4ld.c is the one
         that runs (4.5.0 would select the separate bsim4v5 device), `.option xmu=0.49 klu`.
  ro101  same cards, 101 stages (BASELINE config 2).
  inv    tests/bsim4/{nmos,pmos}/parameters cards, CMOS inverter with PULSE input (config 1).
  b3ring BSIM3v3.3.0 ring of five inverters + buffer on the MC_ring.sp level-8 cards.
  arr    4x4 BSIM4 inverter array with RC links (small instance of config 4's generator).
  ro17tox BSIM4temp tables for 8 oxide-thickness levels + two reference transients with toxe and delvto mismatch.
  b3cap  b3c<capMod>x<xpart> : the 12 capMod / xpart combinations of BSIM3 on a 3-stage ring, 3 ns.
  dio    junction diodes (rectifier, zener clamp, sidewall/tunnel/knee parameters) with R, C, SIN source.

Outputs (tests/golden/): <name>.flat.ngt  flattened circuit after CKTsetup/CKTtemp
                         <name>.trace.ngt.gz recorded CKTload / KLU calls (subset)
                         <name>.wave.ngt  accepted time points, saved waveforms, run statistics
"""
import json
import os
import re
import subprocess
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from parity_util import ngt  # noqa: E402

REF = os.environ.get("NGB_REFERENCE", "/root/reference")
DUMP = os.path.join(ROOT, "oracle", "_ref", "ngspice_dump")
TMP = os.environ.get("NGB_TMP", "/tmp/ngb_golden")


def ro_cards():
    src = open(os.path.join(REF, "examples/mos/ro_17_4.cir")).read()
    i = src.index(".model  N1")
    cards = src[i:]
    cards = cards.replace("version = 4.5.0", "version = 4.8.3")
    return cards


def ro_netlist(stages, tran=".tran .1ns 150ns uic", extra_opts="", kick=False, delvto=None):
    lines = [f"* {stages}-stage BSIM4 ring oscillator (topology of examples/mos/ro_17_4.cir)", "vdd 1 0 2.0"]
    for k in range(1, stages + 1):
        a = k + 1
        out = k + 2 if k < stages else 2
        dp = f" delvto={delvto[2 * (k - 1)]:.17g}" if delvto is not None else ""
        dn = f" delvto={delvto[2 * (k - 1) + 1]:.17g}" if delvto is not None else ""
        lines.append(f"mp{k} {out} {a} 1 1 p1 l=0.1u w=10u ad=5p pd=6u as=5p ps=6u{dp}")
        lines.append(f"mn{k} {out} {a} 0 0 n1 l=0.1u w=5u ad=5p pd=6u as=5p ps=6u{dn}")
    lines.append(f"c1 {stages + 1} 0 .1p")
    if kick:
        # deterministic start: alternate the stage outputs between the rails so the oscillation
        # does not have to grow out of rounding noise (the stock example starts from all-zero)
        ics = " ".join(f"v({k + 1})={2.0 if k % 2 else 0.0}" for k in range(1, stages + 1))
        lines.append(".ic " + ics)
    lines.append(f".option xmu=0.49 klu {extra_opts}")
    lines.append(tran)
    return "\n".join(lines) + "\n" + ro_cards() + "\n.end\n"


def qa_card(kind):
    """model card of the reference's BSIM4 QA suite (tests/bsim4/<kind>/parameters), selected the way
    its qaSpec does for ngspice: level=14 version=4.8.1"""
    body = open(os.path.join(REF, f"tests/bsim4/{kind}/parameters/{kind}Parameters")).read()
    return f".model {kind[0]}inv {kind} level=14 version=4.8.1\n" + body


def inv_netlist():
    """BASELINE config 1: CMOS inverter, QA-suite cards, PULSE input, 10 fF load, DC operating
    point followed by `.tran 10p 10n`"""
    return "\n".join([
        "* CMOS inverter on the tests/bsim4 QA model cards",
        "vdd vdd 0 1.2",
        "vin in 0 pulse(0 1.2 0 50p 50p 1n 2n)",
        "mp out in vdd vdd pinv w=10e-6 l=0.06e-6",
        "mn out in 0 0 ninv w=10e-6 l=0.06e-6",
        "cl out 0 10f",
        ".option klu",
        ".tran 10p 10n",
        qa_card("nmos"), qa_card("pmos"), ".end", ""])


def dio_netlist():
    """rectifier + zener clamp + a diode with sidewall, tunnel and knee parameters: exercises DIOload's
    forward / reverse / breakdown branches, series resistance (internal node) and none, charge storage"""
    return "\n".join([
        "* diode rectifier, zener clamp, sidewall/tunnel diode",
        "vin in 0 sin(0 5 1meg)",
        "d1 in out dmod",
        "c1 out 0 1n",
        "r1 out 0 1k",
        "r2 in z 100",
        "d2 0 z dz",
        "d3 z 0 dsw area=1.5 pj=2",
        "r3 in w 2k",
        "d4 w 0 dnors",
        "d5 0 w dnors off",
        "r4 in u 500",
        "d6 u 0 drec area=2",
        "d7 0 u dssw pj=3",
        ".model dmod d is=1e-14 rs=10 n=1.05 cjo=2p vj=0.7 m=0.45 tt=5n bv=50 ibv=1e-6",
        ".model dz d is=1e-12 rs=5 bv=3.3 ibv=1e-3 cjo=10p nbv=1.2",
        ".model dsw d is=2e-14 rs=2 n=1.1 cjo=1p jsw=1e-13 ns=1.2 cjp=1p php=0.8 mjsw=0.3 ikf=0.05 ikr=0.01 jtun=1e-9 ntun=30 tt=1n",
        ".model dnors d is=5e-15 cjo=0.5p tt=2n",
        ".model drec d is=1e-14 rs=3 n=1.0 isr=1e-11 nr=2 cjo=1p vj=0.8 m=0.4 tt=1n bv=40",
        ".model dssw d level=3 is=1e-14 rs=4 rsw=6 jsw=2e-13 ns=1.3 cjo=1p cjp=2p php=0.75 mjsw=0.35 tt=1n",
        ".option klu",
        ".tran 5n 3u",
        ".end", ""])


def diosh_netlist(selfheat=True, revrec=False):
    """diodes with the thermal terminal `dt` (self-heating: rth0 / cth0, resistance and breakdown temperature coefficients,
    tlev / tlevc mappings re-evaluated at the junction temperature every iteration, dioload.c:317-322) and / or the soft
    reverse-recovery charge node qp (vp, tt: dioload.c:565-582, 836-862), switched hard enough to heat by tens of kelvin"""
    th = lambda t: (f" {t}", " thermal") if selfheat else ("", "")
    rr = " vp=0.4" if revrec else ""
    sh = " rth0={r} cth0={c}" if selfheat else ""
    L = ["* diodes with self-heating / soft reverse recovery",
         "vin in 0 pulse(-8 8 20n 10n 10n 300n 700n)",
         "r1 in a 20",
         "d1 a 0{} dpow area=2{}".format(*th("t1")),
         "r2 in z 60",
         "d2 0 z{} dzen{}".format(*th("t2")),
         "r3 in w 40",
         "d3 w 0{} dsw3 pj=2{}".format(*th("t3")),
         "r4 in u 100",
         "d4 u 0{} dlean{}".format(*th("t4")),
         "d5 0 u dplain",
         ".model dpow d is=2e-13 rs=0.5 n=1.3 cjo=40p vj=0.75 m=0.4 tt=40n bv=60 ibv=1e-5 eg=1.11 xti=3 trs=2e-3 trs2=1e-6 tt1=1e-3 tm1=1e-4 tm2=1e-7" + rr + sh.format(r=60, c="2e-9"),
         ".model dzen d is=1e-12 rs=2 bv=5.1 ibv=1e-3 nbv=1.3 cjo=30p tcv=-3e-4 tlev=1 tlevc=1 cta=2e-4 tpb=1e-3 tt=5n isr=1e-10 nr=2" + rr + sh.format(r=120, c="1e-9"),
         ".model dsw3 d level=3 is=1e-14 rs=1.5 rsw=3 jsw=3e-13 ns=1.25 cjo=5p cjp=4p php=0.8 mjsw=0.33 tt=10n tlev=2 gap1=7.02e-4 gap2=1108 eg=1.16 trs=1e-3 jtun=1e-9 ntun=30 jtunsw=1e-10 ikf=0.5 ikp=0.2 bv=30" + rr + sh.format(r=200, c="5e-10"),
         ".model dlean d is=5e-15 cjo=2p tt=20n" + rr + sh.format(r=400, c="2e-10"),
         ".model dplain d is=5e-15 cjo=2p tt=2n",
         ".option klu",
         ".tran 2n 1.5u",
         ".end", ""]
    return "\n".join(L)


def vbic_netlist():
    """VBIC (level 4): common-emitter stage with the full-featured card of tests/vbic/CEamp.cir (quasi-
    saturation, avalanche, parasitic transistor, knee currents; TD and RTH removed -- excess phase and
    self-heating are outside this path), a pnp follower with a lean card (no series resistances except
    RE: collapsed internal nodes), an emitter-coupled pair with area/m factors; DC operating point and
    PULSE transient"""
    return "\n".join([
        "* VBIC stages: CE amplifier, pnp follower, emitter-coupled pair",
        "vcc vp 0 dc 5",
        "vin in 0 dc 0.7 pulse(0.7 1.1 1n 0.5n 0.5n 6n 14n)",
        "rb in b 1k",
        "rc vp c 1k",
        "q1 c b 0 0 n1",
        "cl c 0 0.2p",
        "q2 0 c e2 vp p1",
        "re2 vp e2 2k",
        "vref r 0 dc 0.9",
        "q3 o1 in t 0 n1 area=2",
        "q4 o2 r t 0 n1 area=2 m=1.5",
        "r5 vp o1 1.5k", "r6 vp o2 1.5k",
        "it t 0 dc 1.2m",
        "c2 o2 0 0.1p",
        ".model n1 npn level=4",
        "+ is=1e-16 ibei=1e-18 iben=5e-15 ibci=2e-17 ibcn=5e-15 isp=1e-15 rcx=10",
        "+ rci=60 rbx=10 rbi=40 re=2 rs=20 rbp=40 vef=10 ver=4 ikf=2e-3 itf=8e-2",
        "+ xtf=20 ikr=2e-4 ikp=2e-4 cje=1e-13 cjc=2e-14 cjep=1e-13 cjcp=4e-13 vo=2",
        "+ gamm=2e-11 hrcf=2 qco=1e-12 avc1=2 avc2=15 tf=10e-12 tr=100e-12",
        "+ cbeo=5e-15 cbco=3e-15 wbe=0.8 ibeip=1e-19 ibenp=1e-16 ibcip=1e-17 ibcnp=1e-15 aje=-0.5 ajc=-0.5 ajs=-0.5",
        ".model p1 pnp level=4",
        "+ is=2e-16 ibei=2e-18 iben=1e-15 ibci=1e-17 ibcn=1e-15 re=3 vef=20 ver=6 ikf=1e-3",
        "+ cje=5e-14 cjc=1e-14 tf=20e-12 tr=200e-12 fc=0.9 qbm=1 nkf=0.6",
        ".option klu",
        ".tran 0.05n 30n",
        ".end", ""])


def mix_netlist(vdd="2.0", r="1k"):
    """BASELINE config 5: one cell with every model family of the path -- an 8-stage BSIM4 inverter chain on
    the swept supply, an RC interconnect with the swept resistor and clamp diodes, a 4-stage BSIM3 level
    shifter and a two-transistor VBIC output stage on 3.3 V; DC operating point + 200-step transient.
    The two swept values are passed as netlist text, exactly what `alter vdd` / `alter r1` would set."""
    lines = ["* mixed-model sweep cell: BSIM4 + BSIM3 + VBIC + diode + R/C",
             f"vdd dd 0 dc {vdd}", "v33 d3 0 dc 3.3",
             "vin in 0 dc 0 pulse(0 2.0 0.2n 0.1n 0.1n 1.2n 3n)"]
    prev = "in"
    for k in range(1, 9):
        lines.append(f"mp{k} a{k} {prev} dd dd p1 l=0.1u w=10u ad=5p pd=6u as=5p ps=6u")
        lines.append(f"mn{k} a{k} {prev} 0 0 n1 l=0.1u w=5u ad=5p pd=6u as=5p ps=6u")
        prev = f"a{k}"
    lines += [f"r1 a8 x {r}", "c1 x 0 30f", "d1 x d3 dcl", "d2 0 x dcl"]
    prev = "x"
    for k in range(1, 5):
        lines.append(f"mnl{k} y{k} {prev} 0 0 nb3 w=2u l=0.35u as=3p ad=3p ps=4u pd=4u")
        lines.append(f"mpl{k} y{k} {prev} d3 d3 pb3 w=4u l=0.35u as=7p ad=7p ps=6u pd=6u")
        prev = f"y{k}"
    lines += ["c2 y4 0 50f",
              "rb1 y4 b1 20k", "q1 cq b1 0 0 nv", "rc1 d3 cq 2k", "q2 d3 cq e2 0 nv area=2", "re2 e2 0 3k", "c3 e2 0 0.1p",
              "r5 d3 m 10k", "d3 m 0 dref", "r6 m b1 200k",
              ".model dcl d is=1e-14 rs=20 n=1.05 cjo=5f vj=0.7 m=0.4 tt=50p bv=12",
              ".model dref d is=5e-15 rs=5 cjo=10f tt=0.1n",
              ".model nv npn level=4",
              "+ is=1e-16 ibei=1e-18 iben=5e-15 ibci=2e-17 ibcn=5e-15 isp=1e-15 rcx=10",
              "+ rci=60 rbx=10 rbi=40 re=2 rs=20 rbp=40 vef=10 ver=4 ikf=2e-3 itf=8e-2",
              "+ xtf=20 ikr=2e-4 ikp=2e-4 cje=1e-13 cjc=2e-14 cjep=1e-13 cjcp=4e-13 vo=2",
              "+ gamm=2e-11 hrcf=2 qco=1e-12 avc1=2 avc2=15 tf=10e-12 tr=100e-12",
              ".option klu", ".tran 25p 5n"]
    b3 = b3_cards().replace(".model n1 nmos", ".model nb3 nmos").replace(".model p1 pmos", ".model pb3 pmos")
    return "\n".join(lines) + "\n" + b3 + "\n" + ro_cards() + "\n.end\n"


# sweep points recorded for parity (value text exactly as it appears in the netlist): centre + 4 corners
MIX_POINTS = [("2.0", "1k"), ("1.6", "200"), ("1.6", "5k"), ("2.4", "200"), ("2.4", "5k"),
              # operating points plain Newton does not reach: the reference goes through dynamic gmin stepping (cktop.c:162)
              ("1.69098", "1348.2"),
              # marginal operating points (66 and 43 Newton iterations in the reference) whose own pivot order differs from
              # the centre's: the batch, bound to the centre's pivot orders, reaches them through gmin stepping instead
              ("1.69412", "2364.7"), ("1.69412", "2383.5")]


def latch_netlist():
    """.nodeset and .ic without UIC: CKTload's row overrides (cktload.c:118-172).  A BSIM4 latch whose state
    is picked by .nodeset (forced during the INITJCT / INITFIX iterations only), an RC ladder whose nodes
    are held by .ic through the whole operating point, one of them next to a source branch; a set pulse
    flips the latch in the transient"""
    mos = "l=0.1u w={w} ad=5p pd=6u as=5p ps=6u"
    return "\n".join([
        "* BSIM4 latch with .nodeset, RC ladder with .ic",
        "vdd dd 0 dc 2.0",
        "vset s 0 dc 0 pulse(0 2 1n 0.1n 0.1n 1n 10n)",
        "mp1 q qb dd dd p1 " + mos.format(w="10u"), "mn1 q qb 0 0 n1 " + mos.format(w="5u"),
        "mp2 qb q dd dd p1 " + mos.format(w="10u"), "mn2 qb q 0 0 n1 " + mos.format(w="5u"),
        "mn3 q s 0 0 n1 " + mos.format(w="20u"),
        "c1 q 0 10f", "c2 qb 0 10f",
        "r1 dd a 10k", "c3 a 0 50f", "r2 a b 5k", "c4 b 0 20f",
        "vm b bm dc 0", "r3 bm 0 100k",
        ".nodeset v(q)=2 v(qb)=0",
        ".ic v(a)=0.5 v(b)=1.2",
        ".option klu", ".tran 20p 4n"]) + "\n" + ro_cards() + "\n.end\n"


def srcs_netlist():
    """every independent-source waveform of the path: PWL (plain, delayed, repeating), EXP, SFFM and AM voltage
    sources and EXP, SFFM, AM and PWL current sources, each into an RC section; a BSIM4 inverter on the PWL input"""
    mos = "l=0.1u w={w} ad=5p pd=6u as=5p ps=6u"
    return "\n".join([
        "* source waveforms: PWL / EXP / SFFM / AM",
        "vdd dd 0 dc 2.0",
        "v1 a 0 pwl(0 0 0.5n 0 0.7n 2 1.5n 2 1.6n 0.4 3n 1.2)",
        "mp1 y a dd dd p1 " + mos.format(w="10u"), "mn1 y a 0 0 n1 " + mos.format(w="5u"), "c1 y 0 20f",
        "v2 b 0 pwl(0 0 0.2n 1 0.5n 1 0.6n 0) r=0.2n td=0.3n",
        "r2 b b2 1k", "c2 b2 0 0.1p",
        "v3 c 0 exp(0 1.5 0.2n 0.3n 1.5n 0.4n)",
        "r3 c c2 2k", "c3 c2 0 50f",
        "v4 d 0 sffm(0.5 0.4 2g 3 0.4g)",
        "r4 d d2 500", "c4 d2 0 20f",
        "v5 e 0 am(0.2 1 0.8 0.5g 3g 0.1n)",
        "r5 e e2 500", "c5 e2 0 20f",
        "i6 0 f exp(0 1m 0.5n 0.2n 2n 0.3n)",
        "r6 f 0 1k", "c6 f 0 0.2p",
        # ISRCload's own SFFM (phases in coefficients 5 and 6, no delay), AM and PWL (no delay / repetition; ISRCaccept
        # sets the next corner only on a breakpoint); the list starts after t = 0
        "i7 0 g sffm(0.2m 1m 2g 3 0.4g 20 30)",
        "r7 g 0 1k", "c7 g 0 50f",
        "i8 0 h am(0.2m 1m 0.8 0.5g 3g 0.1n 10 20)",
        "r8 h 0 1k", "c8 h 0 50f",
        "i9 0 k pwl(0.2n 0 0.6n 1m 1n 1m 1.4n 0.2m 2.5n 0.8m)",
        "r9 k 0 1k", "c9 k 0 0.1p",
        ".option klu", ".tran 10p 4n"]) + "\n" + ro_cards() + "\n.end\n"


def b3_cards():
    """the level-8 (BSIM3v3.3.0) n1/p1 cards of examples/Monte_Carlo/MC_ring.sp"""
    src = open(os.path.join(REF, "examples/Monte_Carlo/MC_ring.sp")).read()
    i = src.index(".model n1 nmos")
    j = src.index(".end", i)
    return src[i:j]


def b3_netlist(stages=5, capmod=None, xpart=None, tran=".tran 0.05n 20n"):
    """BSIM3 ring of inverters (cells of MC_ring.sp: w/l/as/ad/ps/pd as there) kicked by the same
    PULSE source MC_ring.sp uses between input and output, plus an output buffer and load"""
    lines = ["* BSIM3v3.3.0 ring of inverters (MC_ring.sp cells)",
             "vin in out dc 0.5 pulse 0.5 0 0.1n 5n 1 1 1",
             "vdd dd 0 dc 3.3", "vss ss 0 dc 0", "ve sub 0 dc 0", "vpe well 0 dc 3.3"]
    prev = "in"
    for k in range(1, stages + 1):
        nxt = "out" if k == stages else f"n{k}"
        lines.append(f"mn{k} {nxt} {prev} ss sub n1 w=2u l=0.35u as=3p ad=3p ps=4u pd=4u")
        lines.append(f"mp{k} {nxt} {prev} dd well p1 w=4u l=0.35u as=7p ad=7p ps=6u pd=6u")
        prev = nxt
    lines += ["mnb buf out 0 sub n1 w=2u l=0.35u as=3p ad=3p ps=4u pd=4u",
              "mpb buf out dd well p1 w=4u l=0.35u as=7p ad=7p ps=6u pd=6u",
              "cout buf ss 0.2pF", ".option klu", tran]
    cards = b3_cards()
    if capmod is not None:
        extra = f"+capmod={capmod} xpart={xpart}\n"
        cards = cards.replace("+level=8\n", "+level=8\n" + extra)
    return "\n".join(lines) + "\n" + cards + "\n.end\n"


TOX_Z = [-1.5341205443525463, -0.8871465590188759, -0.4887764111146695, -0.15731068461017067,
         0.15731068461017067, 0.4887764111146695, 0.8871465590188759, 1.5341205443525463]   # 8 equal-probability Gaussian bins


def tox_levels(nominal=1.4e-9, sigma=0.03):
    return [nominal * (1.0 + sigma * z) for z in TOX_Z]


def with_toxe(netlist, toxe):
    """both model cards get the same oxide thickness (a per-sample, die-level variation)"""
    out = re.sub(r"toxe\s*=\s*1\.4e-0*9", f"toxe    = {toxe:.17g}", netlist)
    assert out != netlist
    return out


def read_raw(path):
    data = open(path, "rb").read()
    i = data.index(b"Binary:\n")
    head = data[:i].decode(errors="replace")
    nv = int(re.search(r"No. Variables:\s*(\d+)", head).group(1))
    npts = int(re.search(r"No. Points:\s*(\d+)", head).group(1))
    names = []
    vs = head[head.index("Variables:\n") + len("Variables:\n"):]
    for ln in vs.strip().splitlines():
        parts = ln.split()
        if len(parts) >= 3 and parts[0].isdigit():
            names.append(parts[1])
    arr = np.frombuffer(data, dtype=np.float64, count=nv * npts, offset=i + 8).reshape(npts, nv)
    return names, arr


def run(name, netlist, calls, save):
    os.makedirs(TMP, exist_ok=True)
    cir = os.path.join(TMP, name + ".cir")
    open(cir, "w").write(netlist)
    if name in ("ro17", "ro101", "ro17k", "ro17kg", "invg", "vbicsh", "vbicxf", "vbicshxf", "diosh", "diorr", "dioshrr", "inv", "dio", "b3ring", "arr", "vbic", "mix", "latch", "srcs", "mixsrc", "invsrc", "invgmin", "invshunt"):
        # the netlist itself is kept too: the CPU-baseline arm of bench.py feeds it to oracle/_ref/ngspice
        os.makedirs(os.path.join(HERE, "netlists"), exist_ok=True)
        open(os.path.join(HERE, "netlists", name + ".cir"), "w").write(netlist)
    env = dict(os.environ, NGB_DUMP_FLAT=os.path.join(TMP, name + ".flat"),
               NGB_DUMP_TRACE=os.path.join(TMP, name + ".trace"), NGB_DUMP_CALLS=calls,
               NGB_DUMP_STATS=os.path.join(TMP, name + ".stats"))
    raw = os.path.join(TMP, name + ".raw")
    subprocess.run([DUMP, "-b", "-r", raw, cir], env=env, check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    flat = ngt.read(env["NGB_DUMP_FLAT"])
    trace = ngt.read(env["NGB_DUMP_TRACE"])
    stats = json.load(open(env["NGB_DUMP_STATS"]))
    names, arr = read_raw(raw)
    ngt.write(os.path.join(HERE, name + ".flat.ngt"), flat)
    ngt.write(os.path.join(HERE, name + ".trace.ngt.gz"), trace)
    # node name -> equation number from the flat dump
    nn = bytes(flat["node/names_bytes"].astype(np.uint8)).decode().split("\n")
    eq_of = {}
    for ln in nn:
        if ln.strip():
            num, nm = ln.split(" ", 1)
            eq_of[nm.lower()] = int(num)
    cols, eqs = [], []
    for s in save:
        key = s.lower()
        rawname = f"v({key})" if not key.endswith("#branch") else "i(" + key[:-7] + ")"
        cand = [i for i, n in enumerate(names) if n.lower() in (rawname, key)]
        if not cand:
            raise SystemExit(f"{name}: saved vector {s} not in rawfile ({names[:8]}...)")
        cols.append(cand[0]); eqs.append(eq_of[key])
    wave = {"time": arr[:, 0].copy(), "values": arr[:, cols].copy(), "save_eq": np.array(eqs, np.int32),
            "stats": np.array([stats["accepted"], stats["rejected"], stats["numiter"], stats["timepts"],
                               stats["load_calls"], stats.get("op_loads", -1)], np.int32),   # [5]: CKTload calls under MODETRANOP
            "cpu_times": np.array([stats["load_time"], stats["decomp_time"], stats["reorder_time"],
                                   stats["solve_time"], stats["tran_time"]])}
    ngt.write(os.path.join(HERE, name + ".wave.ngt"), wave)
    print(name, "points", arr.shape[0], "stats", stats)


MEAS_CLAUSES = [   # (name, text of the measurement, clauses as (node, kind, count, val, td)): kinds 0 RISE 1 FALL 2 CROSS
    ("tdiff", "TRIG v(18) VAL=0.5 RISE=1 TARG v(18) VAL=0.5 RISE=3", [("18", 0, 1, 0.5, 0.0), ("18", 0, 3, 0.5, 0.0)]),
    ("tfall", "WHEN v(9)=1.0 FALL=2", [("9", 1, 2, 1.0, 0.0)]),
    ("tcross", "WHEN v(2)=1.2 CROSS=5", [("2", 2, 5, 1.2, 0.0)]),
    ("tdel", "WHEN v(18)=0.5 RISE=1 TD=6n", [("18", 0, 1, 0.5, 6e-9)]),
    ("tnone", "WHEN v(18)=0.5 RISE=40", [("18", 0, 40, 0.5, 0.0)]),
]


def run_meas(name, netlist):
    """`.meas tran` results of the stock reference binary (com_measure2.c) for the device-side measurement clauses"""
    os.makedirs(TMP, exist_ok=True)
    cir = os.path.join(TMP, name + "_meas.cir")
    ctl = ".control\nset numdgt=17\nrun\n" + "".join(f"meas tran {nm} {txt}\nprint {nm}\n" for nm, txt, _ in MEAS_CLAUSES) + ".endc\n"
    open(cir, "w").write(netlist.replace(".end\n", ctl + ".end\n"))
    p = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "ngspice"), "-b", cir], capture_output=True, text=True)
    out = {}
    for nm, _, cl in MEAS_CLAUSES:
        m = re.search(rf"^{nm} = (\S+)$", p.stdout, re.M)      # the `print` line carries all 17 digits
        out[nm] = {"value": float(m.group(1)) if m else None,
                   "clauses": [dict(node=c[0], kind=c[1], count=c[2], val=c[3], td=c[4]) for c in cl]}
    json.dump(out, open(os.path.join(HERE, name + ".meas.json"), "w"), indent=1)
    print(name, "meas", {k: v["value"] for k, v in out.items()})


def run_op_only(name, netlist):
    """a run whose operating point FAILS in the reference: circuit, pattern sets and the statistics only -- `op_loads` is
    the number of CKTload calls under MODETRANOP, i.e. the Newton iterations of CKTop's plain NIiter and all its fallbacks"""
    os.makedirs(TMP, exist_ok=True)
    cir = os.path.join(TMP, name + ".cir")
    open(cir, "w").write(netlist)
    open(os.path.join(HERE, "netlists", name + ".cir"), "w").write(netlist)
    env = dict(os.environ, NGB_DUMP_FLAT=os.path.join(TMP, name + ".flat"), NGB_DUMP_TRACE=os.path.join(TMP, name + ".trace"),
               NGB_DUMP_CALLS="0", NGB_DUMP_STATS=os.path.join(TMP, name + ".stats"))
    p = subprocess.run([DUMP, "-b", "-r", os.path.join(TMP, name + ".raw"), cir], env=env, capture_output=True, text=True)
    stats = json.load(open(env["NGB_DUMP_STATS"]))
    log = [ln for ln in (p.stdout + p.stderr).splitlines() if "stepping" in ln]
    ngt.write(os.path.join(HERE, name + ".flat.ngt"), ngt.read(env["NGB_DUMP_FLAT"]))
    ngt.write(os.path.join(HERE, name + ".trace.ngt.gz"), ngt.read(env["NGB_DUMP_TRACE"]))
    ngt.write(os.path.join(HERE, name + ".wave.ngt"),
              {"stats": np.array([stats["accepted"], stats["rejected"], stats["numiter"], stats["timepts"], stats["load_calls"],
                                  stats["op_loads"], stats["ret"]], np.int32)})
    print(name, stats, log)


if __name__ == "__main__":
    which = sys.argv[1:] or ["ro17", "ro101", "ro17k", "ro17mc"]
    if "ro17" in which:
        run("ro17", ro_netlist(17), "0-3,100,101,5000,5001", ["18", "2", "9", "vdd#branch"])
    if "ro17k" in which:
        run("ro17k", ro_netlist(17, tran=".tran .1ns 20ns uic", kick=True), "1,2", ["18", "2", "9", "vdd#branch"])
    if "gear" in which:
        # `.option method=gear` (NIcomCof / NIintegrate / CKTterr GEAR branches: nicomcof.c:52-121, niinteg.c:42-71, cktterr.c:61)
        gear = lambda n: n.replace(".option", ".option method=gear", 1)
        run("ro17kg", gear(ro_netlist(17, tran=".tran .1ns 20ns uic", kick=True)), "1,2", ["18", "2", "9", "vdd#branch"])
        run("invg", gear(inv_netlist()), "0-3", ["out", "in", "vdd#branch", "vin#branch"])
        run("diog", gear(dio_netlist()), "0-3", ["out", "z", "w", "u", "vin#branch"])
        run("b3ringg", gear(b3_netlist(5)), "0-3", ["out", "buf", "n2", "vdd#branch"])
        run("mixg", gear(mix_netlist(*MIX_POINTS[0])), "0-3", ["a8", "x", "y4", "cq", "e2", "vdd#branch", "v33#branch"])
    if "vbicth" in which:
        # VBIC self-heating (RTH / CTH on both cards: thermal node, d/dVrth stamps, DEVlimitlog) and excess phase (TD: the
        # xf1 / xf2 filter nodes), alone and together -- tests/vbic/CEamp.cir's card has both
        base = vbic_netlist()
        # the thermal node is the fifth terminal `dt` (without it RTH is not connected, vbicsetup.c)
        withdt = base.replace("q1 c b 0 0 n1", "q1 c b 0 0 t1 n1").replace("q2 0 c e2 vp p1", "q2 0 c e2 vp t2 p1") \
                     .replace("q3 o1 in t 0 n1 area=2", "q3 o1 in t 0 t3 n1 area=2").replace("q4 o2 r t 0 n1 area=2 m=1.5", "q4 o2 r t 0 t4 n1 area=2 m=1.5")
        assert withdt.count(" t1 ") == 1 and withdt.count(" t4 ") == 1
        sh = withdt.replace("+ is=1e-16 ibei=1e-18", "+ rth=300 cth=1e-9 is=1e-16 ibei=1e-18").replace("+ is=2e-16 ibei=2e-18", "+ rth=500 cth=5e-10 is=2e-16 ibei=2e-18")
        xf = base.replace("+ is=1e-16 ibei=1e-18", "+ td=5e-12 is=1e-16 ibei=1e-18")
        both = sh.replace("+ rth=300 cth=1e-9", "+ td=5e-12 rth=300 cth=1e-9")
        save = ["c", "e2", "o1", "o2", "b", "vcc#branch"]
        run("vbicsh", sh, "0-3", save + ["t1", "t2", "t4"])
        run("vbicxf", xf, "0-3", save)
        run("vbicshxf", both, "0-3", save + ["t1", "t2", "t4"])
    if "diosh" in which:
        save = ["a", "z", "w", "u", "vin#branch"]
        run("diosh", diosh_netlist(True, False), "0-3", save + ["t1", "t2", "t3", "t4"])
        run("diorr", diosh_netlist(False, True), "0-3", save)
        run("dioshrr", diosh_netlist(True, True), "0-3", save + ["t1", "t2", "t3", "t4"])
    if "ro17kmeas" in which:
        run_meas("ro17k", ro_netlist(17, tran=".tran .1ns 20ns uic", kick=True))
    if "ro17mc" in which:
        # four Monte-Carlo samples with per-instance Vth mismatch (delvto ~ N(0, 15 mV), numpy seed 7)
        rng = np.random.default_rng(7)
        dv = rng.normal(0.0, 0.015, size=(4, 34))
        np.save(os.path.join(HERE, "ro17mc.delvto.npy"), dv)
        for i in range(4):
            run(f"ro17mc{i}", ro_netlist(17, tran=".tran .1ns 20ns uic", kick=True, delvto=dv[i]), "1", ["18", "2", "9", "vdd#branch"])
    if "inv" in which:
        run("inv", inv_netlist(), "0-40,100,101,300,301", ["out", "in", "vdd#branch", "vin#branch"])
    if "dio" in which:
        run("dio", dio_netlist(), "0-30,200,201,1000,1001,2000", ["out", "z", "w", "u", "vin#branch"])
    if "vbic" in which:
        run("vbic", vbic_netlist(), "0-40,100,101,300,301,800", ["c", "e2", "o1", "o2", "b", "vcc#branch"])
    if "mix" in which:
        save = ["a8", "x", "y4", "cq", "e2", "vdd#branch", "v33#branch"]
        run("mix", mix_netlist(*MIX_POINTS[0]), "0-40,100,101,300", save)
        for k, (vdd, r) in enumerate(MIX_POINTS[1:], 1):       # corners: waveforms only
            run(f"mix{k}", mix_netlist(vdd, r), "0", save)
            for ext in (".flat.ngt", ".trace.ngt.gz"):
                os.remove(os.path.join(HERE, f"mix{k}" + ext))
    if "invsrc" in which:
        # `.option noopiter` sends CKTop straight to its fallbacks (cktop.c:42-55): with gminsteps=0 that is gillespie_src,
        # otherwise dynamic_gmin.  The pivoting events are the usual four, so these fixtures carry their own pattern sets
        run("invsrc", inv_netlist().replace(".option klu", ".option klu noopiter gminsteps=0"), "0,1", ["out", "in", "vdd#branch", "vin#branch"])
        run("invgmin", inv_netlist().replace(".option klu", ".option klu noopiter"), "0,1", ["out", "in", "vdd#branch", "vin#branch"])
    if "invshunt" in which:
        # `.option gshunt` sets CKTgshunt: dynamic_gmin then ends on
        # MAX(CKTgmin, CKTgshunt) and leaves CKTdiagGmin = CKTgshunt for the whole transient (cktop.c:178, 268)
        run("invshunt", inv_netlist().replace(".option klu", ".option klu noopiter gshunt=1e-9"), "0,1", ["out", "in", "vdd#branch", "vin#branch"])
    if "invfail" in which:
        # tolerances no iteration can meet: the plain NIiter, dynamic_gmin, new_gmin and gillespie_src all fail in turn
        # (cktop.c:62-96) -- the lengths of their ladders are what this pins (the zero-source solve of gillespie_src is the
        # one NIiter that converges: its iterates are exactly zero)
        run_op_only("invfail", inv_netlist().replace(".option klu", ".option klu reltol=1e-15 vntol=1e-20 abstol=1e-22"))
    if "invtstep" in which:
        # tolerances the transient cannot hold: the run is abandoned with "timestep too small" (dctran.c:901-913) after
        # 1044 accepted points; accepted / rejected / iteration counts up to the failure are what this pins
        run_op_only("invtstep", inv_netlist().replace(".option klu", ".option klu reltol=9e-13 vntol=1e-14 abstol=1e-18"))
    if "mixsrc" in which:
        # the operating point of MIX_POINTS[5] with gmin stepping switched off: CKTop goes straight to gillespie_src
        # (cktop.c:87-96, 481-660), which FAILS here after 759 - 101 iterations ("source stepping failed"); the reference
        # then finds the operating point with OPtran, which is not on this path.  Recorded: the circuit, the pivoting
        # factors of the plain NIiter (calls 0, 1), of the zero-source solve (101, 102) and of the transient (759, 760)
        vdd, r = MIX_POINTS[5]
        run("mixsrc", mix_netlist(vdd, r).replace(".option klu", ".option klu gminsteps=0"), "0,1,100-103,758-761",
            ["a8", "x", "y4", "cq", "e2", "vdd#branch", "v33#branch"])
    if "latch" in which:
        run("latch", latch_netlist(), "0-40,100,101,200", ["q", "qb", "a", "b", "vdd#branch", "vm#branch"])
    if "latchns" in which:
        # the same latch with its .nodeset ON the operating point: releasing the nodeset rows changes nothing, the first
        # MODEINITFLOAT iteration passes the node test -- and NIiter still asks for one more (ipass, niiter.c:307-331)
        run("latchns", latch_netlist().replace(".nodeset v(q)=2 v(qb)=0", ".nodeset v(q)=1.88128234 v(qb)=0.0409630224"),
            "0-40,100,101,200", ["q", "qb", "a", "b", "vdd#branch", "vm#branch"])
    if "srcs" in which:
        run("srcs", srcs_netlist(), "0-20,100,101,300", ["y", "b2", "c2", "d2", "e2", "f", "g", "h", "k", "v1#branch"])
    if "b3ring" in which:
        run("b3ring", b3_netlist(5), "0-40,300,301,1000,1001", ["out", "buf", "n2", "vdd#branch"])
    if "arr" in which:
        # 4x4 inverter array of BASELINE config 4 (generator: ngspice-sf-mirror_b200/synth.py)
        import importlib
        synth = importlib.import_module("ngspice-sf-mirror_b200.synth")
        run("arr", synth.inverter_array_netlist(4, 4, ro_cards()), "0-30,100,101,400,401",
            ["out_0_0", "out_3_3", "in_2_1", "vdd#branch"])
    if "arr16" in which:
        # 16x16 array (512 BSIM4, 2 564 unknowns): past what one CTA's shared memory holds, the grid-wide LU's case.  Kept:
        # the circuit (compressed), the pivoting factors of the run, the waveforms -- not the per-call trace (5 MB)
        import importlib
        synth = importlib.import_module("ngspice-sf-mirror_b200.synth")
        run("arr16", synth.inverter_array_netlist(16, 16, ro_cards()), "0-12", ["out_0_0", "out_15_15", "in_2_1", "vdd#branch"])
        flat = ngt.read(os.path.join(HERE, "arr16.flat.ngt")); trace = ngt.read(os.path.join(HERE, "arr16.trace.ngt.gz"))
        pats = {k: v for k, v in trace.items() if "/pat/" in k}
        for k in list(pats):
            pats[k.split("/")[0] + "/mode"] = trace[k.split("/")[0] + "/mode"]
        ngt.write(os.path.join(HERE, "arr16.flat.ngt.gz"), flat)
        ngt.write(os.path.join(HERE, "arr16.pat.ngt.gz"), pats)
        os.remove(os.path.join(HERE, "arr16.flat.ngt")); os.remove(os.path.join(HERE, "arr16.trace.ngt.gz"))
    if "ro17tox" in which:
        # model-parameter mismatch: BSIM4temp results for 8 discrete oxide-thickness levels (what `altermod
        # toxe=...` + CKTtemp produce), plus two complete reference transients with toxe AND delvto mismatch
        levels = tox_levels()
        tabs = {"levels": np.array(levels)}
        for k, tox in enumerate(levels):
            run(f"_tox{k}", with_toxe(ro_netlist(17, tran=".tran .1ns 0.2ns uic", kick=True), tox), "0", ["18"])
            fl = ngt.read(os.path.join(HERE, f"_tox{k}.flat.ngt"))
            tabs[f"mtab{k}"] = fl["b4/mtab"]; tabs[f"ptab{k}"] = fl["b4/ptab"]; tabs[f"inst{k}"] = fl["b4/inst"]
            tabs["prow"] = fl["b4/prow"]
            for ext in (".flat.ngt", ".trace.ngt.gz", ".wave.ngt"):
                os.remove(os.path.join(HERE, f"_tox{k}" + ext))
        ngt.write(os.path.join(HERE, "ro17tox.tables.ngt"), tabs)
        rng = np.random.default_rng(11)
        dv = rng.normal(0.0, 0.015, size=(2, 34)); lev = np.array([1, 6])
        np.save(os.path.join(HERE, "ro17tox.delvto.npy"), dv); np.save(os.path.join(HERE, "ro17tox.level.npy"), lev)
        for i in range(2):
            run(f"ro17tox{i}", with_toxe(ro_netlist(17, tran=".tran .1ns 20ns uic", kick=True, delvto=dv[i]), levels[lev[i]]),
                "1", ["18", "2", "9", "vdd#branch"])
    if "b4temp" in which:
        # BSIM4temp in / out tables (csrc/ngb_b4temp.c restates b4temp.c + b4geo.c + the clamps of b4check.c): the raw model cards
        # and instances the reference's BSIM4temp worked on, and the load tables it left, for cards that walk its branches
        def inv_variant(card_edit=lambda c: c, inst="", opts="", temp=None):
            n = inv_netlist()
            a = n.index(".model"); b = n.rindex(".end")
            cards = card_edit(n[a:b])
            n = n[:a] + cards + n[b:]
            if inst:
                n = re.sub(r"^(m[pn] .*)$", lambda m: m.group(1) + " " + inst, n, flags=re.M)
            n = n.replace(".option klu", ".option klu " + opts + (f" temp={temp}" if temp is not None else ""))
            return n
        def setp(card, **kw):
            out = card
            for k, v in kw.items():
                pat = re.compile(r"(?im)^(\+?\s*" + k + r"\s*=\s*)\S+.*$")
                if pat.search(out):
                    out = pat.sub(lambda m: m.group(1) + str(v), out)
                else:
                    out = re.sub(r"(?im)^(\.model .*)$", lambda m: m.group(1) + f"\n+ {k} = {v}", out)
            return out
        cases = {
            "ro17k": ro_netlist(17, tran=".tran .1ns 0.2ns uic", kick=True),
            "ro17hot": ro_netlist(17, tran=".tran .1ns 0.2ns uic", kick=True).replace(".option xmu", ".option temp=100 xmu"),
            "ro17tox": with_toxe(ro_netlist(17, tran=".tran .1ns 0.2ns uic", kick=True, delvto=np.random.default_rng(3).normal(0, 0.015, 34)), 1.4e-9 * 1.043),
            "inv": inv_variant(),
            "inv_hot_tm1": inv_variant(lambda c: setp(c, tempmod=1, at=2e-4, ua1=1e-3, ub1=-1e-3, uc1=5e-4, ud1=2e-4, prt=1e-3), temp=85),
            "inv_hot_tm2": inv_variant(lambda c: setp(c, tempmod=2, at=2e-4, ua1=1e-3, ub1=-1e-3, uc1=5e-4, ud1=2e-4, prt=1e-3), temp=-20),
            "inv_hot_tm3": inv_variant(lambda c: setp(c, tempmod=3, at=2e-4, ua1=1e-3, ub1=-1e-3, uc1=5e-4, ud1=2e-4, prt=1e-3), temp=125),
            "inv_rds_geo": inv_variant(lambda c: setp(c, rdsmod=1, rsh=7.0), inst="nf=4 rgeomod=3 geomod=5 min=1"),
            "inv_geo9": inv_variant(lambda c: setp(c, rsh=5.0, permod=0), inst="nf=4 rgeomod=1 geomod=9 pd=2e-6 ps=3e-6"),
            "inv_geo2": inv_variant(lambda c: setp(c, rsh=5.0), inst="nf=3 rgeomod=4 geomod=2 nrd=1.5 nrs=0.5 ad=2e-12 as=3e-12"),
            "inv_stress": inv_variant(lambda c: setp(c, ku0=-4e-6, kvsat=0.2, kvth0=-2e-8, stk2=1e-9, steta0=2e-9, tku0=0.1, wpemod=1, kvth0we=0.01, k2we=0.002, ku0we=-0.003),
                                      inst="nf=2 sa=0.4e-6 sb=0.5e-6 sd=0.3e-6 sc=1e-6 mulu0=0.97 delvto=0.011"),
            "inv_rbody2": inv_variant(lambda c: setp(c, rbps0=60.0, rbpd0=70.0, rbsbx0=120.0, rbsby0=110.0, rbdbx0=130.0, rbdby0=90.0, rbpbx0=40.0, rbpby0=45.0),
                                      inst="rbodymod=2 rgatemod=2 ngcon=2 xgw=1e-7"),
            "inv_dio0": inv_variant(lambda c: setp(c, diomod=0, xjbvs=0.5, xjbvd=0.7, bvs=8.0, bvd=9.0)),
            "inv_dio2": inv_variant(lambda c: setp(c, diomod=2, xjbvs=0.5, xjbvd=0.7, bvs=8.0, bvd=9.0, ijthsrev=0.2, ijthdrev=0.3), temp=60),
            "inv_bin": inv_variant(lambda c: setp(c, binunit=1, lvth0=0.004, wvth0=-0.003, pvth0=0.0005, lu0=1e-3, wk2=0.001, lvsat=900.0, pua=1e-12, lnfactor=0.02)),
            "inv_mob3": inv_variant(lambda c: setp(c, mobmod=3, vtl=2.0e5, xn=2.5, lc=5e-9, lambda_=None) if False else setp(c, mobmod=3, vtl=2.0e5, xn=2.5, lc=5e-9)),
            "inv_k1k2": inv_variant(lambda c: setp(c, k1=0.45, k2=-0.02, dvtp4=0.3, dvtp5=0.01, dvtp2=0.02, dvtp3=0.1)),
        }
        out = {}
        for nm, net in cases.items():
            run("_b4t", net, "0", [])
            fl = ngt.read(os.path.join(HERE, "_b4t.flat.ngt"))
            for k in ("b4t/model", "b4t/inst", "b4t/inst_model", "b4t/temp", "b4/mtab", "b4/ptab", "b4/inst", "b4/prow", "opt/vt0"):
                out[f"{nm}/{k}"] = fl[k]
            for ext in (".flat.ngt", ".trace.ngt.gz", ".wave.ngt"):
                os.remove(os.path.join(HERE, "_b4t" + ext))
        out["cases"] = np.frombuffer("\n".join(cases).encode(), dtype=np.uint8).astype(np.int32)
        ngt.write(os.path.join(HERE, "b4temp.tables.ngt.gz"), out)
    if "b3cap" in which:
        # every capMod / xpart combination of BSIM3 on a short 3-stage version of the same ring
        for cm in (0, 1, 2, 3):
            for xp, tag in ((0.0, "0"), (0.5, "5"), (1.0, "1")):
                run(f"b3c{cm}x{tag}", b3_netlist(3, capmod=cm, xpart=xp, tran=".tran 0.05n 3n"), "0-8,40,41,120,121", ["out", "buf"])
    if "ro101" in which:
        run("ro101", ro_netlist(101), "1,2,3000", ["102", "2", "50", "vdd#branch"])
