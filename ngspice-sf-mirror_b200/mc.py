"""Monte-Carlo mismatch on the host side: turns per-instance `delvto` draws into the
instance-parameter table of a batch.

Follows what BSIM4temp does with delvto (src/spicelib/devices/bsim4/b4temp.c:1758-1794):
    vth0 += delvto ; vfb = pParam->vfb + type*delvto ;
    T3 = type*vth0 - vfb - phi ; vtfbphi1 = (NMOS ? 2*T3 : 2.5*T3)+ ; vtfbphi2 = (4*T3)+ ;
    vfbzb = vfbzbfactor + type*vth0
(model-parameter mismatch such as toxe needs the whole of BSIM4temp and is the next step,
SURVEY.md section 8(f) rank 1)."""
import numpy as np


def bsim4_inst_with_delvto(lib, flat, delvto, tables=None):
    """flat: base circuit (delvto = 0); delvto [S][ninst] -> inst table [NI][ninst][S].
    tables = (inst, mtab, ptab) replaces the base circuit's BSIM4temp results (another toxe level)."""
    names = lib.fields["inst"]
    ix = {n: i for i, n in enumerate(names)}
    inst0, mtab0, ptab0 = tables if tables is not None else (flat["b4/inst"], flat["b4/mtab"], flat["b4/ptab"])
    base = np.asarray(inst0, dtype=np.float64)                      # [NI][ninst]
    S, ninst = delvto.shape
    assert ninst == base.shape[1]
    mnames = lib.fields["model"]; pnames = lib.fields["bin"]
    prow = flat["b4/prow"]
    typ = mtab0[prow, mnames.index("type")]                          # [ninst]
    phi = ptab0[prow, pnames.index("phi")]
    out = np.repeat(base[:, :, None], S, axis=2)
    dv = delvto.T                                                    # [ninst][S]
    vth0_b = base[ix["vth0"]][:, None]; vfb_b = base[ix["vfb"]][:, None]
    vfbzbfactor = base[ix["vfbzb"]][:, None] - typ[:, None] * vth0_b
    vth0 = vth0_b + dv
    vfb = vfb_b + typ[:, None] * dv
    T3 = typ[:, None] * vth0 - vfb - phi[:, None]
    out[ix["vth0"]] = vth0
    out[ix["vfb"]] = vfb
    out[ix["vtfbphi1"]] = np.maximum(np.where(typ[:, None] > 0, T3 + T3, 2.5 * T3), 0.0)
    out[ix["vtfbphi2"]] = np.maximum(4.0 * T3, 0.0)
    out[ix["vfbzb"]] = vfbzbfactor + typ[:, None] * vth0
    return out


def bsim4_with_tox_levels(lib, flat, tables, level, delvto):
    """Model-parameter mismatch on top of delvto: every sample s uses the BSIM4temp results of
    oxide-thickness level `level[s]` (tables: tests/golden/ro17tox.tables.ngt -- what `altermod toxe=`
    followed by CKTtemp produces, recorded from the reference for 8 levels).
    Returns (inst [NI][ninst][S], prow_t [ninst*S], mtab [L*rows][NM], ptab [L*rows][NP])."""
    level = np.asarray(level, dtype=np.int64)
    S, ninst = delvto.shape
    L = len(tables["levels"])
    nrows = tables["mtab0"].shape[0]
    prow = np.asarray(flat["b4/prow"], dtype=np.int64)
    inst = np.empty((len(lib.fields["inst"]), ninst, S))
    for k in range(L):
        sel = np.nonzero(level == k)[0]
        if len(sel):
            inst[:, :, sel] = bsim4_inst_with_delvto(lib, flat, delvto[sel], (tables[f"inst{k}"], tables[f"mtab{k}"], tables[f"ptab{k}"]))
    prow_t = (level[None, :] * nrows + prow[:, None]).astype(np.int32).reshape(-1)       # [ninst][S], sample fastest
    mtab = np.concatenate([tables[f"mtab{k}"] for k in range(L)], axis=0)
    ptab = np.concatenate([tables[f"ptab{k}"] for k in range(L)], axis=0)
    return inst, prow_t, mtab, ptab


def bsim4_with_toxe(lib, raw, toxe, delvto, temp=None, vt0=None):
    """Continuous model-parameter mismatch: sample s has oxide thickness toxe[s] (both model cards, a die-level variation)
    and per-instance threshold shifts delvto[s]; its model / bin / instance rows are computed by the library's own BSIM4temp
    (csrc/ngb_b4temp.c) from the nominal raw tables `raw` = {"model", "inst", "inst_model", "temp", "vt0"} (oracle dump keys
    b4t/*).  Returns (inst [NI][ninst][S], prow_t [ninst*S], mtab [R*S][NM], ptab [R*S][NP]) like bsim4_with_tox_levels; R rows
    per sample (one per distinct model and size), numbered r * S + s.  (Measured on B200, profiles/r02_experiments.md: per-sample
    rows cost the load kernel 17 % against rows shared by a warp; storing them field-major, or only the columns that differ
    between samples, was slower than these plain table rows.)"""
    from .b4temp import Bsim4Temp
    T = Bsim4Temp(lib)
    toxe = np.asarray(toxe, dtype=np.float64)
    S, ninst = delvto.shape
    temp = float(raw["temp"]) if temp is None else temp
    vt0 = float(raw["vt0"]) if vt0 is None else vt0
    inst_out = np.empty((len(lib.fields["inst"]), ninst, S))
    prow_t = np.empty((ninst, S), np.int32)
    mt = pt = None
    R = None
    for s in range(S):
        model = np.array(raw["model"], dtype=np.float64, copy=True); inst = np.array(raw["inst"], dtype=np.float64, copy=True)
        T.set_model(model, "toxe", toxe[s])
        inst[:, T.icol["delvto"]] = delvto[s]
        prow, mtab, ptab, itab = T.run(temp, vt0, model, inst, raw["inst_model"])
        if R is None:
            R = mtab.shape[0]
            mt = np.empty((R, S, mtab.shape[1])); pt = np.empty((R, S, ptab.shape[1]))
        assert mtab.shape[0] == R
        inst_out[:, :, s] = itab
        prow_t[:, s] = prow * S + s
        mt[:, s, :] = mtab; pt[:, s, :] = ptab
    return inst_out, prow_t.reshape(-1), mt.reshape(R * S, -1), pt.reshape(R * S, -1)


def group_by_level(level):
    """Batch layout for model-parameter mismatch: positions of the samples ordered by parameter level (stable),
    so that the 32 consecutive samples a warp evaluates read the SAME model / bin rows (one cache line per
    parameter instead of up to one per level).  Returns `order` with batch position p holding draw order[p];
    results come back in batch order, `inverse = np.argsort(order)` puts them in draw order again."""
    return np.argsort(np.asarray(level), kind="stable")


def spice_number(text):
    """the value the reference front end gives a numeric token (digits, '.', e+-NN, scale suffix), which
    is not always the nearest double: INPevaluate accumulates the digits into a double mantissa
    and multiplies by pow(10, exponent) (src/spicelib/parser/inpeval.c:65-201); an integer token
    without suffix is returned as the mantissa itself (:73-81)"""
    import math
    import re
    t = text.strip().lower()
    m_ = re.fullmatch(r"([+-]?)(\d*)(?:\.(\d*))?(?:[ed]([+-]?\d+))?([a-z]*)", t)
    if not m_ or not (m_.group(2) or m_.group(3)):
        raise ValueError(f"not a SPICE number: {text!r}")
    sign = -1.0 if m_.group(1) == "-" else 1.0
    ip, fp, ex, suf = m_.group(2), m_.group(3) or "", m_.group(4), m_.group(5)
    m = 0.0
    for ch in ip + fp:
        m = (10.0 * m + ord(ch)) - 48.0          # `mantis = 10 * mantis + *here - '0'`: the character code is added first (inpeval.c:69), which
                                                 # rounds differently from 10 m + digit once the mantissa passes 2^53 (17-digit tokens)
    if m_.group(3) is None and ex is None and not suf:
        return m * sign
    e = -len(fp) + (int(ex) if ex else 0)
    if suf.startswith("meg"):
        e += 6
    elif suf.startswith("mil"):
        e -= 6
        m *= 25.4
    elif suf:
        e += {"t": 12, "g": 9, "k": 3, "m": -3, "u": -6, "n": -9, "p": -12, "f": -15, "a": -18}.get(suf[0], 0)
    return sign * m * math.pow(10.0, float(e))


def delvto_as_parsed(delvto, fmt="{:.17g}"):
    """delvto draws as the reference sees them after a round trip through netlist text"""
    flat = np.asarray(delvto, dtype=np.float64)
    return np.array([spice_number(fmt.format(v)) for v in flat.ravel()]).reshape(flat.shape)


def draw_delvto(nsamples, ninst, sigma=0.015, seed=1):
    """per-sample, per-instance Vth mismatch ~ N(0, sigma) (seeded, reproducible on any host)"""
    rng = np.random.default_rng(seed)
    return rng.normal(0.0, sigma, size=(nsamples, ninst))


def instance_names(flat, key="b4/names_bytes"):
    """device instance names in the order of the flattened tables (= reference list order)"""
    return bytes(np.asarray(flat[key]).astype(np.uint8)).decode().split("\n")[:-1]
