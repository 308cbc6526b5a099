"""Monte-Carlo mismatch on the host side: turns per-instance `delvto` draws into the
instance-parameter table of a batch.

Follows what BSIM4temp does with delvto (src/spicelib/devices/bsim4/b4temp.c:1758-1794):
    vth0 += delvto ; vfb = pParam->vfb + type*delvto ;
    T3 = type*vth0 - vfb - phi ; vtfbphi1 = (NMOS ? 2*T3 : 2.5*T3)+ ; vtfbphi2 = (4*T3)+ ;
    vfbzb = vfbzbfactor + type*vth0
(model-parameter mismatch such as toxe needs the whole of BSIM4temp and is the next step,
SURVEY.md section 8(f) rank 1)."""
import numpy as np


def bsim4_inst_with_delvto(lib, flat, delvto):
    """flat: base circuit (delvto = 0); delvto [S][ninst] -> inst table [NI][ninst][S]"""
    names = lib.fields["inst"]
    ix = {n: i for i, n in enumerate(names)}
    base = np.asarray(flat["b4/inst"], dtype=np.float64)            # [NI][ninst]
    S, ninst = delvto.shape
    assert ninst == base.shape[1]
    mnames = lib.fields["model"]; pnames = lib.fields["bin"]
    prow = flat["b4/prow"]
    typ = flat["b4/mtab"][prow, mnames.index("type")]                # [ninst]
    phi = flat["b4/ptab"][prow, pnames.index("phi")]
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


def spice_number(text):
    """the value the reference front end gives a plain decimal token (digits, '.', e+-NN), which
    is not always the nearest double: INPevaluate accumulates the digits into a double mantissa
    and multiplies by pow(10, exponent) (src/spicelib/parser/inpeval.c:65-201)"""
    import math
    t = text.strip().lower()
    sign = 1.0
    if t[0] in "+-":
        sign = -1.0 if t[0] == "-" else 1.0
        t = t[1:]
    mant, _, ex = t.partition("e")
    ip, _, fp = mant.partition(".")
    m = 0.0
    for ch in ip + fp:
        m = 10.0 * m + (ord(ch) - 48)
    e = -len(fp) + (int(ex) if ex else 0)
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
