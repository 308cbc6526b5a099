"""BSIM4temp inside the library (csrc/ngb_b4temp.c) against the reference's own BSIM4temp.

tests/golden/b4temp.tables.ngt.gz (make_golden.py b4temp) holds, for 17 card / instance / temperature variants chosen to walk
the branches of b4temp.c and b4geo.c, the raw model cards and instances as the reference's BSIM4temp saw them, and the load
tables (b4/mtab, b4/ptab, b4/inst, b4/prow) it left.  Every entry must come back the same bits.  Because the raw tables were
taken after the reference's run they already hold its results, so the derived fields are poisoned with NaN first, and two
cross-case runs rebuild one case's tables from ANOTHER case's raw inputs (another temperature; another oxide thickness plus
delvto draws)."""
import numpy as np
import pytest
from parity_util import GOLDEN, ngt, pkg

INST_OUT = ["u0temp", "vth0", "vsattemp", "eta0", "k2", "vfb", "vtfbphi1", "vtfbphi2", "vbsc", "k2ox", "vfbzb", "cgso", "cgdo", "grgeltd",
            "Pseff", "Pdeff", "Aseff", "Adeff", "sourceConductance", "drainConductance", "SjctTempRevSatCur", "DjctTempRevSatCur",
            "SswTempRevSatCur", "DswTempRevSatCur", "SswgTempRevSatCur", "DswgTempRevSatCur", "toxp", "coxp"]
MODEL_OUT = ["coxe", "vcrit", "factor1", "vtm0", "Eg0", "vtm", "SjctTempSatCurDensity", "SjctSidewallTempSatCurDensity",
             "SjctGateSidewallTempSatCurDensity", "DjctTempSatCurDensity", "DjctSidewallTempSatCurDensity", "DjctGateSidewallTempSatCurDensity",
             "SunitAreaTempJctCap", "DunitAreaTempJctCap", "SunitLengthSidewallTempJctCap", "DunitLengthSidewallTempJctCap",
             "SunitLengthGateSidewallTempJctCap", "DunitLengthGateSidewallTempJctCap", "PhiBS", "PhiBD", "PhiBSWS", "PhiBSWD", "PhiBSWGS",
             "PhiBSWGD", "njtsstemp", "njtsswstemp", "njtsswgstemp", "njtsdtemp", "njtsswdtemp", "njtsswgdtemp"]


@pytest.fixture(scope="module")
def tables():
    tab = ngt.read(f"{GOLDEN}/b4temp.tables.ngt.gz")
    cases = bytes(tab["cases"].astype(np.uint8)).decode().split("\n")
    return tab, cases


def _get(tab, case, key):
    return tab[f"{case}/{key}"]


def _run_case(T, tab, case, raw_from=None, temp=None, edit=None):
    src = raw_from or case
    model = _get(tab, src, "b4t/model").copy(); inst = _get(tab, src, "b4t/inst").copy()
    for n in MODEL_OUT:
        model[:, T.mcol[n]] = np.nan
    for n in INST_OUT:
        inst[:, T.icol[n]] = np.nan
    if edit:
        edit(model, inst)
    t = float(_get(tab, case, "b4t/temp")[0, 0]) if temp is None else temp
    return T.run(t, float(_get(tab, case, "opt/vt0")[0]), model, inst, _get(tab, src, "b4t/inst_model"))


def _assert_same(tab, case, got):
    prow, mtab, ptab, itab = got
    assert np.array_equal(prow, _get(tab, case, "b4/prow"))
    for a, key in ((mtab, "b4/mtab"), (ptab, "b4/ptab"), (itab, "b4/inst")):
        b = _get(tab, case, key)
        assert a.shape == b.shape and np.array_equal(a, b), (case, key, np.argwhere(a != b)[:5])


def test_every_case_bit_identical(hostsim_lib, tables):
    tab, cases = tables
    T = pkg.b4temp.Bsim4Temp(hostsim_lib)
    assert len(cases) >= 17
    for case in cases:
        _assert_same(tab, case, _run_case(T, tab, case))


def test_variants_differ_from_their_base(tables):
    """the card / instance edits of make_golden.py took effect: every variant's tables differ from the plain inverter's"""
    tab, cases = tables
    for case in cases:
        if case.startswith("inv_"):
            same = all(np.array_equal(_get(tab, case, k), _get(tab, "inv", k)) for k in ("b4/mtab", "b4/ptab", "b4/inst"))
            assert not same, case


def test_other_temperature_from_the_same_raw_card(hostsim_lib, tables):
    """the 27 C oscillator's raw tables at the hot run's temperature give the hot run's tables"""
    tab, _ = tables
    T = pkg.b4temp.Bsim4Temp(hostsim_lib)
    hot = float(_get(tab, "ro17hot", "b4t/temp")[0, 0])
    assert hot != float(_get(tab, "ro17k", "b4t/temp")[0, 0])
    _assert_same(tab, "ro17hot", _run_case(T, tab, "ro17hot", raw_from="ro17k", temp=hot))


def test_continuous_toxe_and_delvto_from_the_nominal_card(hostsim_lib, tables):
    """model-parameter mismatch the reference was NOT asked for beforehand: the nominal card with toxe and per-instance delvto
    set through the library gives the tables of the reference run on the edited netlist"""
    tab, _ = tables
    T = pkg.b4temp.Bsim4Temp(hostsim_lib)
    want_m = _get(tab, "ro17tox", "b4t/model"); want_i = _get(tab, "ro17tox", "b4t/inst")

    def edit(model, inst):
        T.set_model(model, "toxe", want_m[0, T.mcol["toxe"]])
        inst[:, T.icol["delvto"]] = want_i[:, T.icol["delvto"]]
    _assert_same(tab, "ro17tox", _run_case(T, tab, "ro17tox", raw_from="ro17k", edit=edit))


def test_fatal_parameter_is_reported(hostsim_lib, tables):
    tab, _ = tables
    T = pkg.b4temp.Bsim4Temp(hostsim_lib)

    def edit(model, inst):
        model[:, T.mcol["Lint"]] = 1e-6          # effective channel length <= 0
    with pytest.raises(pkg.NgbError):
        _run_case(T, tab, "inv", edit=edit)
