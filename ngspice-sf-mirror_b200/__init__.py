"""ngspice-sf-mirror_b200 -- host-side mirror of the hot-path interface.

Thin ctypes layer over the C ABI of include/ngb200.h (libngb200.so: hand-written sm_100a
kernels + C host code).  The product path has NO CPU fallback: importing works anywhere, but
creating a batch needs the CUDA library and a GPU, and fails loudly otherwise.  The names
mirror the reference entry points they stand in for (CKTload, SMPluFac, SMPsolve, DCtran ...).
"""
import ctypes
import os
import numpy as np
from . import ngt  # noqa: F401
from . import mc  # noqa: F401
from . import parallel  # noqa: F401
from . import sweep  # noqa: F401
from . import b4temp  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NGB200_LIB", os.path.join(_HERE, "libngb200.so"))

_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_dbl_p = ctypes.POINTER(ctypes.c_double)


class NgbError(RuntimeError):
    pass


def _ip(a):
    return a.ctypes.data_as(_c_int_p)


def _dp(a):
    return a.ctypes.data_as(_c_dbl_p)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Library:
    """Loaded libngb200.so (or, in CPU-only tests, the hostsim test double given by path)."""

    def __init__(self, path=None):
        path = path or LIB_PATH
        if not os.path.exists(path):
            raise NgbError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the hot path)")
        self.path = path
        L = self.L = ctypes.CDLL(path)
        L.ngbBackend.restype = ctypes.c_char_p
        L.ngbLastError.restype = ctypes.c_char_p
        L.ngbLaunchCount.restype = ctypes.c_long
        L.ngbBsim4FieldName.restype = ctypes.c_char_p
        L.ngbCircuitCreate.restype = ctypes.c_void_p
        L.ngbBatchCreate.restype = ctypes.c_void_p
        L.ngbBatchArrayBytes.restype = ctypes.c_long
        L.ngbBatchDevPtr.restype = ctypes.c_void_p
        L.ngbTranWaveBytes.restype = ctypes.c_long
        L.ngbTranTicks.restype = ctypes.c_long
        L.ngbTranDevWaves.restype = ctypes.c_void_p
        lay = (ctypes.c_int * 8)()
        L.ngbBsim4Layout(lay)
        self.layout = list(lay)
        self.fields = {}
        for li, key in ((0, "model"), (1, "bin"), (2, "inst"), (3, "node"), (5, "stamp"), (7, "op")):
            n = self.layout[li]
            self.fields[key] = [L.ngbBsim4FieldName(li, i).decode() for i in range(n)]
        dl = (ctypes.c_int * 3)()
        L.ngbDioLayout(dl)
        self.dio_layout = list(dl)                       # parameters, states, stamp rows
        self.fields["dio"] = [L.ngbBsim4FieldName(8, i).decode() for i in range(dl[0])]
        vl = (ctypes.c_int * 5)()
        L.ngbVbicLayout(vl)
        self.vbic_layout = list(vl)                      # parameters, aux, node roles, states, stamp rows
        bl = (ctypes.c_int * 6)()
        L.ngbBsim3Layout(bl)
        self.b3_layout = list(bl)                        # model, bin, instance, node roles, stamp rows, states
        for li, key, n in ((9, "b3model", bl[0]), (10, "b3bin", bl[1]), (11, "b3inst", bl[2])):
            self.fields[key] = [L.ngbBsim4FieldName(li, i).decode() for i in range(n)]

    @property
    def backend(self):
        return self.L.ngbBackend().decode()

    def check(self, rc, what=""):
        if rc != 0:
            raise NgbError(f"{what} failed with code {rc}: {self.L.ngbLastError().decode()}")

    def launch_count(self):
        return int(self.L.ngbLaunchCount())


_default_lib = None


def library(path=None):
    global _default_lib
    if path is not None:
        return Library(path)
    if _default_lib is None:
        _default_lib = Library()
    return _default_lib


DOPT_KEYS = ["reltol", "abstol", "vntol", "chgtol", "trtol", "temp", "vt0", "xmu", "tstep", "tstop",
             "tmax", "tstart", "delmin", "minbreak", "gmin"]
IOPT_KEYS = ["method", "maxorder", "itl4", "itl1", "uic"]


class Circuit:
    """Flattened circuit: what CKTsetup + CKTtemp leave behind (see include/ngb200.h)."""

    def __init__(self, lib, neq, node_type):
        self.lib = lib
        self.neq = int(neq)
        nt = _i32(node_type)
        self.h = ctypes.c_void_p(lib.L.ngbCircuitCreate(self.neq, _ip(nt)))
        self.flat = None

    def __del__(self):
        try:
            if self.h:
                self.lib.L.ngbCircuitDestroy(self.h)
        except Exception:
            pass

    @classmethod
    def from_flat(cls, lib, flat, lu_pattern=None):
        """Build from a flat-circuit dict (oracle dump / fixture / synthetic generator)."""
        sc = ngt.scalar
        c = cls(lib, sc(flat, "meta/neq"), flat["node/type"])
        c.flat = flat
        c.uic = int(sc(flat, "tran/uic", 0))
        d = np.array([sc(flat, "opt/reltol"), sc(flat, "opt/abstol"), sc(flat, "opt/vntol"),
                      sc(flat, "opt/chgtol"), sc(flat, "opt/trtol"), sc(flat, "opt/temp"), sc(flat, "opt/vt0"),
                      sc(flat, "opt/xmu"), sc(flat, "tran/tstep", 0.0), sc(flat, "tran/tstop", 0.0),
                      sc(flat, "tran/tmax", 0.0), sc(flat, "tran/tstart", 0.0), sc(flat, "opt/delmin", 0.0),
                      sc(flat, "opt/minbreak", 0.0), sc(flat, "opt/gmin")], dtype=np.float64)
        i = np.array([sc(flat, "opt/method"), sc(flat, "opt/maxorder"), sc(flat, "opt/itl4"),
                      sc(flat, "opt/itl1"), sc(flat, "tran/uic", 0)], dtype=np.int32)
        lib.check(lib.L.ngbCircuitSetOptions(c.h, _dp(d), _ip(i)), "ngbCircuitSetOptions")
        if "opt/gminsteps" in flat:      # CKTop's fallbacks (cktop.c:62-96); fixtures recorded before the key existed use the defaults
            lib.check(lib.L.ngbCircuitSetOpFallbacks(c.h, int(sc(flat, "opt/gminsteps")), int(sc(flat, "opt/srcsteps")), int(sc(flat, "opt/itl2")),
                                                     ctypes.c_double(float(sc(flat, "opt/gminfactor"))), int(sc(flat, "opt/noopiter", 0)),
                                                     ctypes.c_double(float(sc(flat, "opt/gshunt", 0.0)))), "ngbCircuitSetOpFallbacks")
        if sc(flat, "opt/bypass", 0):
            raise NgbError("CKTbypass != 0 is not supported on this path")
        n = sc(flat, "b4/ninst", 0)
        if n:
            nodes, flags, prow = _i32(flat["b4/nodes"]), _i32(flat["b4/flags"]), _i32(flat["b4/prow"])
            inst, mtab, ptab = _f64(flat["b4/inst"]), _f64(flat["b4/mtab"]), _f64(flat["b4/ptab"])
            assert inst.shape[0] == lib.layout[2] and mtab.shape[1] == lib.layout[0] and ptab.shape[1] == lib.layout[1], \
                "fixture built against different BSIM4 field lists"
            lib.check(lib.L.ngbCircuitAddBsim4(c.h, int(n), _ip(nodes), _ip(flags), _ip(prow), _dp(inst),
                                               int(mtab.shape[0]), _dp(mtab), _dp(ptab)), "ngbCircuitAddBsim4")
        n = sc(flat, "b3/ninst", 0)
        if n:
            inst, mtab, ptab = _f64(flat["b3/inst"]), _f64(flat["b3/mtab"]), _f64(flat["b3/ptab"])
            assert inst.shape[0] == lib.b3_layout[2] and mtab.shape[1] == lib.b3_layout[0] and ptab.shape[1] == lib.b3_layout[1], \
                "fixture built against different BSIM3 field lists"
            lib.check(lib.L.ngbCircuitAddBsim3(c.h, int(n), _ip(_i32(flat["b3/nodes"])), _ip(_i32(flat["b3/flags"])),
                                               _ip(_i32(flat["b3/prow"])), _dp(inst), int(mtab.shape[0]), _dp(mtab), _dp(ptab)),
                      "ngbCircuitAddBsim3")
        n = sc(flat, "res/n", 0)
        if n:
            lib.check(lib.L.ngbCircuitAddResistors(c.h, int(n), _ip(_i32(flat["res/nodes"])), _dp(_f64(flat["res/g"]))),
                      "ngbCircuitAddResistors")
        n = sc(flat, "cap/n", 0)
        if n:
            lib.check(lib.L.ngbCircuitAddCapacitors(c.h, int(n), _ip(_i32(flat["cap/nodes"])), _dp(_f64(flat["cap/par"]))),
                      "ngbCircuitAddCapacitors")
        n = sc(flat, "dio/n", 0)
        if n:
            par = _f64(flat["dio/par"])
            if par.shape[0] < lib.dio_layout[0]:             # fixtures recorded before the raw rows of self-heating existed
                par = np.ascontiguousarray(np.concatenate([par, np.zeros((lib.dio_layout[0] - par.shape[0], par.shape[1]))], axis=0))
            assert par.shape[0] == lib.dio_layout[0], "fixture built against a different diode field list"
            dfl = _i32(flat["dio/flags"])
            dnodes = np.array(np.asarray(flat["dio/nodes"])[[0, 1, 3, 4, 2, 5]])    # the dump's order is pos neg temp pos' posSw' qp
            dnodes[4] = np.where(dfl & 0x800, dnodes[4], 0)
            dnodes[5] = np.where(dfl & 0x1000, dnodes[5], 0)
            lib.check(lib.L.ngbCircuitAddDiodes(c.h, int(n), _ip(_i32(dnodes)), _ip(dfl), _dp(par)), "ngbCircuitAddDiodes")
        n = sc(flat, "vbic/n", 0)
        if n:
            par = _f64(flat["vbic/par"]); aux = _f64(flat["vbic/aux"])
            assert par.shape[0] == lib.vbic_layout[0] and aux.shape[0] == lib.vbic_layout[1]
            vnodes = np.asarray(flat["vbic/nodes"])
            if vnodes.shape[0] < lib.vbic_layout[2]:        # fixtures recorded before the thermal / excess-phase node roles existed
                vnodes = np.concatenate([vnodes, np.zeros((lib.vbic_layout[2] - vnodes.shape[0], vnodes.shape[1]), vnodes.dtype)], axis=0)
            lib.check(lib.L.ngbCircuitAddVbic(c.h, int(n), _ip(_i32(vnodes)), _ip(_i32(flat["vbic/flags"])),
                                              _dp(par), _dp(aux)), "ngbCircuitAddVbic")
        n = sc(flat, "vsrc/n", 0)
        if n:
            lib.check(lib.L.ngbCircuitAddVsources(c.h, int(n), _ip(_i32(flat["vsrc/nodes"])), _ip(_i32(flat["vsrc/fn"])),
                                                  _dp(_f64(flat["vsrc/par"]))), "ngbCircuitAddVsources")
            if "vsrc/pwl_ptr" in flat:
                ptr = _i32(flat["vsrc/pwl_ptr"]); co = _f64(flat["vsrc/pwl"])
                for k in range(int(n)):
                    if ptr[k + 1] > ptr[k]:
                        seg = np.ascontiguousarray(co[ptr[k]:ptr[k + 1]])
                        lib.check(lib.L.ngbCircuitSetVsourcePwl(c.h, k, len(seg), _dp(seg), ctypes.c_double(float(flat["vsrc/pwl_rdelay"][k])),
                                                                int(flat["vsrc/pwl_rep"][k])), "ngbCircuitSetVsourcePwl")
        n = sc(flat, "isrc/n", 0)
        if n:
            lib.check(lib.L.ngbCircuitAddIsources(c.h, int(n), _ip(_i32(flat["isrc/nodes"])), _ip(_i32(flat["isrc/fn"])),
                                                  _dp(_f64(flat["isrc/par"]))), "ngbCircuitAddIsources")
            if "isrc/pwl_ptr" in flat:
                ptr = _i32(flat["isrc/pwl_ptr"]); co = _f64(flat["isrc/pwl"])
                for k in range(int(n)):
                    if ptr[k + 1] > ptr[k]:
                        seg = np.ascontiguousarray(co[ptr[k]:ptr[k + 1]])
                        lib.check(lib.L.ngbCircuitSetIsourcePwl(c.h, k, len(seg), _dp(seg)), "ngbCircuitSetIsourcePwl")
        lib.check(lib.L.ngbCircuitFinalize(c.h), "ngbCircuitFinalize")
        n = sc(flat, "node/nov", 0)
        if n:
            lib.check(lib.L.ngbCircuitSetNodeOverrides(c.h, int(n), _ip(_i32(flat["node/ov_eq"])), _ip(_i32(flat["node/ov_kind"])),
                                                       _dp(_f64(flat["node/ov_val"]))), "ngbCircuitSetNodeOverrides")
        if lu_pattern is not None:
            c.set_lu_pattern(lu_pattern)
        return c

    def pattern(self):
        n, nnz, nrows = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self.lib.check(self.lib.L.ngbCircuitPatternSize(self.h, ctypes.byref(n), ctypes.byref(nnz), ctypes.byref(nrows)))
        Ap = np.zeros(n.value + 1, np.int32); Ai = np.zeros(nnz.value, np.int32); dg = np.zeros(n.value, np.int32)
        self.lib.check(self.lib.L.ngbCircuitGetPattern(self.h, _ip(Ap), _ip(Ai), _ip(dg)))
        return dict(n=n.value, nnz=nnz.value, nrows=nrows.value, Ap=Ap, Ai=Ai, diag=dg)

    def bsim4_slots(self, ninst):
        s = np.zeros((self.lib.layout[4], ninst), np.int32)
        self.lib.check(self.lib.L.ngbCircuitGetBsim4Slots(self.h, _ip(s)))
        return s

    def set_op_fallbacks(self, gminsteps=1, srcsteps=1, itl2=50, gminfactor=10.0, noopiter=0, gshunt=0.0):
        """CKTop's fallbacks after a failed plain NIiter (cktop.c:62-96): `.option gminsteps= srcsteps= itl2= gminfactor=`;
        0 skips the route, 1 is dynamic_gmin + new_gmin / gillespie_src, larger counts are refused (E_UNSUPP); `noopiter` skips
        the plain NIiter"""
        self.lib.check(self.lib.L.ngbCircuitSetOpFallbacks(self.h, int(gminsteps), int(srcsteps), int(itl2), ctypes.c_double(float(gminfactor)), int(noopiter), ctypes.c_double(float(gshunt))),
                       "ngbCircuitSetOpFallbacks")

    def set_lu_pattern(self, pat, prefix="", which=0, uic=None):
        """pat: dict with n nblocks Q R Pnum Lp Li Up Ui Offp Offi (KLU symbolic + numeric pattern).
        A list/tuple gives the pivoting factors of a run in the order the reference computed them
        (NIiter re-pivots in the INITJCT iteration, in the one after it, and in the first two
        iterations of the first time point; a UIC run has only the last two): identical factors share
        a pattern set, and every pivoting event is mapped onto its set (ngbCircuitSetLuEvents)."""
        if isinstance(pat, (list, tuple)):
            if uic is None:
                uic = bool(getattr(self, "uic", 0))
            pats = list(pat)
            keys = ("Pnum", "Q", "Lp", "Li", "Up", "Ui", "Offp", "Offi")
            sets, set_of = [], []
            for p1 in pats:
                for k, q in enumerate(sets):
                    if all(np.array_equal(np.asarray(q[prefix + kk]), np.asarray(p1[prefix + kk])) for kk in keys):
                        set_of.append(k)
                        break
                else:
                    sets.append(p1)
                    set_of.append(len(sets) - 1)
            assert len(sets) <= 4
            for w, p1 in enumerate(sets):
                self.set_lu_pattern(p1, prefix, which=w)
            # events 0..3; a run with fewer factors than events keeps the last one (e.g. recorded runs of one factor)
            if uic:
                ev = [set_of[0], set_of[0]] + [set_of[min(k, len(set_of) - 1)] for k in (0, 1)]
            else:
                ev = [set_of[min(k, len(set_of) - 1)] for k in range(4)]
            self.lib.check(self.lib.L.ngbCircuitSetLuEvents(self.h, _ip(np.array(ev, np.int32))), "ngbCircuitSetLuEvents")
            self.lib.check(self.lib.L.ngbCircuitSelectLuSet(self.h, 0))
            return
        self.lib.check(self.lib.L.ngbCircuitSelectLuSet(self.h, int(which)))
        g = lambda k: _i32(pat[prefix + k])
        self._keep = [g(k) for k in ("Q", "R", "Pnum", "Lp", "Li", "Up", "Ui", "Offp", "Offi")]
        n = int(np.asarray(pat[prefix + "n"]).reshape(-1)[0]); nb = int(np.asarray(pat[prefix + "nblocks"]).reshape(-1)[0])
        if (prefix + "P") in pat:
            # klu_analyze's row permutation travels with the recorded factor: the library can then re-pivot a sample
            # whose refactor meets a zero pivot (ngbCircuitSetSymbolic, csrc/ngb_pivot.c)
            sym = [_i32(pat[prefix + k]) for k in ("P", "Q", "R")]
            self.lib.check(self.lib.L.ngbCircuitSetSymbolic(self.h, n, nb, *[_ip(a) for a in sym]), "ngbCircuitSetSymbolic")
        self.lib.check(self.lib.L.ngbCircuitSetLuPattern(self.h, n, nb, *[_ip(a) for a in self._keep]),
                       "ngbCircuitSetLuPattern")

    def lu_info(self):
        info = (ctypes.c_int * 9)()
        self.lib.check(self.lib.L.ngbCircuitLuInfo(self.h, info))
        return dict(zip(["nV", "nlev", "npairs", "ntask", "nslev", "nsolvepairs", "lnz", "unz", "nzoff"], list(info)))


_INT_ARRAYS = {"ctl.lusel", "ctl.mode", "ctl.active", "ctl.head", "ctl.order", "ctl.noncon", "ctl.xsel", "ctl.err",
               "b4.prow", "lu.nodeconv", "lu.singular", "errflag"}


class Batch:
    """S device-resident samples of one circuit."""

    def __init__(self, circuit, nsamples, device=0):
        self.c = circuit
        self.lib = circuit.lib
        self.S = int(nsamples)
        h = self.lib.L.ngbBatchCreate(circuit.h, self.S, int(device))
        if not h:
            raise NgbError("ngbBatchCreate failed: " + self.lib.L.ngbLastError().decode())
        self.h = ctypes.c_void_p(h)

    def __del__(self):
        try:
            if self.h:
                self.lib.L.ngbBatchDestroy(self.h)
        except Exception:
            pass

    def nbytes(self, name):
        n = self.lib.L.ngbBatchArrayBytes(self.h, name.encode())
        if n < 0:
            raise NgbError(self.lib.L.ngbLastError().decode())
        return n

    def put(self, name, arr, offset=0):
        dt = np.int32 if name in _INT_ARRAYS else np.float64
        a = np.ascontiguousarray(arr, dtype=dt)
        self.lib.check(self.lib.L.ngbBatchUpload(self.h, name.encode(), a.ctypes.data_as(ctypes.c_void_p),
                                                 ctypes.c_long(a.nbytes), ctypes.c_long(offset)), f"upload {name}")

    def get(self, name, shape=None):
        dt = np.int32 if name in _INT_ARRAYS else np.float64
        n = self.nbytes(name) // np.dtype(dt).itemsize
        a = np.zeros(n, dt)
        self.lib.check(self.lib.L.ngbBatchDownload(self.h, name.encode(), a.ctypes.data_as(ctypes.c_void_p),
                                                   ctypes.c_long(a.nbytes), ctypes.c_long(0)), f"download {name}")
        return a.reshape(shape) if shape is not None else a

    def set_resistors(self, g):
        """per-sample conductances g [nres][S] (parameter sweeps)"""
        g = np.ascontiguousarray(g, dtype=np.float64)
        assert g.shape[1] == self.S
        self.lib.check(self.lib.L.ngbBatchSetResistors(self.h, _dp(g)), "ngbBatchSetResistors")

    def bsim4_variant(self):
        """(variant key of the batch's BSIM4 instances, True when the kernel specialised on it is in use) -- csrc/bsim4_variants.h"""
        k = (ctypes.c_uint * 2)()
        self.lib.check(self.lib.L.ngbBatchBsim4Variant(self.h, k), "ngbBatchBsim4Variant")
        return int(k[0]), bool(k[1])

    def bsim4_overlay(self):
        """(model columns, bin columns) read per sample when the rows of set_bsim4_rows are read as an overlay, else None"""
        nm, nb = ctypes.c_int(), ctypes.c_int()
        on = self.lib.L.ngbBatchBsim4Overlay(self.h, ctypes.byref(nm), ctypes.byref(nb))
        return (nm.value, nb.value) if on else None

    def set_bsim4_generic(self, on=True):
        """run the generic BSIM4 load kernel whatever the variant key (same bits; parity tests and measurements)"""
        self.lib.L.ngbBatchSetBsim4Generic(self.h, 1 if on else 0)

    def set_op_full(self, on=True):
        self.lib.L.ngbBatchSetOpFull(self.h, 1 if on else 0)

    # hot path: one call = one step of NIiter for every active sample
    def load(self):
        self.lib.check(self.lib.L.ngbLoad(self.h), "ngbLoad (CKTload)")

    def set_bsim4_rows(self, prow_t, mtab, ptab):
        """per-(instance, sample) model/bin parameter rows: Monte-Carlo with model-parameter mismatch.  mtab [nrows][NM],
        ptab [nrows][NP], prow_t [ninst * S] (sample fastest)"""
        pr, mt, pt = _i32(prow_t), _f64(mtab), _f64(ptab)
        assert mt.shape[1] == self.lib.layout[0] and pt.shape[1] == self.lib.layout[1] and mt.shape[0] == pt.shape[0]
        self.lib.check(self.lib.L.ngbBatchSetBsim4Rows(self.h, _ip(pr), int(mt.shape[0]), _dp(mt), _dp(pt)), "ngbBatchSetBsim4Rows")

    def lufac(self):
        self.lib.check(self.lib.L.ngbLuFac(self.h), "ngbLuFac (SMPluFac)")

    def solve(self):
        self.lib.check(self.lib.L.ngbSolve(self.h), "ngbSolve (SMPsolve)")

    def lufac_solve(self):
        self.lib.check(self.lib.L.ngbLuFacSolve(self.h), "ngbLuFacSolve")

    def set_measures(self, clauses):
        """`.meas tran` clauses evaluated on the device while the points are produced (com_measure2.c:378-663):
        clauses = [(equation, kind 0 RISE / 1 FALL / 2 CROSS, count, level, td), ...]; [] removes them"""
        n = len(clauses)
        eq = _i32([c[0] for c in clauses]); kind = _i32([c[1] for c in clauses]); cnt = _i32([c[2] for c in clauses])
        val = _f64([c[3] for c in clauses]); td = _f64([c[4] for c in clauses])
        self.lib.check(self.lib.L.ngbTranSetMeasures(self.h, n, _ip(eq), _ip(kind), _ip(cnt), _dp(val), _dp(td)), "ngbTranSetMeasures")
        self.nmeas = n

    # DCtran for the whole batch, resident on the device
    def tran(self, max_points, save_eq):
        """Run the transient analysis (options tstep/tstop/tmax/uic of the circuit) for every
        sample; returns a TranResult."""
        se = _i32(save_eq)
        self.lib.check(self.lib.L.ngbTranRun(self.h, int(max_points), _ip(se), int(len(se))), "ngbTranRun (DCtran)")
        return TranResult(self, int(max_points), int(len(se)))


class TranResult:
    def __init__(self, batch, max_points, nsave):
        self.b, self.max_points, self.nsave = batch, max_points, nsave
        S = batch.S
        L = batch.lib.L
        a = [np.zeros(S, np.int32) for _ in range(4)]
        batch.lib.check(L.ngbTranStats(batch.h, *[_ip(v) for v in a]), "ngbTranStats")
        self.accepted, self.rejected, self.numiter, self.npoints = a
        self.ticks = int(L.ngbTranTicks(batch.h))
        self.repivots = int(L.ngbTranRepivots(batch.h))     # samples x events factored with pivoting on the host
        self.err = np.zeros(S, np.int32)          # DCtran's return value per sample (0 = completed)
        batch.lib.check(L.ngbTranErrors(batch.h, _ip(self.err)), "ngbTranErrors")

    def measures(self):
        """[nclauses][S] measured times, NaN where the transition did not occur"""
        out = np.zeros((self.b.nmeas, self.b.S))
        self.b.lib.check(self.b.lib.L.ngbTranMeasures(self.b.h, _dp(out)), "ngbTranMeasures")
        return out

    def write_raw(self, path, names, types=None, title="", date=None, first_sample=0, nsamples=None):
        """binary rawfile with one `Transient Analysis` plot per sample (header and rows as `ngspice -b -r` writes them,
        outitf.c:881-1092); names = the saved equations' vector names in the order given to tran()"""
        n = self.b.S - first_sample if nsamples is None else nsamples
        types = types or ["current" if x.startswith("i(") else "voltage" for x in names]
        assert len(names) == self.nsave and len(types) == self.nsave
        arr = lambda xs: (ctypes.c_char_p * len(xs))(*[x.encode() for x in xs])
        self.b.lib.check(self.b.lib.L.ngbTranWriteRaw(self.b.h, path.encode(), title.encode(), date.encode() if date else None,
                                                      arr(names), arr(types), int(first_sample), int(n)), "ngbTranWriteRaw")

    def waves(self):
        S = self.b.S
        t = np.zeros((S, self.max_points)); v = np.zeros((S, self.max_points, max(self.nsave, 1)))
        self.b.lib.check(self.b.lib.L.ngbTranWaves(self.b.h, _dp(t), _dp(v)), "ngbTranWaves")
        return t, v
