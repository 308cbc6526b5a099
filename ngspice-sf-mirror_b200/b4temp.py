"""BSIM4temp inside the library (csrc/ngb_b4temp.c): the BSIM4 load's model / bin / instance tables from raw model cards and
instance geometry, for any parameter value -- continuous model-parameter mismatch.

The raw tables are what oracle/ref_hooks.c dumps as b4t/model, b4t/inst, b4t/inst_model (one row per model card / instance,
columns by the name lists of csrc/bsim4_temp_fields.h)."""
import ctypes
import numpy as np

_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_dbl_p = ctypes.POINTER(ctypes.c_double)


class Bsim4Temp:
    def __init__(self, lib):
        self.lib = lib
        L = lib.L
        L.ngbBsim4TempFieldName.restype = ctypes.c_char_p
        lay = (ctypes.c_int * 4)()
        L.ngbBsim4TempLayout(lay)
        self.nm, self.ns, self.ni, self.nbinned = list(lay)
        self.model_fields = [L.ngbBsim4TempFieldName(0, i).decode() for i in range(self.nm)]
        self.inst_fields = [L.ngbBsim4TempFieldName(2, i).decode() for i in range(self.ni)]
        self.mcol = {n: i for i, n in enumerate(self.model_fields)}
        self.icol = {n: i for i, n in enumerate(self.inst_fields)}

    def set_model(self, model, name, value, card=None):
        """model [nmodel][NM]: set parameter `name` of card `card` (all cards when None) the way a netlist would: the value,
        its Given flag, and the defaults BSIM4setup derives from it (b4set.c:235-240: toxp, toxm follow toxe unless given)"""
        rows = range(model.shape[0]) if card is None else [card]
        for r in rows:
            model[r, self.mcol[name]] = value
            if name + "Given" in self.mcol:
                model[r, self.mcol[name + "Given"]] = 1.0
            if name == "toxe":
                if not model[r, self.mcol["toxpGiven"]]:
                    model[r, self.mcol["toxp"]] = value
                if not model[r, self.mcol["toxmGiven"]]:
                    model[r, self.mcol["toxm"]] = value
        return model

    def run(self, temp, vt0, model, inst, inst_model):
        """returns (prow [ninst], mtab [nrows][78], ptab [nrows][143], itab [51][ninst]); model and inst are updated in place
        like the reference's structures"""
        model = np.ascontiguousarray(model, dtype=np.float64); inst = np.ascontiguousarray(inst, dtype=np.float64)
        im = np.ascontiguousarray(inst_model, dtype=np.int32)
        ninst = inst.shape[0]
        assert model.shape[1] == self.nm and inst.shape[1] == self.ni
        lay = self.lib.layout
        prow = np.zeros(ninst, np.int32); nrows = ctypes.c_int(0)
        mtab = np.zeros((ninst, lay[0])); ptab = np.zeros((ninst, lay[1])); itab = np.zeros((lay[2], ninst))
        self.lib.check(self.lib.L.ngbBsim4Temp(ctypes.c_double(float(temp)), ctypes.c_double(float(vt0)), int(model.shape[0]),
                                               model.ctypes.data_as(_c_dbl_p), ninst, im.ctypes.data_as(_c_int_p),
                                               inst.ctypes.data_as(_c_dbl_p), prow.ctypes.data_as(_c_int_p), ctypes.byref(nrows),
                                               mtab.ctypes.data_as(_c_dbl_p), ptab.ctypes.data_as(_c_dbl_p), itab.ctypes.data_as(_c_dbl_p)),
                       "ngbBsim4Temp")
        n = nrows.value
        return prow, mtab[:n].copy(), ptab[:n].copy(), itab
