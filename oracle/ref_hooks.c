/* oracle/ref_hooks.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Linked into oracle/_ref/ngspice_dump together with the UNMODIFIED reference objects
 * (ld --wrap, no reference source is touched).  It records, from inside a real ngspice
 * run, the data the parity tests need:
 *
 *   NGB_DUMP_FLAT=<file>   after CKTsetup+CKTtemp (first CKTload): the flattened circuit --
 *                          equation count, node types, options, the KLU CSC pattern built by
 *                          SMPconvertCOOtoCSC, and for every BSIM4/RES/CAP/VSRC/ISRC instance
 *                          its node numbers, CSC slots (pointer - Ax) and the parameter values
 *                          BSIM4setup/BSIM4temp produced, in the field order of
 *                          ngspice-sf-mirror_b200/csrc/bsim4_fields.h.
 *   NGB_DUMP_TRACE=<file>  for the CKTload calls listed in NGB_DUMP_CALLS ("0-9,40,41" ...):
 *                          inputs (mode, time, ag, rhsOld, state vectors) and outputs
 *                          (Ax/rhs after the BSIM4 load alone and after the whole CKTload,
 *                          state0, noncon, op-point), plus the KLU pattern after each pivoting
 *                          factor, LU values after each (re)factor, and the SMPsolve result.
 *   NGB_DUMP_STATS=<file>  STATaccepted/STATrejected/STATnumIter/STATtimePts and the load /
 *                          decomposition / solve times when DCtran returns.
 *
 * File format "NGT1": magic, then records { u16 name_len, name, u8 dtype ('d' f64 | 'i' i32),
 * u8 ndim, i64 dims[ndim], raw little-endian data }.
 */
#include "ngspice/ngspice.h"
#include "ngspice/cktdefs.h"
#include "ngspice/devdefs.h"
#include "ngspice/smpdefs.h"
#include "ngspice/sperror.h"
#include "ngspice/trandefs.h"
#include "ngspice/klu.h"
#include "bsim4/bsim4def.h"
#include "res/resdefs.h"
#include "cap/capdefs.h"
#include "vsrc/vsrcdefs.h"
#include "isrc/isrcdefs.h"
#include "dio/diodefs.h"
#include "vbic/vbicdefs.h"
#include "bsim3/bsim3def.h"
#include "klu_internal.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../ngspice-sf-mirror_b200/csrc/bsim4_fields.h"
#include "../ngspice-sf-mirror_b200/csrc/bsim4_temp_fields.h"
#include "../ngspice-sf-mirror_b200/csrc/dio_fields.h"
#include "../ngspice-sf-mirror_b200/csrc/vbic_types.h"
#include "../ngspice-sf-mirror_b200/csrc/bsim3_fields.h"

extern SPICEdev **DEVices;
extern int DEVmaxnum;
int CKTtypelook(char *);

int __real_CKTload(CKTcircuit *ckt);
void __real_SMPsolve(SMPmatrix *, double[], double[]);
int __real_SMPluFac(SMPmatrix *, double, double);
int __real_SMPreorder(SMPmatrix *, double, double, double);
int __real_DCtran(CKTcircuit *ckt, int restart);

/* ---------------------------------------------------------------- NGT writer */
static FILE *ngt_open(const char *path)
{
    FILE *f = fopen(path, "wb");
    if (!f) { perror(path); exit(2); }
    fwrite("NGT1", 1, 4, f);
    return f;
}
static void ngt_put(FILE *f, const char *name, char dtype, int ndim, const long long *dims, const void *data)
{
    unsigned short nl = (unsigned short)strlen(name);
    unsigned char nd = (unsigned char)ndim;
    size_t n = 1, esz = (dtype == 'd') ? 8 : 4;
    int i;
    fwrite(&nl, 2, 1, f); fwrite(name, 1, nl, f);
    fwrite(&dtype, 1, 1, f); fwrite(&nd, 1, 1, f);
    for (i = 0; i < ndim; i++) { fwrite(&dims[i], 8, 1, f); n *= (size_t)dims[i]; }
    if (n) fwrite(data, esz, n, f);
}
static void put_d1(FILE *f, const char *name, const double *v, long long n)
{ ngt_put(f, name, 'd', 1, &n, v); }
static void put_i1(FILE *f, const char *name, const int *v, long long n)
{ ngt_put(f, name, 'i', 1, &n, v); }
static void put_d2(FILE *f, const char *name, const double *v, long long a, long long b)
{ long long d[2]; d[0] = a; d[1] = b; ngt_put(f, name, 'd', 2, d, v); }
static void put_i2(FILE *f, const char *name, const int *v, long long a, long long b)
{ long long d[2]; d[0] = a; d[1] = b; ngt_put(f, name, 'i', 2, d, v); }
static void put_ds(FILE *f, const char *name, double v) { put_d1(f, name, &v, 1); }
static void put_is(FILE *f, const char *name, int v) { put_i1(f, name, &v, 1); }

/* ---------------------------------------------------------------- helpers */
static KLUmatrix *klu_of(CKTcircuit *ckt)
{
    if (!ckt->CKTmatrix || !ckt->CKTmatrix->CKTkluMODE) return NULL;
    return ckt->CKTmatrix->SMPkluMatrix;
}
static int slot_of(KLUmatrix *K, double *p)
{
    if (!K || !p) return -1;
    if (p >= K->KLUmatrixAx && p < K->KLUmatrixAx + K->KLUmatrixNZ) return (int)(p - K->KLUmatrixAx);
    return -1;     /* trash cell (ground row/column) */
}

static int b4_type = -2, res_type, cap_type, vsrc_type, isrc_type, dio_type, b3_type, vbic_type;
static void lookup_types(void)
{
    if (b4_type != -2) return;
    b4_type = CKTtypelook("BSIM4");
    res_type = CKTtypelook("Resistor");
    cap_type = CKTtypelook("Capacitor");
    vsrc_type = CKTtypelook("Vsource");
    isrc_type = CKTtypelook("Isource");
    dio_type = CKTtypelook("Diode");
    vbic_type = CKTtypelook("VBIC");
    b3_type = CKTtypelook("BSIM3");
}

/* ---------------------------------------------------------------- flat circuit dump */
typedef struct { BSIM4model *m; struct bsim4SizeDependParam *p; } prow_t;

static void dump_bsim4(FILE *f, CKTcircuit *ckt, KLUmatrix *K)
{
    BSIM4model *model;
    BSIM4instance *here;
    int n = 0, nrows = 0, i, r;
    prow_t *rows;
    int *nodes, *slots, *flags, *sbase, *prow;
    double *inst, *mtab, *ptab;
    char *names; size_t nameslen = 0, namescap = 1 << 16;

    if (b4_type < 0 || !ckt->CKThead[b4_type]) { put_is(f, "b4/ninst", 0); return; }
    for (model = (BSIM4model *)ckt->CKThead[b4_type]; model; model = BSIM4nextModel(model))
        for (here = BSIM4instances(model); here; here = BSIM4nextInstance(here)) n++;
    rows = (prow_t *)calloc((size_t)n, sizeof(prow_t));
    nodes = (int *)calloc((size_t)n * B4N_COUNT, sizeof(int));
    slots = (int *)calloc((size_t)n * B4S_MAT_COUNT, sizeof(int));
    flags = (int *)calloc((size_t)n, sizeof(int));
    sbase = (int *)calloc((size_t)n, sizeof(int));
    prow = (int *)calloc((size_t)n, sizeof(int));
    inst = (double *)calloc((size_t)n * B4I_COUNT, sizeof(double));
    names = (char *)malloc(namescap);

    i = 0;
    for (model = (BSIM4model *)ckt->CKThead[b4_type]; model; model = BSIM4nextModel(model))
        for (here = BSIM4instances(model); here; here = BSIM4nextInstance(here), i++) {
            int k = 0;
            size_t l = strlen(here->BSIM4name);
            if (nameslen + l + 2 > namescap) { namescap *= 2; names = (char *)realloc(names, namescap); }
            memcpy(names + nameslen, here->BSIM4name, l); nameslen += l; names[nameslen++] = '\n';
#define X(nm) nodes[(k++) * n + i] = here->BSIM4##nm;
            NGB_B4_NODE_FIELDS(X)
#undef X
            k = 0;
#define X(nm) slots[(k++) * n + i] = slot_of(K, here->BSIM4##nm##Ptr);
            NGB_B4_MAT_FIELDS(X)
#undef X
            k = 0;
#define X(nm) inst[(size_t)(k++) * n + i] = (double)here->BSIM4##nm;
            NGB_B4_INST_FIELDS(X)
#undef X
            flags[i] = (here->BSIM4off ? B4F_OFF : 0) | ((here->BSIM4rbodyMod & 3) << B4F_RBODY_SH)
                     | ((here->BSIM4rgateMod & 3) << B4F_RGATE_SH)
                     | (here->BSIM4trnqsMod ? 0x100 : 0) | (here->BSIM4acnqsMod ? 0x200 : 0);
            sbase[i] = here->BSIM4states;
            for (r = 0; r < nrows; r++)
                if (rows[r].m == model && rows[r].p == here->pParam) break;
            if (r == nrows) { rows[nrows].m = model; rows[nrows].p = here->pParam; nrows++; }
            prow[i] = r;
        }
    mtab = (double *)calloc((size_t)nrows * B4M_COUNT, sizeof(double));
    ptab = (double *)calloc((size_t)nrows * B4P_COUNT, sizeof(double));
    for (r = 0; r < nrows; r++) {
        int k = 0;
        BSIM4model *m = rows[r].m;
        struct bsim4SizeDependParam *pParam = rows[r].p;
#define X(nm) mtab[(size_t)r * B4M_COUNT + (k++)] = (double)m->BSIM4##nm;
        NGB_B4_MODEL_FIELDS(X)
#undef X
        k = 0;
#define X(nm) ptab[(size_t)r * B4P_COUNT + (k++)] = (double)pParam->BSIM4##nm;
        NGB_B4_BIN_FIELDS(X)
#undef X
    }
    {   /* the raw tables BSIM4temp works on (csrc/ngb_b4temp.c restates it): every model-card and instance quantity it reads or
         * writes, by the generated name lists.  Taken after CKTtemp: the inputs are unchanged by it, re-running it is idempotent */
        int nmodel = 0, mi = 0, *imodel = (int *)calloc((size_t)n, sizeof(int));
        double *tm, *ti;
        for (model = (BSIM4model *)ckt->CKThead[b4_type]; model; model = BSIM4nextModel(model)) nmodel++;
        tm = (double *)calloc((size_t)nmodel * B4TM_COUNT, sizeof(double));
        ti = (double *)calloc((size_t)n * B4TI_COUNT, sizeof(double));
        i = 0;
        for (model = (BSIM4model *)ckt->CKThead[b4_type]; model; model = BSIM4nextModel(model), mi++) {
            int k = 0;
#define X(nm) tm[(size_t)mi * B4TM_COUNT + (k++)] = (double)model->BSIM4##nm;
            NGB_B4T_MODEL_FIELDS(X)
#undef X
            for (here = BSIM4instances(model); here; here = BSIM4nextInstance(here), i++) {
                k = 0;
                imodel[i] = mi;
#define X(nm) ti[(size_t)i * B4TI_COUNT + (k++)] = (double)here->BSIM4##nm;
                NGB_B4T_INST_FIELDS(X)
#undef X
            }
        }
        put_d2(f, "b4t/model", tm, nmodel, B4TM_COUNT);
        put_d2(f, "b4t/inst", ti, n, B4TI_COUNT);
        put_i1(f, "b4t/inst_model", imodel, n);
        { double tk = ckt->CKTtemp; put_d2(f, "b4t/temp", &tk, 1, 1); }
        free(tm); free(ti); free(imodel);
    }
    put_is(f, "b4/ninst", n);
    put_i2(f, "b4/nodes", nodes, B4N_COUNT, n);
    put_i2(f, "b4/slots", slots, B4S_MAT_COUNT, n);
    put_d2(f, "b4/inst", inst, B4I_COUNT, n);
    put_i1(f, "b4/flags", flags, n);
    put_i1(f, "b4/state_base", sbase, n);
    put_i1(f, "b4/prow", prow, n);
    put_d2(f, "b4/mtab", mtab, nrows, B4M_COUNT);
    put_d2(f, "b4/ptab", ptab, nrows, B4P_COUNT);
    { /* names as int-coded bytes keeps the reader trivial */
      int *nb = (int *)malloc(sizeof(int) * (nameslen ? nameslen : 1)); size_t q;
      for (q = 0; q < nameslen; q++) nb[q] = (unsigned char)names[q];
      put_i1(f, "b4/names_bytes", nb, (long long)nameslen); free(nb); }
    free(rows); free(nodes); free(slots); free(flags); free(sbase); free(prow);
    free(inst); free(mtab); free(ptab); free(names);
}

static void put_names(FILE *f, const char *key, GENmodel *head);
static void dump_linear(FILE *f, CKTcircuit *ckt, KLUmatrix *K)
{
    int n, i;
    /* resistors */
    n = 0;
    if (res_type >= 0) {
        RESmodel *m; RESinstance *h;
        for (m = (RESmodel *)ckt->CKThead[res_type]; m; m = RESnextModel(m))
            for (h = RESinstances(m); h; h = RESnextInstance(h)) n++;
        if (n) {
            int *nodes = (int *)calloc((size_t)n * 2, sizeof(int)), *slots = (int *)calloc((size_t)n * 4, sizeof(int));
            double *g = (double *)calloc((size_t)n, sizeof(double));
            i = 0;
            for (m = (RESmodel *)ckt->CKThead[res_type]; m; m = RESnextModel(m))
                for (h = RESinstances(m); h; h = RESnextInstance(h), i++) {
                    nodes[i] = h->RESposNode; nodes[n + i] = h->RESnegNode;
                    slots[i] = slot_of(K, h->RESposPosPtr); slots[n + i] = slot_of(K, h->RESnegNegPtr);
                    slots[2 * n + i] = slot_of(K, h->RESposNegPtr); slots[3 * n + i] = slot_of(K, h->RESnegPosPtr);
                    g[i] = h->RESconduct;
                }
            put_i2(f, "res/nodes", nodes, 2, n); put_i2(f, "res/slots", slots, 4, n); put_d1(f, "res/g", g, n);
            free(nodes); free(slots); free(g);
        }
    }
    put_is(f, "res/n", n);
    if (n) put_names(f, "res/names_bytes", ckt->CKThead[res_type]);
    /* capacitors */
    n = 0;
    if (cap_type >= 0) {
        CAPmodel *m; CAPinstance *h;
        for (m = (CAPmodel *)ckt->CKThead[cap_type]; m; m = CAPnextModel(m))
            for (h = CAPinstances(m); h; h = CAPnextInstance(h)) n++;
        if (n) {
            int *nodes = (int *)calloc((size_t)n * 2, sizeof(int)), *slots = (int *)calloc((size_t)n * 4, sizeof(int));
            int *sb = (int *)calloc((size_t)n, sizeof(int));
            double *par = (double *)calloc((size_t)n * 3, sizeof(double));
            i = 0;
            for (m = (CAPmodel *)ckt->CKThead[cap_type]; m; m = CAPnextModel(m))
                for (h = CAPinstances(m); h; h = CAPnextInstance(h), i++) {
                    nodes[i] = h->CAPposNode; nodes[n + i] = h->CAPnegNode;
                    slots[i] = slot_of(K, h->CAPposPosPtr); slots[n + i] = slot_of(K, h->CAPnegNegPtr);
                    slots[2 * n + i] = slot_of(K, h->CAPposNegPtr); slots[3 * n + i] = slot_of(K, h->CAPnegPosPtr);
                    par[i] = h->CAPcapac; par[n + i] = h->CAPm; par[2 * n + i] = h->CAPinitCond;
                    sb[i] = h->CAPstate;
                }
            put_i2(f, "cap/nodes", nodes, 2, n); put_i2(f, "cap/slots", slots, 4, n);
            put_d2(f, "cap/par", par, 3, n); put_i1(f, "cap/state_base", sb, n);
            free(nodes); free(slots); free(par); free(sb);
        }
    }
    put_is(f, "cap/n", n);
    /* voltage sources */
    n = 0;
    if (vsrc_type >= 0) {
        VSRCmodel *m; VSRCinstance *h;
        for (m = (VSRCmodel *)ckt->CKThead[vsrc_type]; m; m = VSRCnextModel(m))
            for (h = VSRCinstances(m); h; h = VSRCnextInstance(h)) n++;
        if (n) {
            int *nodes = (int *)calloc((size_t)n * 3, sizeof(int)), *slots = (int *)calloc((size_t)n * 4, sizeof(int));
            int *fn = (int *)calloc((size_t)n * 3, sizeof(int));
            double *par = (double *)calloc((size_t)n * 9, sizeof(double));
            int k;
            i = 0;
            for (m = (VSRCmodel *)ckt->CKThead[vsrc_type]; m; m = VSRCnextModel(m))
                for (h = VSRCinstances(m); h; h = VSRCnextInstance(h), i++) {
                    nodes[i] = h->VSRCposNode; nodes[n + i] = h->VSRCnegNode; nodes[2 * n + i] = h->VSRCbranch;
                    slots[i] = slot_of(K, h->VSRCposIbrPtr); slots[n + i] = slot_of(K, h->VSRCnegIbrPtr);
                    slots[2 * n + i] = slot_of(K, h->VSRCibrPosPtr); slots[3 * n + i] = slot_of(K, h->VSRCibrNegPtr);
                    fn[i] = h->VSRCfunctionType; fn[n + i] = h->VSRCfunctionOrder; fn[2 * n + i] = h->VSRCdcGiven;
                    par[i] = h->VSRCdcValue;
                    for (k = 0; k < 8; k++)
                        par[(size_t)(1 + k) * n + i] = (h->VSRCcoeffs && k < h->VSRCfunctionOrder) ? h->VSRCcoeffs[k] : 0.0;
                }
            put_i2(f, "vsrc/nodes", nodes, 3, n); put_i2(f, "vsrc/slots", slots, 4, n);
            put_i2(f, "vsrc/fn", fn, 3, n); put_d2(f, "vsrc/par", par, 9, n);
            {   /* PWL corner lists */
                int *ptr = (int *)calloc((size_t)n + 1, sizeof(int)), *rep = (int *)calloc((size_t)n, sizeof(int)), tot = 0, any = 0;
                double *rd = (double *)calloc((size_t)n, sizeof(double)), *co;
                i = 0;
                for (m = (VSRCmodel *)ckt->CKThead[vsrc_type]; m; m = VSRCnextModel(m))
                    for (h = VSRCinstances(m); h; h = VSRCnextInstance(h), i++) {
                        ptr[i] = tot; rep[i] = -1;
                        if (h->VSRCfunctionType == PWL) { tot += h->VSRCfunctionOrder; any = 1; rd[i] = h->VSRCrdelay; if (h->VSRCrGiven) rep[i] = h->VSRCrBreakpt; }
                    }
                ptr[n] = tot;
                if (any) {
                    co = (double *)calloc((size_t)tot + 1, sizeof(double)); i = 0;
                    for (m = (VSRCmodel *)ckt->CKThead[vsrc_type]; m; m = VSRCnextModel(m))
                        for (h = VSRCinstances(m); h; h = VSRCnextInstance(h), i++)
                            if (h->VSRCfunctionType == PWL) memcpy(co + ptr[i], h->VSRCcoeffs, sizeof(double) * (size_t)h->VSRCfunctionOrder);
                    put_i1(f, "vsrc/pwl_ptr", ptr, (long long)n + 1); put_i1(f, "vsrc/pwl_rep", rep, n);
                    put_d1(f, "vsrc/pwl_rdelay", rd, n); put_d1(f, "vsrc/pwl", co, tot);
                    free(co);
                }
                free(ptr); free(rep); free(rd);
            }
            free(nodes); free(slots); free(fn); free(par);
        }
    }
    put_is(f, "vsrc/n", n);
    if (n) put_names(f, "vsrc/names_bytes", ckt->CKThead[vsrc_type]);
    /* current sources */
    n = 0;
    if (isrc_type >= 0) {
        ISRCmodel *m; ISRCinstance *h;
        for (m = (ISRCmodel *)ckt->CKThead[isrc_type]; m; m = ISRCnextModel(m))
            for (h = ISRCinstances(m); h; h = ISRCnextInstance(h)) n++;
        if (n) {
            int *nodes = (int *)calloc((size_t)n * 2, sizeof(int));
            int *fn = (int *)calloc((size_t)n * 3, sizeof(int));
            double *par = (double *)calloc((size_t)n * 10, sizeof(double));
            int k;
            i = 0;
            for (m = (ISRCmodel *)ckt->CKThead[isrc_type]; m; m = ISRCnextModel(m))
                for (h = ISRCinstances(m); h; h = ISRCnextInstance(h), i++) {
                    nodes[i] = h->ISRCposNode; nodes[n + i] = h->ISRCnegNode;
                    fn[i] = h->ISRCfunctionType; fn[n + i] = h->ISRCfunctionOrder; fn[2 * n + i] = h->ISRCdcGiven;
                    par[i] = h->ISRCdcValue; par[n + i] = h->ISRCmValue;
                    for (k = 0; k < 8; k++)
                        par[(size_t)(2 + k) * n + i] = (h->ISRCcoeffs && k < h->ISRCfunctionOrder) ? h->ISRCcoeffs[k] : 0.0;
                }
            put_i2(f, "isrc/nodes", nodes, 2, n); put_i2(f, "isrc/fn", fn, 3, n); put_d2(f, "isrc/par", par, 10, n);
            {   /* PWL corner lists */
                int *ptr = (int *)calloc((size_t)n + 1, sizeof(int)), tot = 0;
                double *co;
                i = 0;
                for (m = (ISRCmodel *)ckt->CKThead[isrc_type]; m; m = ISRCnextModel(m))
                    for (h = ISRCinstances(m); h; h = ISRCnextInstance(h), i++) {
                        ptr[i] = tot;
                        if (h->ISRCfunctionType == PWL) tot += h->ISRCfunctionOrder;
                    }
                ptr[n] = tot;
                if (tot) {
                    co = (double *)calloc((size_t)tot + 1, sizeof(double)); i = 0;
                    for (m = (ISRCmodel *)ckt->CKThead[isrc_type]; m; m = ISRCnextModel(m))
                        for (h = ISRCinstances(m); h; h = ISRCnextInstance(h), i++)
                            if (h->ISRCfunctionType == PWL) memcpy(co + ptr[i], h->ISRCcoeffs, sizeof(double) * (size_t)h->ISRCfunctionOrder);
                    put_i1(f, "isrc/pwl_ptr", ptr, (long long)n + 1); put_d1(f, "isrc/pwl", co, tot);
                    free(co);
                }
                free(ptr);
            }
            free(nodes); free(fn); free(par);
        }
    }
    put_is(f, "isrc/n", n);
}

/* BSIM3v3.3.0 instances: the same three-table form as BSIM4 (bsim3_fields.h) */
typedef struct { BSIM3model *m; struct bsim3SizeDependParam *p; } b3row_t;
static void dump_bsim3(FILE *f, CKTcircuit *ckt, KLUmatrix *K)
{
    BSIM3model *model; BSIM3instance *here;
    int n = 0, nrows = 0, i, r;
    b3row_t *rows; int *nodes, *slots, *flags, *sbase, *prow; double *inst, *mtab, *ptab;
    char *names; size_t nb = 0;
    if (b3_type < 0 || !ckt->CKThead[b3_type]) { put_is(f, "b3/ninst", 0); return; }
    for (model = (BSIM3model *)ckt->CKThead[b3_type]; model; model = BSIM3nextModel(model))
        for (here = BSIM3instances(model); here; here = BSIM3nextInstance(here)) n++;
    rows = (b3row_t *)calloc((size_t)n + 1, sizeof(b3row_t));
    nodes = (int *)calloc((size_t)n * B3N_COUNT + 1, sizeof(int));
    slots = (int *)calloc((size_t)n * (B3S_COUNT - B3S_RHS_COUNT) + 1, sizeof(int));
    flags = (int *)calloc((size_t)n + 1, sizeof(int)); sbase = (int *)calloc((size_t)n + 1, sizeof(int));
    prow = (int *)calloc((size_t)n + 1, sizeof(int));
    inst = (double *)calloc((size_t)n * B3I_COUNT + 1, sizeof(double));
    names = (char *)calloc((size_t)n * 64 + 1, 1);
    i = 0;
    for (model = (BSIM3model *)ckt->CKThead[b3_type]; model; model = BSIM3nextModel(model))
        for (here = BSIM3instances(model); here; here = BSIM3nextInstance(here), i++) {
            int k = 0;
            nodes[0 * n + i] = here->BSIM3dNode; nodes[1 * n + i] = here->BSIM3gNode; nodes[2 * n + i] = here->BSIM3sNode;
            nodes[3 * n + i] = here->BSIM3bNode; nodes[4 * n + i] = here->BSIM3dNodePrime; nodes[5 * n + i] = here->BSIM3sNodePrime;
#define X(nm, rr, cc) if (k >= B3S_RHS_COUNT) slots[(k - B3S_RHS_COUNT) * n + i] = slot_of(K, here->BSIM3##nm##Ptr); k++;
#define BSIM3rGPtr BSIM3GgPtr
#define BSIM3rBPtr BSIM3GgPtr
#define BSIM3rDPPtr BSIM3GgPtr
#define BSIM3rSPPtr BSIM3GgPtr
            NGB_B3_STAMPS(X)
#undef X
            k = 0;
#define X(nm) inst[(size_t)(k++) * n + i] = (double)here->BSIM3##nm;
            NGB_B3_INST_FIELDS(X)
#undef X
            flags[i] = (here->BSIM3off ? B3F_OFF : 0) | ((here->BSIM3nqsMod || here->BSIM3acnqsMod) ? B3F_NQS : 0);
            sbase[i] = here->BSIM3states;
            for (r = 0; r < nrows; r++) if (rows[r].m == model && rows[r].p == here->pParam) break;
            if (r == nrows) { rows[nrows].m = model; rows[nrows].p = here->pParam; nrows++; }
            prow[i] = r;
            nb += (size_t)snprintf(names + nb, 64, "%s\n", here->BSIM3name);
        }
    mtab = (double *)calloc((size_t)nrows * B3M_COUNT + 1, sizeof(double));
    ptab = (double *)calloc((size_t)nrows * B3P_COUNT + 1, sizeof(double));
    for (r = 0; r < nrows; r++) {
        int k = 0; BSIM3model *m = rows[r].m; struct bsim3SizeDependParam *pParam = rows[r].p;
#define X(nm) mtab[(size_t)r * B3M_COUNT + (k++)] = (double)m->BSIM3##nm;
        NGB_B3_MODEL_FIELDS(X)
#undef X
        k = 0;
#define X(nm) ptab[(size_t)r * B3P_COUNT + (k++)] = (double)pParam->BSIM3##nm;
        NGB_B3_BIN_FIELDS(X)
#undef X
    }
    put_is(f, "b3/ninst", n);
    put_i2(f, "b3/nodes", nodes, B3N_COUNT, n); put_i2(f, "b3/slots", slots, B3S_COUNT - B3S_RHS_COUNT, n);
    put_d2(f, "b3/inst", inst, B3I_COUNT, n); put_i1(f, "b3/flags", flags, n); put_i1(f, "b3/state_base", sbase, n);
    put_i1(f, "b3/prow", prow, n); put_d2(f, "b3/mtab", mtab, nrows, B3M_COUNT); put_d2(f, "b3/ptab", ptab, nrows, B3P_COUNT);
    { int *nb_i = (int *)calloc(nb + 1, sizeof(int)); size_t q; for (q = 0; q < nb; q++) nb_i[q] = (unsigned char)names[q];
      put_i1(f, "b3/names_bytes", nb_i, (long long)nb); free(nb_i); }
    free(rows); free(nodes); free(slots); free(flags); free(sbase); free(prow); free(inst); free(mtab); free(ptab); free(names);
}

/* diodes: node numbers, flags and the DIOtemp results DIOload reads (dio_fields.h) */
static void dump_dio(FILE *f, CKTcircuit *ckt, KLUmatrix *K)
{
    int n = 0, i;
    if (dio_type >= 0) {
        DIOmodel *m; DIOinstance *h;
        for (m = (DIOmodel *)ckt->CKThead[dio_type]; m; m = DIOnextModel(m))
            for (h = DIOinstances(m); h; h = DIOnextInstance(h)) n++;
        if (n) {
            int *nodes = (int *)calloc((size_t)n * 6, sizeof(int)), *slots = (int *)calloc((size_t)n * 7, sizeof(int));
            int *flags = (int *)calloc((size_t)n, sizeof(int)), *sb = (int *)calloc((size_t)n, sizeof(int));
            double *par = (double *)calloc((size_t)n * DIOP_COUNT, sizeof(double));
            size_t nb = 0; char *names = (char *)calloc((size_t)n * 64 + 1, 1);
            i = 0;
            for (m = (DIOmodel *)ckt->CKThead[dio_type]; m; m = DIOnextModel(m))
                for (h = DIOinstances(m); h; h = DIOnextInstance(h), i++) {
                    int fl = 0, k = 0;
                    nodes[i] = h->DIOposNode; nodes[n + i] = h->DIOnegNode; nodes[2 * n + i] = h->DIOtempNode;
                    nodes[3 * n + i] = h->DIOposPrimeNode; nodes[4 * n + i] = h->DIOposSwPrimeNode; nodes[5 * n + i] = h->DIOqpNode;
                    slots[i] = slot_of(K, h->DIOposPosPtr); slots[n + i] = slot_of(K, h->DIOnegNegPtr);
                    slots[2 * n + i] = slot_of(K, h->DIOposPrimePosPrimePtr); slots[3 * n + i] = slot_of(K, h->DIOposPosPrimePtr);
                    slots[4 * n + i] = slot_of(K, h->DIOnegPosPrimePtr); slots[5 * n + i] = slot_of(K, h->DIOposPrimePosPtr);
                    slots[6 * n + i] = slot_of(K, h->DIOposPrimeNegPtr);
                    if (h->DIOoff) fl |= DIOF_OFF;
                    if (m->DIObreakdownVoltageGiven) fl |= DIOF_BV;
                    if (m->DIOsatSWCurGiven) fl |= DIOF_SATSW;
                    if (m->DIOswEmissionCoeffGiven) fl |= DIOF_NSW;
                    if (m->DIOtunSatSWCurGiven) fl |= DIOF_TUNSW;
                    if (m->DIOtunSatCurGiven) fl |= DIOF_TUN;
                    if (m->DIOforwardKneeCurrentGiven) fl |= DIOF_IKF;
                    if (m->DIOreverseKneeCurrentGiven) fl |= DIOF_IKR;
                    if (m->DIOforwardSWKneeCurrentGiven) fl |= DIOF_IKP;
                    if (m->DIOrecSatCurGiven) fl |= DIOF_RECSAT;
                    if (m->DIOresistSWGiven) fl |= DIOF_RESISTSW;
                    if (m->DIOresistGiven) fl |= DIOF_RESIST;
                    if ((h->DIOtempNode > 0) && h->DIOthermal && m->DIOrth0Given) fl |= DIOF_SELFHEAT;
                    if ((h->DIOqpNode > 0) && (m->DIOsoftRevRecParam != 0) && (h->DIOtTransitTime != 0)) fl |= DIOF_REVREC;
                    flags[i] = fl; sb[i] = h->DIOstate;
#define X(nm) par[(size_t)(k++) * n + i] = h->DIO##nm;
                    NGB_DIO_INST_FIELDS(X)
#undef X
#define X(nm) par[(size_t)(k++) * n + i] = m->DIO##nm;
                    NGB_DIO_MODEL_FIELDS(X)
#undef X
#define X(nm) par[(size_t)(k++) * n + i] = h->DIO##nm;
                    NGB_DIO_RAW_INST_FIELDS(X)
#undef X
#define X(nm) par[(size_t)(k++) * n + i] = m->DIO##nm;
                    NGB_DIO_RAW_MODEL_FIELDS(X)
#undef X
                    nb += (size_t)snprintf(names + nb, 64, "%s\n", h->DIOname);
                }
            put_i2(f, "dio/nodes", nodes, 6, n); put_i2(f, "dio/slots", slots, 7, n);
            put_i1(f, "dio/flags", flags, n); put_i1(f, "dio/state_base", sb, n);
            put_d2(f, "dio/par", par, DIOP_COUNT, n);
            { int *nb_i = (int *)calloc(nb + 1, sizeof(int)); size_t q; for (q = 0; q < nb; q++) nb_i[q] = (unsigned char)names[q];
              put_i1(f, "dio/names_bytes", nb_i, (long long)nb); free(nb_i); }
            free(nodes); free(slots); free(flags); free(sb); free(par); free(names);
        }
    }
    put_is(f, "dio/n", n);
}

static void put_names(FILE *f, const char *key, GENmodel *head)
{
    GENmodel *m; GENinstance *h; size_t nb = 0, cap = 64; int *out = (int *)calloc(cap, sizeof(int));
    for (m = head; m; m = m->GENnextModel)
        for (h = m->GENinstances; h; h = h->GENnextInstance) {
            const char *s = h->GENname; size_t L = strlen(s), q;
            if (nb + L + 2 > cap) { cap = 2 * (nb + L + 2); out = (int *)realloc(out, cap * sizeof(int)); }
            for (q = 0; q < L; q++) out[nb++] = (unsigned char)s[q];
            out[nb++] = '\n';
        }
    put_i1(f, key, out, (long long)nb);
    free(out);
}

/* VBIC: the parameter vector as VBICload assembles it (vbicload.c:127-166), node numbers, flags */
static void dump_vbic(FILE *f, CKTcircuit *ckt)
{
    int n = 0, i, k;
    if (vbic_type >= 0) {
        VBICmodel *m; VBICinstance *h;
        for (m = (VBICmodel *)ckt->CKThead[vbic_type]; m; m = VBICnextModel(m))
            for (h = VBICinstances(m); h; h = VBICnextInstance(h)) n++;
        if (n) {
            int *nodes = (int *)calloc((size_t)n * VBN_COUNT, sizeof(int));
            int *flags = (int *)calloc((size_t)n, sizeof(int)), *sb = (int *)calloc((size_t)n, sizeof(int));
            double *par = (double *)calloc((size_t)n * VBIC_NP, sizeof(double)), *aux = (double *)calloc((size_t)n * VBA_COUNT, sizeof(double));
            size_t nb = 0; char *names = (char *)calloc((size_t)n * 64 + 1, 1);
            i = 0;
            for (m = (VBICmodel *)ckt->CKThead[vbic_type]; m; m = VBICnextModel(m))
                for (h = VBICinstances(m); h; h = VBICnextInstance(h), i++) {
                    double p[VBIC_NP];
                    const int nd[VBN_COUNT] = { h->VBICcollNode, h->VBICbaseNode, h->VBICemitNode, h->VBICsubsNode, h->VBICcollCXNode,
                        h->VBICcollCINode, h->VBICbaseBXNode, h->VBICbaseBINode, h->VBICemitEINode, h->VBICbaseBPNode, h->VBICsubsSINode,
                        (h->VBIC_selfheat && h->VBICtempNode > 0) ? h->VBICtempNode : 0, h->VBIC_excessPhase ? h->VBICxf1Node : 0, h->VBIC_excessPhase ? h->VBICxf2Node : 0 };
                    for (k = 0; k < VBN_COUNT; k++) nodes[(size_t)k * n + i] = nd[k];
                    memcpy(p, &m->VBICtnom, sizeof p);
                    p[0] = h->VBICtemp - CONSTCtoK + p[105];
                    p[1] = h->VBICtextCollResist; p[2] = h->VBICtintCollResist; p[3] = h->VBICtepiSatVoltage; p[4] = h->VBICtepiDoping;
                    p[6] = h->VBICtextBaseResist; p[7] = h->VBICtintBaseResist; p[8] = h->VBICtemitterResist; p[9] = h->VBICtsubstrateResist;
                    p[10] = h->VBICtparBaseResist; p[11] = h->VBICtsatCur; p[12] = h->VBICtemissionCoeffF; p[13] = h->VBICtemissionCoeffR;
                    p[16] = h->VBICtdepletionCapBE; p[17] = h->VBICtpotentialBE; p[21] = h->VBICtdepletionCapBC; p[23] = h->VBICtextCapBC;
                    p[24] = h->VBICtpotentialBC; p[27] = h->VBICtextCapSC; p[28] = h->VBICtpotentialSC; p[31] = h->VBICtidealSatCurBE;
                    p[34] = h->VBICtnidealSatCurBE; p[36] = h->VBICtidealSatCurBC; p[38] = h->VBICtnidealSatCurBC; p[41] = h->VBICtavalanchePar2BC;
                    p[42] = h->VBICtparasitSatCur; p[45] = h->VBICtidealParasitSatCurBE; p[46] = h->VBICtnidealParasitSatCurBE;
                    p[47] = h->VBICtidealParasitSatCurBC; p[49] = h->VBICtnidealParasitSatCurBC; p[53] = h->VBICtrollOffF;
                    p[94] = h->VBICtsepISRR; p[98] = h->VBICtvbbe; p[99] = h->VBICtnbbe;
                    for (k = 0; k < VBIC_NP; k++) par[(size_t)k * n + i] = p[k];
                    aux[(size_t)VBA_type * n + i] = m->VBICtype; aux[(size_t)VBA_tVcrit * n + i] = h->VBICtVcrit;
                    aux[(size_t)VBA_icVBE * n + i] = h->VBICicVBE; aux[(size_t)VBA_icVCE * n + i] = h->VBICicVCE;
                    aux[(size_t)VBA_scale * n + i] = h->VBICarea * h->VBICm; aux[(size_t)VBA_temp * n + i] = h->VBICtemp;
                    flags[i] = (h->VBICoff ? VBF_OFF : 0) | ((h->VBIC_selfheat && h->VBICtempNode > 0) ? VBF_SELFHEAT : 0) | (h->VBIC_excessPhase ? VBF_EXCESS : 0);
                    sb[i] = h->VBICstate;
                    nb += (size_t)snprintf(names + nb, 64, "%s\n", h->VBICname);
                }
            put_i2(f, "vbic/nodes", nodes, VBN_COUNT, n); put_i1(f, "vbic/flags", flags, n); put_i1(f, "vbic/state_base", sb, n);
            put_d2(f, "vbic/par", par, VBIC_NP, n); put_d2(f, "vbic/aux", aux, VBA_COUNT, n);
            { int *nb_i = (int *)calloc(nb + 1, sizeof(int)); size_t q; for (q = 0; q < nb; q++) nb_i[q] = (unsigned char)names[q];
              put_i1(f, "vbic/names_bytes", nb_i, (long long)nb); free(nb_i); }
            free(nodes); free(flags); free(sb); free(par); free(aux); free(names);
        }
    }
    put_is(f, "vbic/n", n);
}

static void dump_flat(CKTcircuit *ckt, const char *path)
{
    FILE *f = ngt_open(path);
    KLUmatrix *K = klu_of(ckt);
    int neq = ckt->CKTmaxEqNum, i;
    int *ntype = (int *)calloc((size_t)neq + 1, sizeof(int));
    double *nic = (double *)calloc((size_t)neq + 1, sizeof(double));
    int *nicg = (int *)calloc((size_t)neq + 1, sizeof(int));
    CKTnode *node;
    size_t nl = 0, ncap = 1 << 16; char *names = (char *)malloc(ncap);
    TRANan *job = (TRANan *)ckt->CKTcurJob;

    lookup_types();
    for (node = ckt->CKTnodes; node; node = node->next) {
        if (node->number >= 0 && node->number <= neq) {
            size_t l = strlen(node->name);
            ntype[node->number] = node->type;
            nic[node->number] = node->ic; nicg[node->number] = node->icGiven;
            if (nl + l + 24 > ncap) { ncap *= 2; names = (char *)realloc(names, ncap); }
            nl += (size_t)sprintf(names + nl, "%d %s\n", node->number, node->name);
        }
    }
    put_is(f, "meta/neq", neq);
    put_is(f, "meta/nstates", ckt->CKTnumStates);
    put_i1(f, "node/type", ntype, neq + 1);
    put_d1(f, "node/ic", nic, neq + 1);
    put_i1(f, "node/ic_given", nicg, neq + 1);
    { int *nb = (int *)malloc(sizeof(int) * (nl ? nl : 1)); size_t q;
      for (q = 0; q < nl; q++) nb[q] = (unsigned char)names[q];
      put_i1(f, "node/names_bytes", nb, (long long)nl); free(nb); }
    put_ds(f, "opt/reltol", ckt->CKTreltol); put_ds(f, "opt/abstol", ckt->CKTabstol);
    put_ds(f, "opt/vntol", ckt->CKTvoltTol); put_ds(f, "opt/chgtol", ckt->CKTchgtol);
    put_ds(f, "opt/trtol", ckt->CKTtrtol); put_ds(f, "opt/gmin", ckt->CKTgmin);
    put_ds(f, "opt/diag_gmin", ckt->CKTdiagGmin);
    put_ds(f, "opt/temp", ckt->CKTtemp); put_ds(f, "opt/nomtemp", ckt->CKTnomTemp);
    put_ds(f, "opt/xmu", ckt->CKTxmu); put_is(f, "opt/maxorder", ckt->CKTmaxOrder);
    put_is(f, "opt/method", ckt->CKTintegrateMethod);
    put_is(f, "opt/itl4", ckt->CKTtranMaxIter); put_is(f, "opt/itl1", ckt->CKTdcMaxIter);
    put_ds(f, "opt/pivabstol", ckt->CKTpivotAbsTol); put_ds(f, "opt/pivreltol", ckt->CKTpivotRelTol);
    put_ds(f, "opt/delmin", ckt->CKTdelmin); put_ds(f, "opt/minbreak", ckt->CKTminBreak);
    put_is(f, "opt/bypass", ckt->CKTbypass); put_is(f, "opt/newtrunc", ckt->CKTnewtrunc);
    put_ds(f, "opt/vt0", CONSTvt0);
    put_is(f, "opt/gminsteps", ckt->CKTnumGminSteps); put_is(f, "opt/srcsteps", ckt->CKTnumSrcSteps);
    put_is(f, "opt/itl2", ckt->CKTdcTrcvMaxIter); put_ds(f, "opt/gminfactor", ckt->CKTgminFactor);
    put_is(f, "opt/noopiter", ckt->CKTnoOpIter); put_ds(f, "opt/gshunt", ckt->CKTgshunt);
    put_ds(f, "tran/tstep", ckt->CKTstep); put_ds(f, "tran/tstop", ckt->CKTfinalTime);
    put_ds(f, "tran/tmax", ckt->CKTmaxStep); put_ds(f, "tran/tstart", ckt->CKTinitTime);
    put_is(f, "tran/uic", (ckt->CKTmode & MODEUIC) ? 1 : 0);
    (void)job;
    if (K) {
        int n = (int)K->KLUmatrixN, nz = (int)K->KLUmatrixNZ;
        int *diag = (int *)calloc((size_t)n, sizeof(int));
        int *n2o = (int *)calloc((size_t)n + 2, sizeof(int));
        for (i = 0; i < n; i++) diag[i] = slot_of(K, K->KLUmatrixDiag[i]);
        for (i = 0; i <= n; i++) n2o[i] = (int)K->KLUmatrixNodeCollapsingNewToOld[i];
        put_is(f, "klu/n", n); put_is(f, "klu/nz", nz); put_is(f, "klu/nrhs", (int)K->KLUmatrixNrhs);
        put_i1(f, "klu/Ap", K->KLUmatrixAp, n + 1); put_i1(f, "klu/Ai", K->KLUmatrixAi, nz);
        put_i1(f, "klu/diag", diag, n); put_i1(f, "klu/new2old", n2o, n + 1);
        free(diag); free(n2o);
    } else {
        put_is(f, "klu/n", 0);
    }
    dump_bsim4(f, ckt, K);
    dump_linear(f, ckt, K);
    dump_dio(f, ckt, K);
    dump_vbic(f, ckt);
    {   /* .nodeset / .ic nodes as CKTic left them: nodesets first, then initial conditions, each in node order */
        CKTnode *nd; int n = 0, pass, i = 0; int *eq, *kind; double *val;
        for (nd = ckt->CKTnodes; nd; nd = nd->next) n += (nd->nsGiven ? 1 : 0) + (nd->icGiven ? 1 : 0);
        eq = (int *)calloc((size_t)n + 1, sizeof(int)); kind = (int *)calloc((size_t)n + 1, sizeof(int)); val = (double *)calloc((size_t)n + 1, sizeof(double));
        for (pass = 0; pass < 2; pass++)
            for (nd = ckt->CKTnodes; nd; nd = nd->next)
                if (pass ? nd->icGiven : nd->nsGiven) { eq[i] = nd->number; kind[i] = pass; val[i] = pass ? nd->ic : nd->nodeset; i++; }
        put_is(f, "node/nov", n);
        if (n) { put_i1(f, "node/ov_eq", eq, n); put_i1(f, "node/ov_kind", kind, n); put_d1(f, "node/ov_val", val, n); }
        free(eq); free(kind); free(val);
    }
    dump_bsim3(f, ckt, K);
    fclose(f);
    free(ntype); free(nic); free(nicg); free(names);
}

/* ---------------------------------------------------------------- per-call trace */
static FILE *trace_f = NULL;
static int call_no = -1;          /* index of the current CKTload call */
static int call_selected = 0;
static char sel_spec[4096];
static CKTcircuit *cur_ckt = NULL;
static int (*orig_b4_load)(GENmodel *, CKTcircuit *) = NULL;

static int is_selected(int k)
{
    const char *p = sel_spec;
    while (*p) {
        int a, b, n = 0;
        if (sscanf(p, "%d-%d%n", &a, &b, &n) == 2 && n > 0) { if (k >= a && k <= b) return 1; }
        else if (sscanf(p, "%d%n", &a, &n) == 1 && n > 0) { if (k == a) return 1; }
        else break;
        p += n;
        if (*p == ',') p++;
    }
    return 0;
}
static void tname(char *buf, const char *what) { sprintf(buf, "c%d/%s", call_no, what); }

static void trace_b4_op(CKTcircuit *ckt, const char *what)
{
    BSIM4model *model; BSIM4instance *here; int n = 0, i = 0; double *op; char nm[64];
    if (b4_type < 0 || !ckt->CKThead[b4_type]) return;
    for (model = (BSIM4model *)ckt->CKThead[b4_type]; model; model = BSIM4nextModel(model))
        for (here = BSIM4instances(model); here; here = BSIM4nextInstance(here)) n++;
    op = (double *)calloc((size_t)n * B4O_COUNT, sizeof(double));
    for (model = (BSIM4model *)ckt->CKThead[b4_type]; model; model = BSIM4nextModel(model))
        for (here = BSIM4instances(model); here; here = BSIM4nextInstance(here), i++) {
            int k = 0;
#define X(f_) op[(size_t)(k++) * n + i] = (double)here->BSIM4##f_;
            NGB_B4_OP_FIELDS(X)
#undef X
        }
    tname(nm, what);
    put_d2(trace_f, nm, op, B4O_COUNT, n);
    free(op);
}

static int b4_load_wrapper(GENmodel *head, CKTcircuit *ckt)
{
    int err, before = ckt->CKTnoncon;
    char nm[64];
    if (call_selected && trace_f) trace_b4_op(ckt, "b4_op_in");
    err = orig_b4_load(head, ckt);
    if (call_selected && trace_f) {
        KLUmatrix *K = klu_of(ckt);
        if (K) { tname(nm, "b4_Ax"); put_d1(trace_f, nm, K->KLUmatrixAx, K->KLUmatrixNZ); }
        tname(nm, "b4_rhs"); put_d1(trace_f, nm, ckt->CKTrhs, ckt->CKTmaxEqNum + 1);
        tname(nm, "b4_state0"); put_d1(trace_f, nm, ckt->CKTstate0, ckt->CKTnumStates);
        tname(nm, "b4_state1"); put_d1(trace_f, nm, ckt->CKTstate1, ckt->CKTnumStates);
        tname(nm, "b4_noncon"); put_is(trace_f, nm, ckt->CKTnoncon - before);
        trace_b4_op(ckt, "b4_op_out");
    }
    return err;
}

static void dump_klu_pattern(KLUmatrix *K, const char *prefix)
{
    klu_symbolic *Sy = K->KLUmatrixSymbolic; klu_numeric *Nu = K->KLUmatrixNumeric;
    int n, nb, b, k, lnz = 0, unz = 0; int *Lp, *Up, *Li, *Ui; char nm[96];
    if (!Sy || !Nu) return;
    n = Sy->n; nb = Sy->nblocks;
    Lp = (int *)calloc((size_t)n + 1, sizeof(int)); Up = (int *)calloc((size_t)n + 1, sizeof(int));
    for (b = 0; b < nb; b++) {
        int k1 = Sy->R[b], k2 = Sy->R[b + 1];
        for (k = k1; k < k2; k++) {
            Lp[k] = lnz; Up[k] = unz;
            if (k2 - k1 > 1) { lnz += Nu->Llen[k]; unz += Nu->Ulen[k]; }   /* singletons: lengths unset */
        }
    }
    Lp[n] = lnz; Up[n] = unz;
    Li = (int *)calloc((size_t)lnz + 1, sizeof(int)); Ui = (int *)calloc((size_t)unz + 1, sizeof(int));
    for (b = 0; b < nb; b++) {
        int k1 = Sy->R[b], k2 = Sy->R[b + 1];
        if (k2 - k1 == 1) continue;
        for (k = k1; k < k2; k++) {
            Unit *LU = (Unit *)Nu->LUbx[b];
            int *li = (int *)(LU + Nu->Lip[k]), *ui = (int *)(LU + Nu->Uip[k]), p;
            for (p = 0; p < Nu->Llen[k]; p++) Li[Lp[k] + p] = li[p] + k1;
            for (p = 0; p < Nu->Ulen[k]; p++) Ui[Up[k] + p] = ui[p] + k1;
        }
    }
    sprintf(nm, "%s/n", prefix); put_is(trace_f, nm, n);
    sprintf(nm, "%s/nblocks", prefix); put_is(trace_f, nm, nb);
    sprintf(nm, "%s/nzoff", prefix); put_is(trace_f, nm, Sy->nzoff);
    sprintf(nm, "%s/P", prefix); put_i1(trace_f, nm, Sy->P, n);
    sprintf(nm, "%s/Q", prefix); put_i1(trace_f, nm, Sy->Q, n);
    sprintf(nm, "%s/R", prefix); put_i1(trace_f, nm, Sy->R, nb + 1);
    sprintf(nm, "%s/Pnum", prefix); put_i1(trace_f, nm, Nu->Pnum, n);
    sprintf(nm, "%s/Pinv", prefix); put_i1(trace_f, nm, Nu->Pinv, n);
    sprintf(nm, "%s/Lp", prefix); put_i1(trace_f, nm, Lp, n + 1);
    sprintf(nm, "%s/Li", prefix); put_i1(trace_f, nm, Li, lnz);
    sprintf(nm, "%s/Up", prefix); put_i1(trace_f, nm, Up, n + 1);
    sprintf(nm, "%s/Ui", prefix); put_i1(trace_f, nm, Ui, unz);
    sprintf(nm, "%s/Offp", prefix); put_i1(trace_f, nm, Nu->Offp, n + 1);
    sprintf(nm, "%s/Offi", prefix); put_i1(trace_f, nm, Nu->Offi, Sy->nzoff);
    free(Lp); free(Up); free(Li); free(Ui);
}

static void dump_klu_values(KLUmatrix *K, const char *prefix)
{
    klu_symbolic *Sy = K->KLUmatrixSymbolic; klu_numeric *Nu = K->KLUmatrixNumeric;
    int n, nb, b, k, lnz = 0, unz = 0, pl = 0, pu = 0; double *Lx, *Ux; char nm[96];
    if (!Sy || !Nu) return;
    n = Sy->n; nb = Sy->nblocks;
    for (b = 0; b < nb; b++) {
        int k1 = Sy->R[b], k2 = Sy->R[b + 1];
        if (k2 - k1 > 1) for (k = k1; k < k2; k++) { lnz += Nu->Llen[k]; unz += Nu->Ulen[k]; }
    }
    Lx = (double *)calloc((size_t)lnz + 1, sizeof(double)); Ux = (double *)calloc((size_t)unz + 1, sizeof(double));
    for (b = 0; b < nb; b++) {
        int k1 = Sy->R[b], k2 = Sy->R[b + 1];
        for (k = k1; k < k2; k++) {
            if (k2 - k1 > 1) {
                Unit *LU = (Unit *)Nu->LUbx[b];
                double *lx = (double *)(LU + Nu->Lip[k] + UNITS(Int, Nu->Llen[k]));
                double *ux = (double *)(LU + Nu->Uip[k] + UNITS(Int, Nu->Ulen[k]));
                int p;
                for (p = 0; p < Nu->Llen[k]; p++) Lx[pl + p] = lx[p];
                for (p = 0; p < Nu->Ulen[k]; p++) Ux[pu + p] = ux[p];
                pl += Nu->Llen[k]; pu += Nu->Ulen[k];
            }
        }
    }
    sprintf(nm, "%s/Lx", prefix); put_d1(trace_f, nm, Lx, lnz);
    sprintf(nm, "%s/Ux", prefix); put_d1(trace_f, nm, Ux, unz);
    sprintf(nm, "%s/Udiag", prefix); put_d1(trace_f, nm, (double *)Nu->Udiag, n);
    if (Nu->Rs) { sprintf(nm, "%s/Rs", prefix); put_d1(trace_f, nm, Nu->Rs, n); }
    sprintf(nm, "%s/Offx", prefix); put_d1(trace_f, nm, (double *)Nu->Offx, Sy->nzoff);
    free(Lx); free(Ux);
}

static int op_loads = 0;
int __wrap_CKTload(CKTcircuit *ckt)
{
    static int flat_done = 0;
    const char *flat = getenv("NGB_DUMP_FLAT"), *trace = getenv("NGB_DUMP_TRACE");
    int err;
    char nm[64];
    cur_ckt = ckt;
    if (ckt->CKTmode & MODETRANOP) op_loads++;          /* loads of CKTop: plain NIiter + its fallbacks (OPtran loads under MODETRAN) */
    lookup_types();
    if (flat && !flat_done) { flat_done = 1; dump_flat(ckt, flat); }
    if (trace && !trace_f) {
        const char *sel = getenv("NGB_DUMP_CALLS");
        trace_f = ngt_open(trace);
        strncpy(sel_spec, sel ? sel : "0-3", sizeof(sel_spec) - 1);
    }
    if (trace_f && b4_type >= 0 && DEVices[b4_type] && DEVices[b4_type]->DEVload != b4_load_wrapper) {
        orig_b4_load = DEVices[b4_type]->DEVload;
        DEVices[b4_type]->DEVload = b4_load_wrapper;
    }
    call_no++;
    call_selected = trace_f ? is_selected(call_no) : 0;
    if (call_selected) {
        tname(nm, "mode"); put_is(trace_f, nm, (int)ckt->CKTmode);
        tname(nm, "time"); put_ds(trace_f, nm, ckt->CKTtime);
        tname(nm, "delta"); put_ds(trace_f, nm, ckt->CKTdelta);
        tname(nm, "deltaOld"); put_d1(trace_f, nm, ckt->CKTdeltaOld, 7);
        tname(nm, "ag"); put_d1(trace_f, nm, ckt->CKTag, 7);
        tname(nm, "order"); put_is(trace_f, nm, ckt->CKTorder);
        tname(nm, "gmin"); put_ds(trace_f, nm, ckt->CKTgmin);
        tname(nm, "diag_gmin"); put_ds(trace_f, nm, ckt->CKTdiagGmin);
        tname(nm, "srcfact"); put_ds(trace_f, nm, ckt->CKTsrcFact);
        tname(nm, "rhsOld"); put_d1(trace_f, nm, ckt->CKTrhsOld, ckt->CKTmaxEqNum + 1);
        tname(nm, "state0_in"); put_d1(trace_f, nm, ckt->CKTstate0, ckt->CKTnumStates);
        tname(nm, "state1_in"); put_d1(trace_f, nm, ckt->CKTstate1, ckt->CKTnumStates);
        if (ckt->CKTstate2) { tname(nm, "state2_in"); put_d1(trace_f, nm, ckt->CKTstate2, ckt->CKTnumStates); }
        if (b3_type >= 0 && ckt->CKThead[b3_type]) {       /* previous-iterate von: input of DEVfetlim */
            BSIM3model *m; BSIM3instance *hh; int n3 = 0, q = 0; double *v;
            for (m = (BSIM3model *)ckt->CKThead[b3_type]; m; m = BSIM3nextModel(m))
                for (hh = BSIM3instances(m); hh; hh = BSIM3nextInstance(hh)) n3++;
            v = (double *)calloc((size_t)n3 + 1, sizeof(double));
            for (m = (BSIM3model *)ckt->CKThead[b3_type]; m; m = BSIM3nextModel(m))
                for (hh = BSIM3instances(m); hh; hh = BSIM3nextInstance(hh)) v[q++] = hh->BSIM3von;
            tname(nm, "b3_von_in"); put_d1(trace_f, nm, v, n3); free(v);
        }
    }
    err = __real_CKTload(ckt);
    if (call_selected) {
        KLUmatrix *K = klu_of(ckt);
        if (K) { tname(nm, "Ax"); put_d1(trace_f, nm, K->KLUmatrixAx, K->KLUmatrixNZ); }
        tname(nm, "rhs"); put_d1(trace_f, nm, ckt->CKTrhs, ckt->CKTmaxEqNum + 1);
        tname(nm, "state0_out"); put_d1(trace_f, nm, ckt->CKTstate0, ckt->CKTnumStates);
        tname(nm, "state1_out"); put_d1(trace_f, nm, ckt->CKTstate1, ckt->CKTnumStates);
        tname(nm, "noncon"); put_is(trace_f, nm, ckt->CKTnoncon);
        fflush(trace_f);
    }
    return err;
}

int __wrap_SMPreorder(SMPmatrix *M, double a, double b, double g)
{
    int r = __real_SMPreorder(M, a, b, g);
    if (trace_f && M->CKTkluMODE && r == 0) {
        char nm[64];
        /* the pivot order can change at every pivoting factor: record it once per event for
         * selected calls, and always for the first one */
        static int first = 1;
        if (call_selected || first) {
            sprintf(nm, "c%d/pat", call_no); dump_klu_pattern(M->SMPkluMatrix, nm);
            first = 0;
        }
        if (call_selected) {
            sprintf(nm, "c%d/lu", call_no); dump_klu_values(M->SMPkluMatrix, nm);
            sprintf(nm, "c%d/factor_kind", call_no); put_is(trace_f, nm, 2);
            sprintf(nm, "c%d/Ax_fact", call_no);
            put_d1(trace_f, nm, M->SMPkluMatrix->KLUmatrixAx, M->SMPkluMatrix->KLUmatrixNZ);
        }
    }
    return r;
}

int __wrap_SMPluFac(SMPmatrix *M, double a, double g)
{
    int r = __real_SMPluFac(M, a, g);
    if (trace_f && M->CKTkluMODE && call_selected) {
        char nm[64];
        sprintf(nm, "c%d/lu", call_no); dump_klu_values(M->SMPkluMatrix, nm);
        sprintf(nm, "c%d/factor_kind", call_no); put_is(trace_f, nm, 1);
        sprintf(nm, "c%d/factor_ret", call_no); put_is(trace_f, nm, r);
        sprintf(nm, "c%d/Ax_fact", call_no);
        put_d1(trace_f, nm, M->SMPkluMatrix->KLUmatrixAx, M->SMPkluMatrix->KLUmatrixNZ);
    }
    return r;
}

void __wrap_SMPsolve(SMPmatrix *M, double RHS[], double Spare[])
{
    __real_SMPsolve(M, RHS, Spare);
    if (trace_f && call_selected && cur_ckt) {
        char nm[64];
        sprintf(nm, "c%d/sol", call_no);
        put_d1(trace_f, nm, RHS, cur_ckt->CKTmaxEqNum + 1);
        fflush(trace_f);
    }
}

int __wrap_DCtran(CKTcircuit *ckt, int restart)
{
    int r = __real_DCtran(ckt, restart);
    const char *st = getenv("NGB_DUMP_STATS");
    if (st) {
        FILE *f = fopen(st, "w");
        if (f) {
            STATistics *s = ckt->CKTstat;
            fprintf(f, "{\"ret\": %d, \"accepted\": %d, \"rejected\": %d, \"numiter\": %d, \"timepts\": %d, "
                       "\"load_calls\": %d, \"load_time\": %.9g, \"decomp_time\": %.9g, \"reorder_time\": %.9g, "
                       "\"solve_time\": %.9g, \"tran_time\": %.9g, \"op_loads\": %d}\n",
                    r, s->STATaccepted, s->STATrejected, s->STATnumIter, s->STATtimePts, call_no + 1,
                    s->STATloadTime, s->STATdecompTime, s->STATreorderTime, s->STATsolveTime, s->STATtranTime, op_loads);
            fclose(f);
        }
    }
    if (trace_f) fflush(trace_f);
    return r;
}
