/* ngb_shim.c -- reference-side binding of libngb200 (level 1).
 *
 * Linked into the UNMODIFIED ngspice objects with
 *     -Wl,--wrap=CKTload,--wrap=SMPreorder,--wrap=SMPluFac,--wrap=SMPsolve
 * NIiter, DCop, DCtran and the front end run unchanged; what changes is who executes the hot path:
 *
 *   CKTload     -> ngbLoad      device evaluation + stamping + assembly on the GPU.  The circuit is
 *                               flattened once from what CKTsetup/CKTtemp left in the instance and model
 *                               structures; per call the shim ships CKTmode / CKTag / CKTdelta ...,
 *                               CKTrhsOld and the state history down, and Ax, CKTrhs, CKTstate0/1 and
 *                               CKTnoncon back, so every other reference routine (DEVtrunc, DEVaccept,
 *                               CKTdump, the host KLU) keeps seeing the data it expects.
 *   SMPreorder  -> reference klu_factor on the host (pivoting stays on the CPU), then its pattern and
 *                  pivot order become the device schedule (ngbCircuitSetLuPattern) and the device
 *                  refactors on it.
 *   SMPluFac    -> ngbLuFac     (row scaling + refactor on the GPU; 102 == E_SINGULAR like KLU)
 *   SMPsolve    -> ngbSolve     (triangular solves on the GPU)
 *
 * TABLE MODE (NGB_SHIM_TABLE=1, and automatically whenever the circuit holds a device type the library does not
 * implement): the library is installed through the SPICEdev plugin table instead (devdefs.h:50-129, dev.c:142-209).
 * DEVices[t]->DEVload of every device type the library implements is replaced; the reference's own CKTload then runs
 * UNCHANGED -- it clears the matrix, walks DEVices[] (cktload.c:70-91), applies the .nodeset / .ic rows and counts
 * CKTnoncon.  The first replaced DEVload of a pass evaluates ALL the replaced types in one ngbLoad and adds their
 * assembled contributions into the KLU matrix and CKTrhs, the others return OK; device types without a replacement
 * (BJT, inductors, controlled sources, ...) keep their CPU DEVload and stamp the same matrix.  With such types present
 * the factorisation stays on the host KLU (their entries are not in the device pattern); without, SMPluFac / SMPsolve
 * run on the device on the matrix and right-hand side the host CKTload finished.
 *
 * The shim is inactive -- every call goes to the reference -- when the matrix is not in KLU mode, when a replaced
 * device type uses an option outside the GPU path, or with NGB_SHIM=0.
 * NGB_SHIM_LU=0 keeps the LU on the host (device load only).
 */
#include "ngspice/ngspice.h"
#include "ngspice/cktdefs.h"
#include "ngspice/devdefs.h"
#include "ngspice/smpdefs.h"
#include "ngspice/sperror.h"
#include "ngspice/const.h"
#include "ngspice/klu.h"
#include "bsim4/bsim4def.h"
#include "bsim3/bsim3def.h"
#include "dio/diodefs.h"
#include "vbic/vbicdefs.h"
#include "res/resdefs.h"
#include "cap/capdefs.h"
#include "vsrc/vsrcdefs.h"
#include "isrc/isrcdefs.h"
#include "klu_internal.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../include/ngb200.h"
#include "../ngspice-sf-mirror_b200/csrc/bsim4_fields.h"
#include "../ngspice-sf-mirror_b200/csrc/bsim3_fields.h"
#include "../ngspice-sf-mirror_b200/csrc/dio_fields.h"
#include "../ngspice-sf-mirror_b200/csrc/vbic_types.h"

extern SPICEdev **DEVices;
extern int DEVmaxnum;
int CKTtypelook(char *);
int __real_CKTload(CKTcircuit *ckt);
int __real_SMPreorder(SMPmatrix *, double, double, double);
int __real_SMPluFac(SMPmatrix *, double, double);
void __real_SMPsolve(SMPmatrix *, double[], double[]);

static struct {
    int tried, active, use_lu, have_lu, dev_factored;
    ngb_circuit *C; ngb_batch *B;
    CKTcircuit *ckt;
    int neq, nnz;
    int tB4, tB3, tDIO, tVBIC, tRES, tCAP, tVSRC, tISRC;
    int n4, n3, nd, nq, nc;
    int *sb4, *sb3, *sbd, *sbq, *sbc;          /* state bases per instance */
    double *buf; size_t buf_len;
    long loads, facs, solves;
    /* table mode */
    int table, mixed, leader;                  /* leader: the first replaced device type in DEVices order */
    int *slot_map;                             /* our CSC slot -> slot of the reference's KLU matrix (NULL: identical patterns) */
    int (*orig_load[8])(GENmodel *, CKTcircuit *); int orig_type[8], norig;
} G;

static void *xc(size_t n, size_t sz) { void *p = calloc(n ? n : 1, sz); if (!p) { fprintf(stderr, "ngb_shim: out of memory\n"); exit(1); } return p; }
static double *scratch(size_t n) { if (n > G.buf_len) { free(G.buf); G.buf = (double *)xc(n, sizeof(double)); G.buf_len = n; } return G.buf; }

#define COUNT(T, type, var) do { var = 0; if (type >= 0) { T##model *m_; T##instance *h_; \
    for (m_ = (T##model *)ckt->CKThead[type]; m_; m_ = T##nextModel(m_)) for (h_ = T##instances(m_); h_; h_ = T##nextInstance(h_)) var++; } } while (0)

static int shim_fail(const char *why) { fprintf(stderr, "ngb_shim: inactive (%s)\n", why); G.active = 0; return 0; }

/* ------------------------------------------------------------------ flatten the circuit (once) */
static int flatten_bsim4(CKTcircuit *ckt)
{
    typedef struct { BSIM4model *m; struct bsim4SizeDependParam *p; } row_t;
    BSIM4model *model; BSIM4instance *here;
    int n = G.n4, nrows = 0, i = 0, r, rc;
    row_t *rows; int *nodes, *flags, *prow; double *inst, *mtab, *ptab;
    if (!n) return 0;
    rows = (row_t *)xc((size_t)n, sizeof *rows); nodes = (int *)xc((size_t)n * B4N_COUNT, sizeof(int));
    flags = (int *)xc((size_t)n, sizeof(int)); prow = (int *)xc((size_t)n, sizeof(int)); G.sb4 = (int *)xc((size_t)n, sizeof(int));
    inst = (double *)xc((size_t)n * B4I_COUNT, sizeof(double));
    for (model = (BSIM4model *)ckt->CKThead[G.tB4]; model; model = BSIM4nextModel(model))
        for (here = BSIM4instances(model); here; here = BSIM4nextInstance(here), i++) {
            int k = 0;
#define X(nm) nodes[(k++) * n + i] = here->BSIM4##nm;
            NGB_B4_NODE_FIELDS(X)
#undef X
            k = 0;
#define X(nm) inst[(size_t)(k++) * n + i] = (double)here->BSIM4##nm;
            NGB_B4_INST_FIELDS(X)
#undef X
            flags[i] = (here->BSIM4off ? B4F_OFF : 0) | ((here->BSIM4rbodyMod & 3) << B4F_RBODY_SH) | ((here->BSIM4rgateMod & 3) << B4F_RGATE_SH)
                     | (here->BSIM4trnqsMod ? 0x100 : 0) | (here->BSIM4acnqsMod ? 0x200 : 0);
            G.sb4[i] = here->BSIM4states;
            for (r = 0; r < nrows; r++) if (rows[r].m == model && rows[r].p == here->pParam) break;
            if (r == nrows) { rows[nrows].m = model; rows[nrows].p = here->pParam; nrows++; }
            prow[i] = r;
        }
    mtab = (double *)xc((size_t)nrows * B4M_COUNT, sizeof(double)); ptab = (double *)xc((size_t)nrows * B4P_COUNT, sizeof(double));
    for (r = 0; r < nrows; r++) {
        int k = 0; BSIM4model *m = rows[r].m; struct bsim4SizeDependParam *pParam = rows[r].p;
#define X(nm) mtab[(size_t)r * B4M_COUNT + (k++)] = (double)m->BSIM4##nm;
        NGB_B4_MODEL_FIELDS(X)
#undef X
        k = 0;
#define X(nm) ptab[(size_t)r * B4P_COUNT + (k++)] = (double)pParam->BSIM4##nm;
        NGB_B4_BIN_FIELDS(X)
#undef X
    }
    rc = ngbCircuitAddBsim4(G.C, n, nodes, flags, prow, inst, nrows, mtab, ptab);
    free(rows); free(nodes); free(flags); free(prow); free(inst); free(mtab); free(ptab);
    return rc;
}

static int flatten_bsim3(CKTcircuit *ckt)
{
    typedef struct { BSIM3model *m; struct bsim3SizeDependParam *p; } row_t;
    BSIM3model *model; BSIM3instance *here;
    int n = G.n3, nrows = 0, i = 0, r, rc;
    row_t *rows; int *nodes, *flags, *prow; double *inst, *mtab, *ptab;
    if (!n) return 0;
    rows = (row_t *)xc((size_t)n, sizeof *rows); nodes = (int *)xc((size_t)n * B3N_COUNT, sizeof(int));
    flags = (int *)xc((size_t)n, sizeof(int)); prow = (int *)xc((size_t)n, sizeof(int)); G.sb3 = (int *)xc((size_t)n, sizeof(int));
    inst = (double *)xc((size_t)n * B3I_COUNT, sizeof(double));
    for (model = (BSIM3model *)ckt->CKThead[G.tB3]; model; model = BSIM3nextModel(model))
        for (here = BSIM3instances(model); here; here = BSIM3nextInstance(here), i++) {
            int k = 0;
            nodes[0 * n + i] = here->BSIM3dNode; nodes[1 * n + i] = here->BSIM3gNode; nodes[2 * n + i] = here->BSIM3sNode;
            nodes[3 * n + i] = here->BSIM3bNode; nodes[4 * n + i] = here->BSIM3dNodePrime; nodes[5 * n + i] = here->BSIM3sNodePrime;
#define X(nm) inst[(size_t)(k++) * n + i] = (double)here->BSIM3##nm;
            NGB_B3_INST_FIELDS(X)
#undef X
            flags[i] = (here->BSIM3off ? B3F_OFF : 0) | ((here->BSIM3nqsMod || here->BSIM3acnqsMod) ? B3F_NQS : 0);
            G.sb3[i] = here->BSIM3states;
            for (r = 0; r < nrows; r++) if (rows[r].m == model && rows[r].p == here->pParam) break;
            if (r == nrows) { rows[nrows].m = model; rows[nrows].p = here->pParam; nrows++; }
            prow[i] = r;
        }
    mtab = (double *)xc((size_t)nrows * B3M_COUNT, sizeof(double)); ptab = (double *)xc((size_t)nrows * B3P_COUNT, sizeof(double));
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
    rc = ngbCircuitAddBsim3(G.C, n, nodes, flags, prow, inst, nrows, mtab, ptab);
    free(rows); free(nodes); free(flags); free(prow); free(inst); free(mtab); free(ptab);
    return rc;
}

static int flatten_dio(CKTcircuit *ckt)
{
    DIOmodel *m; DIOinstance *h;
    int n = G.nd, i = 0, rc; int *nodes, *flags; double *par;
    if (!n) return 0;
    nodes = (int *)xc((size_t)n * DION_COUNT, sizeof(int)); flags = (int *)xc((size_t)n, sizeof(int)); G.sbd = (int *)xc((size_t)n, sizeof(int));
    par = (double *)xc((size_t)n * DIOP_COUNT, sizeof(double));
    for (m = (DIOmodel *)ckt->CKThead[G.tDIO]; m; m = DIOnextModel(m))
        for (h = DIOinstances(m); h; h = DIOnextInstance(h), i++) {
            int fl = 0, k = 0;
            nodes[i] = h->DIOposNode; nodes[n + i] = h->DIOnegNode; nodes[2 * n + i] = h->DIOposPrimeNode;
            nodes[3 * n + i] = h->DIOposSwPrimeNode;
            if (h->DIOoff) fl |= DIOF_OFF;
            if (m->DIOresistGiven) fl |= DIOF_RESIST;
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
            if ((h->DIOtempNode > 0) && h->DIOthermal && m->DIOrth0Given) fl |= DIOF_SELFHEAT;
            if ((h->DIOqpNode > 0) && (m->DIOsoftRevRecParam != 0) && (h->DIOtTransitTime != 0)) fl |= DIOF_REVREC;
            flags[i] = fl; G.sbd[i] = h->DIOstate;
            nodes[4 * n + i] = (fl & DIOF_SELFHEAT) ? h->DIOtempNode : 0;      /* unused thermal / qp nodes stay with the reference's own stamps (none) */
            nodes[5 * n + i] = (fl & DIOF_REVREC) ? h->DIOqpNode : 0;
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
        }
    rc = ngbCircuitAddDiodes(G.C, n, nodes, flags, par);
    free(nodes); free(flags); free(par);
    return rc;
}

/* VBIC: the parameter vector is put together the way VBICload does it per instance (vbicload.c:127-166) */
static int flatten_vbic(CKTcircuit *ckt)
{
    VBICmodel *m; VBICinstance *h;
    int n = G.nq, i = 0, k, rc; int *nodes, *flags; double *par, *aux;
    if (!n) return 0;
    nodes = (int *)xc((size_t)n * VBN_COUNT, sizeof(int)); flags = (int *)xc((size_t)n, sizeof(int)); G.sbq = (int *)xc((size_t)n, sizeof(int));
    par = (double *)xc((size_t)n * VBIC_NP, sizeof(double)); aux = (double *)xc((size_t)n * VBA_COUNT, sizeof(double));
    for (m = (VBICmodel *)ckt->CKThead[G.tVBIC]; m; m = VBICnextModel(m))
        for (h = VBICinstances(m); h; h = VBICnextInstance(h), i++) {
            double p[VBIC_NP];
            const int nd[VBN_COUNT] = { h->VBICcollNode, h->VBICbaseNode, h->VBICemitNode, h->VBICsubsNode, h->VBICcollCXNode,
                h->VBICcollCINode, h->VBICbaseBXNode, h->VBICbaseBINode, h->VBICemitEINode, h->VBICbaseBPNode, h->VBICsubsSINode,
                        (h->VBIC_selfheat && h->VBICtempNode > 0) ? h->VBICtempNode : 0, h->VBIC_excessPhase ? h->VBICxf1Node : 0, h->VBIC_excessPhase ? h->VBICxf2Node : 0 };
            static const struct { int k; size_t off; } upd[] = {
#define U(kk, f) { kk, offsetof(VBICinstance, f) }
                U(1, VBICtextCollResist), U(2, VBICtintCollResist), U(3, VBICtepiSatVoltage), U(4, VBICtepiDoping), U(6, VBICtextBaseResist),
                U(7, VBICtintBaseResist), U(8, VBICtemitterResist), U(9, VBICtsubstrateResist), U(10, VBICtparBaseResist), U(11, VBICtsatCur),
                U(12, VBICtemissionCoeffF), U(13, VBICtemissionCoeffR), U(16, VBICtdepletionCapBE), U(17, VBICtpotentialBE),
                U(21, VBICtdepletionCapBC), U(23, VBICtextCapBC), U(24, VBICtpotentialBC), U(27, VBICtextCapSC), U(28, VBICtpotentialSC),
                U(31, VBICtidealSatCurBE), U(34, VBICtnidealSatCurBE), U(36, VBICtidealSatCurBC), U(38, VBICtnidealSatCurBC),
                U(41, VBICtavalanchePar2BC), U(42, VBICtparasitSatCur), U(45, VBICtidealParasitSatCurBE), U(46, VBICtnidealParasitSatCurBE),
                U(47, VBICtidealParasitSatCurBC), U(49, VBICtnidealParasitSatCurBC), U(53, VBICtrollOffF), U(94, VBICtsepISRR),
                U(98, VBICtvbbe), U(99, VBICtnbbe)
#undef U
            };
            for (k = 0; k < VBN_COUNT; k++) nodes[(size_t)k * n + i] = nd[k];
            memcpy(p, &m->VBICtnom, sizeof p);
            p[0] = h->VBICtemp - CONSTCtoK + p[105];
            for (k = 0; k < (int)(sizeof upd / sizeof upd[0]); k++) p[upd[k].k] = *(const double *)((const char *)h + upd[k].off);
            for (k = 0; k < VBIC_NP; k++) par[(size_t)k * n + i] = p[k];
            aux[(size_t)VBA_type * n + i] = m->VBICtype; aux[(size_t)VBA_tVcrit * n + i] = h->VBICtVcrit;
            aux[(size_t)VBA_icVBE * n + i] = h->VBICicVBE; aux[(size_t)VBA_icVCE * n + i] = h->VBICicVCE;
            aux[(size_t)VBA_scale * n + i] = h->VBICarea * h->VBICm; aux[(size_t)VBA_temp * n + i] = h->VBICtemp;
            flags[i] = (h->VBICoff ? VBF_OFF : 0) | ((h->VBIC_selfheat && h->VBICtempNode > 0) ? VBF_SELFHEAT : 0) | (h->VBIC_excessPhase ? VBF_EXCESS : 0);
            G.sbq[i] = h->VBICstate;
        }
    rc = ngbCircuitAddVbic(G.C, n, nodes, flags, par, aux);
    free(nodes); free(flags); free(par); free(aux);
    return rc;
}

static int flatten_linear(CKTcircuit *ckt)
{
    int n, i, rc = 0, k;
    COUNT(RES, G.tRES, n);
    if (n) {
        RESmodel *m; RESinstance *h; int *nodes = (int *)xc((size_t)n * 2, sizeof(int)); double *g = (double *)xc((size_t)n, sizeof(double));
        i = 0;
        for (m = (RESmodel *)ckt->CKThead[G.tRES]; m; m = RESnextModel(m))
            for (h = RESinstances(m); h; h = RESnextInstance(h), i++) { nodes[i] = h->RESposNode; nodes[n + i] = h->RESnegNode; g[i] = h->RESconduct; }
        rc = ngbCircuitAddResistors(G.C, n, nodes, g); free(nodes); free(g);
        if (rc) return rc;
    }
    n = G.nc;
    if (n) {
        CAPmodel *m; CAPinstance *h; int *nodes = (int *)xc((size_t)n * 2, sizeof(int)); double *par = (double *)xc((size_t)n * 3, sizeof(double));
        G.sbc = (int *)xc((size_t)n, sizeof(int));
        i = 0;
        for (m = (CAPmodel *)ckt->CKThead[G.tCAP]; m; m = CAPnextModel(m))
            for (h = CAPinstances(m); h; h = CAPnextInstance(h), i++) {
                nodes[i] = h->CAPposNode; nodes[n + i] = h->CAPnegNode;
                par[i] = h->CAPcapac; par[n + i] = h->CAPm; par[2 * n + i] = h->CAPinitCond; G.sbc[i] = h->CAPstate; }
        rc = ngbCircuitAddCapacitors(G.C, n, nodes, par); free(nodes); free(par);
        if (rc) return rc;
    }
    COUNT(VSRC, G.tVSRC, n);
    if (n) {
        VSRCmodel *m; VSRCinstance *h; int *nodes = (int *)xc((size_t)n * 3, sizeof(int)), *fn = (int *)xc((size_t)n * 3, sizeof(int));
        double *par = (double *)xc((size_t)n * 9, sizeof(double));
        i = 0;
        for (m = (VSRCmodel *)ckt->CKThead[G.tVSRC]; m; m = VSRCnextModel(m))
            for (h = VSRCinstances(m); h; h = VSRCnextInstance(h), i++) {
                nodes[i] = h->VSRCposNode; nodes[n + i] = h->VSRCnegNode; nodes[2 * n + i] = h->VSRCbranch;
                fn[i] = h->VSRCfunctionType; fn[n + i] = h->VSRCfunctionOrder; fn[2 * n + i] = h->VSRCdcGiven;
                par[i] = h->VSRCdcValue;
                for (k = 0; k < 8; k++) par[(size_t)(1 + k) * n + i] = (h->VSRCcoeffs && k < h->VSRCfunctionOrder) ? h->VSRCcoeffs[k] : 0.0;
            }
        rc = ngbCircuitAddVsources(G.C, n, nodes, fn, par); free(nodes); free(fn); free(par);
        if (rc) return rc;
        i = 0;
        for (m = (VSRCmodel *)ckt->CKThead[G.tVSRC]; m; m = VSRCnextModel(m))
            for (h = VSRCinstances(m); h; h = VSRCnextInstance(h), i++)
                if (h->VSRCfunctionType == PWL &&
                    (rc = ngbCircuitSetVsourcePwl(G.C, i, h->VSRCfunctionOrder, h->VSRCcoeffs, h->VSRCrdelay, h->VSRCrGiven ? h->VSRCrBreakpt : -1)))
                    return rc;
    }
    COUNT(ISRC, G.tISRC, n);
    if (n) {
        ISRCmodel *m; ISRCinstance *h; int *nodes = (int *)xc((size_t)n * 2, sizeof(int)), *fn = (int *)xc((size_t)n * 3, sizeof(int));
        double *par = (double *)xc((size_t)n * 10, sizeof(double));
        i = 0;
        for (m = (ISRCmodel *)ckt->CKThead[G.tISRC]; m; m = ISRCnextModel(m))
            for (h = ISRCinstances(m); h; h = ISRCnextInstance(h), i++) {
                nodes[i] = h->ISRCposNode; nodes[n + i] = h->ISRCnegNode;
                fn[i] = h->ISRCfunctionType; fn[n + i] = h->ISRCfunctionOrder; fn[2 * n + i] = h->ISRCdcGiven;
                par[i] = h->ISRCdcValue; par[n + i] = h->ISRCmValue;
                for (k = 0; k < 8; k++) par[(size_t)(2 + k) * n + i] = (h->ISRCcoeffs && k < h->ISRCfunctionOrder) ? h->ISRCcoeffs[k] : 0.0;
            }
        rc = ngbCircuitAddIsources(G.C, n, nodes, fn, par); free(nodes); free(fn); free(par);
        if (rc) return rc;
        i = 0;
        for (m = (ISRCmodel *)ckt->CKThead[G.tISRC]; m; m = ISRCnextModel(m))
            for (h = ISRCinstances(m); h; h = ISRCnextInstance(h), i++)
                if (h->ISRCfunctionType == PWL && (rc = ngbCircuitSetIsourcePwl(G.C, i, h->ISRCfunctionOrder, h->ISRCcoeffs)))
                    return rc;
    }
    return rc;
}

static void table_install(CKTcircuit *ckt);
static int shim_attach(CKTcircuit *ckt)
{
    const char *e = getenv("NGB_SHIM");
    KLUmatrix *K;
    CKTnode *node;
    int t, neq = ckt->CKTmaxEqNum, rc, n, nnz, nrows;
    int *ntype;
    double dopt[15]; int iopt[5];
    G.tried = 1; G.ckt = ckt;
    if (e && !strcmp(e, "0")) return shim_fail("NGB_SHIM=0");
    if (!ckt->CKTmatrix || !ckt->CKTmatrix->CKTkluMODE) return shim_fail("matrix is not in KLU mode (.option klu)");
    if (ckt->CKTbypass) return shim_fail("CKTbypass is on");
    K = ckt->CKTmatrix->SMPkluMatrix;
    G.tB4 = CKTtypelook("BSIM4"); G.tB3 = CKTtypelook("BSIM3"); G.tDIO = CKTtypelook("Diode"); G.tVBIC = CKTtypelook("VBIC"); G.tRES = CKTtypelook("Resistor");
    G.tCAP = CKTtypelook("Capacitor"); G.tVSRC = CKTtypelook("Vsource"); G.tISRC = CKTtypelook("Isource");
    e = getenv("NGB_SHIM_TABLE");
    G.table = (e && !strcmp(e, "1")) ? 1 : 0;
    for (t = 0; t < DEVmaxnum; t++)
        if (DEVices[t] && ckt->CKThead[t] && DEVices[t]->DEVload &&
            t != G.tB4 && t != G.tB3 && t != G.tDIO && t != G.tVBIC && t != G.tRES && t != G.tCAP && t != G.tVSRC && t != G.tISRC) {
            fprintf(stderr, "ngb_shim: device type %s stays on the CPU (its own DEVload)\n", DEVices[t]->DEVpublic.name);
            G.table = 1; G.mixed = 1;
        }
    ntype = (int *)xc((size_t)neq + 1, sizeof(int));
    for (node = ckt->CKTnodes; node; node = node->next) if (node->number >= 0 && node->number <= neq) ntype[node->number] = node->type;
    G.C = ngbCircuitCreate(neq, ntype);
    free(ntype);
    memset(dopt, 0, sizeof dopt); memset(iopt, 0, sizeof iopt);
    dopt[0] = ckt->CKTreltol; dopt[1] = ckt->CKTabstol; dopt[2] = ckt->CKTvoltTol; dopt[3] = ckt->CKTchgtol; dopt[4] = ckt->CKTtrtol;
    dopt[5] = ckt->CKTtemp; dopt[6] = CONSTvt0; dopt[7] = ckt->CKTxmu; dopt[8] = ckt->CKTstep; dopt[9] = ckt->CKTfinalTime;
    dopt[10] = ckt->CKTmaxStep; dopt[11] = ckt->CKTinitTime; dopt[12] = ckt->CKTdelmin; dopt[13] = ckt->CKTminBreak; dopt[14] = ckt->CKTgmin;
    iopt[0] = ckt->CKTintegrateMethod; iopt[1] = ckt->CKTmaxOrder; iopt[2] = ckt->CKTtranMaxIter; iopt[3] = ckt->CKTdcMaxIter;
    iopt[4] = (ckt->CKTmode & MODEUIC) ? 1 : 0;
    ngbCircuitSetOptions(G.C, dopt, iopt);
    if ((rc = ngbCircuitSetOpFallbacks(G.C, ckt->CKTnumGminSteps, ckt->CKTnumSrcSteps, ckt->CKTdcTrcvMaxIter, ckt->CKTgminFactor, ckt->CKTnoOpIter, ckt->CKTgshunt))) return shim_fail(ngbLastError());
    COUNT(BSIM4, G.tB4, G.n4); COUNT(BSIM3, G.tB3, G.n3); COUNT(DIO, G.tDIO, G.nd); COUNT(VBIC, G.tVBIC, G.nq); COUNT(CAP, G.tCAP, G.nc);
    if ((rc = flatten_bsim3(ckt)) || (rc = flatten_bsim4(ckt)) || (rc = flatten_dio(ckt)) || (rc = flatten_vbic(ckt)) || (rc = flatten_linear(ckt)) ||
        (rc = ngbCircuitFinalize(G.C))) {
        fprintf(stderr, "ngb_shim: %s\n", ngbLastError());
        return shim_fail("circuit uses an option outside the GPU path");
    }
    if (!G.table) {   /* .nodeset / .ic rows (cktload.c:118-172): nodesets first, then initial conditions, each in node order;
                       * in table mode the reference's CKTload applies them itself */
        int nov = 0, pass, i = 0, *oeq, *okind; double *oval;
        for (node = ckt->CKTnodes; node; node = node->next) nov += (node->nsGiven ? 1 : 0) + (node->icGiven ? 1 : 0);
        if (nov) {
            oeq = (int *)xc((size_t)nov, sizeof(int)); okind = (int *)xc((size_t)nov, sizeof(int)); oval = (double *)xc((size_t)nov, sizeof(double));
            for (pass = 0; pass < 2; pass++)
                for (node = ckt->CKTnodes; node; node = node->next)
                    if (pass ? node->icGiven : node->nsGiven) { oeq[i] = node->number; okind[i] = pass; oval[i] = pass ? node->ic : node->nodeset; i++; }
            rc = ngbCircuitSetNodeOverrides(G.C, nov, oeq, okind, oval);
            free(oeq); free(okind); free(oval);
            if (rc) { fprintf(stderr, "ngb_shim: %s\n", ngbLastError()); return shim_fail("node overrides rejected"); }
        }
    }
    ngbCircuitPatternSize(G.C, &n, &nnz, &nrows);
    if (!G.mixed && (n != (int)K->KLUmatrixN || nnz != (int)K->KLUmatrixNZ)) return shim_fail("CSC pattern differs from SMPconvertCOOtoCSC's");
    {
        int *Ap = (int *)xc((size_t)n + 1, sizeof(int)), *Ai = (int *)xc((size_t)nnz, sizeof(int)), *dg = (int *)xc((size_t)n, sizeof(int)), same;
        ngbCircuitGetPattern(G.C, Ap, Ai, dg);
        same = n == (int)K->KLUmatrixN && nnz == (int)K->KLUmatrixNZ &&
               !memcmp(Ap, K->KLUmatrixAp, sizeof(int) * ((size_t)n + 1)) && !memcmp(Ai, K->KLUmatrixAi, sizeof(int) * (size_t)nnz);
        if (!same && G.mixed) {
            /* the replaced types' entries are a subset of the reference's pattern: find every slot of ours in it */
            int col, q, bad = 0, *eqn = (int *)xc((size_t)n, sizeof(int));
            ngbCircuitGetPatternEquations(G.C, eqn);          /* our index k is equation eqn[k]; the reference's is equation - 1 */
            G.slot_map = (int *)xc((size_t)nnz, sizeof(int));
            for (col = 0; col < n && !bad; col++)
                for (q = Ap[col]; q < Ap[col + 1]; q++) {
                    const int kc = eqn[col] - 1, kr = eqn[Ai[q]] - 1;
                    int lo, hi = -1;
                    if (kc < (int)K->KLUmatrixN)
                        for (lo = K->KLUmatrixAp[kc]; lo < K->KLUmatrixAp[kc + 1]; lo++) if (K->KLUmatrixAi[lo] == kr) { hi = lo; break; }
                    if (hi < 0) { bad = 1; break; }
                    G.slot_map[q] = hi;
                }
            free(eqn);
            if (bad) { free(Ap); free(Ai); free(dg); return shim_fail("an entry of the replaced device types is missing from the reference's CSC pattern"); }
            same = 1;
        }
        free(Ap); free(Ai); free(dg);
        if (!same) return shim_fail("CSC pattern differs from SMPconvertCOOtoCSC's");
    }
    if (ngbInit(0)) { fprintf(stderr, "ngb_shim: %s\n", ngbLastError()); return shim_fail("no CUDA device"); }
    G.B = ngbBatchCreate(G.C, 1, 0);
    if (!G.B) { fprintf(stderr, "ngb_shim: %s\n", ngbLastError()); return shim_fail("batch creation failed"); }
    G.neq = neq; G.nnz = nnz;
    e = getenv("NGB_SHIM_LU");
    G.use_lu = !(e && !strcmp(e, "0")) && !G.mixed;       /* CPU device types: their entries are not in the device pattern */
    G.active = 1;
    if (G.table) table_install(ckt);
    fprintf(stderr, "ngb_shim: %s%s on %s (%d BSIM4, %d BSIM3, %d diodes, %d VBIC, %d unknowns, %d nonzeros)\n",
            G.table ? "CKTload through the SPICEdev table (DEVload of the replaced types)" : "CKTload",
            G.use_lu ? " + SMPluFac + SMPsolve" : "", ngbBackend(), G.n4, G.n3, G.nd, G.nq, n, nnz);
    return 1;
}

/* ------------------------------------------------------------------ per-call data movement */
static void states_down(const char *name, int K, int n, const int *base)
{
    CKTcircuit *ckt = G.ckt;
    double *buf; int h, k, i;
    if (!n) return;
    buf = scratch((size_t)4 * K * n);
    memset(buf, 0, sizeof(double) * 4 * (size_t)K * n);
    for (h = 0; h < 4; h++) {
        const double *st = (h <= ckt->CKTmaxOrder + 1) ? ckt->CKTstates[h] : NULL;
        if (!st) continue;
        for (k = 0; k < K; k++) for (i = 0; i < n; i++) buf[((size_t)h * K + k) * n + i] = st[base[i] + k];
    }
    ngbBatchUpload(G.B, name, buf, (long)(sizeof(double) * 4 * (size_t)K * n), 0);
}
static void states_up(const char *name, int K, int n, const int *base)
{
    CKTcircuit *ckt = G.ckt;
    double *buf; int h, k, i;
    if (!n) return;
    buf = scratch((size_t)2 * K * n);
    ngbBatchDownload(G.B, name, buf, (long)(sizeof(double) * 2 * (size_t)K * n), 0);
    for (h = 0; h < 2; h++) {
        double *st = ckt->CKTstates[h];
        if (!st) continue;
        for (k = 0; k < K; k++) for (i = 0; i < n; i++) st[base[i] + k] = buf[((size_t)h * K + k) * n + i];
    }
}

/* CKTload's state -> device, ngbLoad, results back.  into_matrix = 0: Ax and CKTrhs are the device's (wrap mode, every
 * device type on the device); 1: the assembled contributions are ADDED to what the reference's CKTload holds (table mode) */
static int device_load(CKTcircuit *ckt, int into_matrix)
{
    int iv, rc; double dv;
#define PUT_I(name, v) do { iv = (v); ngbBatchUpload(G.B, name, &iv, sizeof(int), 0); } while (0)
#define PUT_D(name, v) do { dv = (v); ngbBatchUpload(G.B, name, &dv, sizeof(double), 0); } while (0)
    PUT_I("ctl.mode", (int)ckt->CKTmode); PUT_I("ctl.active", 1); PUT_I("ctl.head", 0); PUT_I("ctl.order", ckt->CKTorder);
    PUT_I("ctl.xsel", 0); PUT_I("ctl.stateop", 0); PUT_I("ctl.err", 0);
    PUT_D("ctl.ag0", ckt->CKTag[0]); PUT_D("ctl.ag1", ckt->CKTag[1]); PUT_D("ctl.ag2", ckt->CKTag[2]); PUT_D("ctl.delta", ckt->CKTdelta); PUT_D("ctl.time", ckt->CKTtime);
    PUT_D("ctl.gmin", ckt->CKTgmin); PUT_D("ctl.srcfact", ckt->CKTsrcFact);
    PUT_D("ctl.diag_gmin", 0.0);        /* LoadGmin_CSC stays with whoever factors; see __wrap_SMPluFac */
    ngbBatchUpload(G.B, "ctl.delta_old", ckt->CKTdeltaOld, sizeof(double) * 7, 0);
    ngbBatchUpload(G.B, "x", ckt->CKTrhsOld, (long)(sizeof(double) * ((size_t)G.neq + 1)), 0);
    states_down("b4.state", B4ST_COUNT, G.n4, G.sb4);
    states_down("b3.state", B3ST_COUNT, G.n3, G.sb3);
    states_down("dio.state", DIOST_COUNT, G.nd, G.sbd);
    states_down("vbic.state", VBS_COUNT, G.nq, G.sbq);
    states_down("cap.state", 2, G.nc, G.sbc);
    rc = ngbLoad(G.B);
    if (rc) { fprintf(stderr, "ngb_shim: ngbLoad failed (%d): %s\n", rc, ngbLastError()); return rc; }
    if (!into_matrix) {
        ngbBatchDownload(G.B, "Ax", ckt->CKTmatrix->SMPkluMatrix->KLUmatrixAx, (long)(sizeof(double) * (size_t)G.nnz), 0);
        ngbBatchDownload(G.B, "x", ckt->CKTrhs, (long)(sizeof(double) * ((size_t)G.neq + 1)), (long)(sizeof(double) * ((size_t)G.neq + 1)));
        ckt->CKTrhs[0] = 0.0;
    } else {
        double *Ax = ckt->CKTmatrix->SMPkluMatrix->KLUmatrixAx, *buf = scratch((size_t)G.nnz + (size_t)G.neq + 1);
        int k;
        ngbBatchDownload(G.B, "Ax", buf, (long)(sizeof(double) * (size_t)G.nnz), 0);
        ngbBatchDownload(G.B, "x", buf + G.nnz, (long)(sizeof(double) * ((size_t)G.neq + 1)), (long)(sizeof(double) * ((size_t)G.neq + 1)));
        if (G.slot_map) for (k = 0; k < G.nnz; k++) Ax[G.slot_map[k]] += buf[k];
        else for (k = 0; k < G.nnz; k++) Ax[k] += buf[k];
        for (k = 1; k <= G.neq; k++) ckt->CKTrhs[k] += buf[G.nnz + k];
    }
    states_up("b4.state", B4ST_COUNT, G.n4, G.sb4);
    states_up("b3.state", B3ST_COUNT, G.n3, G.sb3);
    states_up("dio.state", DIOST_COUNT, G.nd, G.sbd);
    states_up("vbic.state", VBS_COUNT, G.nq, G.sbq);
    states_up("cap.state", 2, G.nc, G.sbc);
    ngbBatchDownload(G.B, "ctl.noncon", &iv, sizeof(int), 0);
    ckt->CKTnoncon += iv;
    G.dev_factored = 0;
    G.loads++;
    return OK;
}

/* ------------------------------------------------------------------ table mode: DEVices[t]->DEVload replacements */
static int table_load(int type, GENmodel *head, CKTcircuit *ckt)
{
    int k;
    if (!G.active || ckt != G.ckt) {          /* another circuit: the type's own DEVload */
        for (k = 0; k < G.norig; k++) if (G.orig_type[k] == type) return G.orig_load[k](head, ckt);
        return E_PANIC;
    }
    if (type != G.leader) return OK;          /* evaluated together with the leader's pass */
    return device_load(ckt, 1);
}
#define TABLE_FN(nm, field) static int table_load_##nm(GENmodel *m, CKTcircuit *ckt) { return table_load(G.field, m, ckt); }
TABLE_FN(b4, tB4) TABLE_FN(b3, tB3) TABLE_FN(dio, tDIO) TABLE_FN(vbic, tVBIC) TABLE_FN(res, tRES) TABLE_FN(cap, tCAP) TABLE_FN(vsrc, tVSRC) TABLE_FN(isrc, tISRC)

static void table_install(CKTcircuit *ckt)
{
    struct { int type; int (*fn)(GENmodel *, CKTcircuit *); } rep[8] = {
        { G.tB4, table_load_b4 }, { G.tB3, table_load_b3 }, { G.tDIO, table_load_dio }, { G.tVBIC, table_load_vbic },
        { G.tRES, table_load_res }, { G.tCAP, table_load_cap }, { G.tVSRC, table_load_vsrc }, { G.tISRC, table_load_isrc } };
    int k;
    G.leader = -1; G.norig = 0;
    for (k = 0; k < 8; k++) {
        const int t = rep[k].type;
        if (t < 0 || !DEVices[t] || !ckt->CKThead[t]) continue;
        G.orig_type[G.norig] = t; G.orig_load[G.norig] = DEVices[t]->DEVload; G.norig++;
        DEVices[t]->DEVload = rep[k].fn;
        if (G.leader < 0 || t < G.leader) G.leader = t;
    }
}

int __wrap_CKTload(CKTcircuit *ckt)
{
    int rc;
    double startTime;
    if (!G.tried) shim_attach(ckt);
    if (!G.active || ckt != G.ckt || G.table) return __real_CKTload(ckt);      /* table mode: the reference's CKTload, replaced DEVloads inside */
    startTime = SPfrontEnd->IFseconds();
    rc = device_load(ckt, 0);
    ckt->CKTstat->STATloadTime += SPfrontEnd->IFseconds() - startTime;
    return rc;
}

/* pattern + pivot order of the pivoting factor that just ran on the host -> device schedule */
static int pattern_to_device(KLUmatrix *K)
{
    klu_symbolic *Sy = K->KLUmatrixSymbolic; klu_numeric *Nu = K->KLUmatrixNumeric;
    int n = Sy->n, nb = Sy->nblocks, lnz = 0, unz = 0, b, k, pl = 0, pu = 0, rc;
    int *Lp = (int *)xc((size_t)n + 1, sizeof(int)), *Up = (int *)xc((size_t)n + 1, sizeof(int)), *Li, *Ui;
    for (b = 0; b < nb; b++) {
        int k1 = Sy->R[b], nk = Sy->R[b + 1] - k1;
        if (nk > 1) for (k = 0; k < nk; k++) { lnz += Nu->Llen[k1 + k]; unz += Nu->Ulen[k1 + k]; }
    }
    Li = (int *)xc((size_t)lnz + 1, sizeof(int)); Ui = (int *)xc((size_t)unz + 1, sizeof(int));
    for (b = 0; b < nb; b++) {
        int k1 = Sy->R[b], nk = Sy->R[b + 1] - k1;
        for (k = 0; k < nk; k++) {
            Lp[k1 + k] = pl; Up[k1 + k] = pu;
            if (nk > 1) {
                double *LU = (double *)Nu->LUbx[b];
                int *Lip = Nu->Lip + k1, *Llen = Nu->Llen + k1, *Uip = Nu->Uip + k1, *Ulen = Nu->Ulen + k1, q;
                int *li = (int *)(LU + Lip[k]), *ui = (int *)(LU + Uip[k]);
                for (q = 0; q < Llen[k]; q++) Li[pl++] = li[q] + k1;
                for (q = 0; q < Ulen[k]; q++) Ui[pu++] = ui[q] + k1;
            }
        }
    }
    Lp[n] = pl; Up[n] = pu;
    ngbCircuitSelectLuSet(G.C, 0);
    rc = ngbCircuitSetLuPattern(G.C, n, nb, Sy->Q, Sy->R, Nu->Pnum, Lp, Li, Up, Ui, Nu->Offp, Nu->Offi);
    free(Lp); free(Up); free(Li); free(Ui);
    if (!rc) rc = ngbBatchRefreshLu(G.B);
    if (rc) fprintf(stderr, "ngb_shim: LU pattern rejected (%d): %s -- LU stays on the host\n", rc, ngbLastError());
    return rc;
}

int __wrap_SMPreorder(SMPmatrix *M, double PivTol, double PivRel, double Gmin)
{
    int r = __real_SMPreorder(M, PivTol, PivRel, Gmin);
    if (G.active && G.use_lu && M == G.ckt->CKTmatrix && r == 0) {
        const int gmin_loaded = (Gmin != 0.0 && M->SMPkluMatrix->KLUloadDiagGmin);   /* host Ax now differs from the device copy */
        G.have_lu = (pattern_to_device(M->SMPkluMatrix) == 0);
        G.dev_factored = 0;
        if (G.have_lu && !gmin_loaded) {
            /* the device refactors the same Ax on the new order, so the solve that follows runs there too */
            int rc;
            if (G.table) ngbBatchUpload(G.B, "Ax", M->SMPkluMatrix->KLUmatrixAx, (long)(sizeof(double) * (size_t)G.nnz), 0);
            rc = ngbLuFac(G.B);
            if (rc == 0) { G.dev_factored = 1; G.facs++; }
        }
    }
    return r;
}

int __wrap_SMPluFac(SMPmatrix *M, double PivTol, double Gmin)
{
    if (G.active && G.use_lu && G.have_lu && M == G.ckt->CKTmatrix &&
        !(Gmin != 0.0 && M->SMPkluMatrix->KLUloadDiagGmin)) {      /* gmin stepping: the host factors (it owns LoadGmin_CSC) */
        int rc;
        if (G.table) ngbBatchUpload(G.B, "Ax", M->SMPkluMatrix->KLUmatrixAx, (long)(sizeof(double) * (size_t)G.nnz), 0);   /* what the host CKTload finished */
        rc = ngbLuFac(G.B);
        G.facs++;
        if (rc == 0) { G.dev_factored = 1; return 0; }
        G.dev_factored = 0;
        if (rc == 102) return E_SINGULAR;
        fprintf(stderr, "ngb_shim: ngbLuFac failed (%d): %s\n", rc, ngbLastError());
        return rc;
    }
    return __real_SMPluFac(M, PivTol, Gmin);
}

void __wrap_SMPsolve(SMPmatrix *M, double RHS[], double Spare[])
{
    if (G.active && G.use_lu && G.dev_factored && M == G.ckt->CKTmatrix && RHS == G.ckt->CKTrhs) {
        int rc;
        if (G.table) ngbBatchUpload(G.B, "x", RHS, (long)(sizeof(double) * ((size_t)G.neq + 1)), (long)(sizeof(double) * ((size_t)G.neq + 1)));
        rc = ngbSolve(G.B);
        if (rc) { fprintf(stderr, "ngb_shim: ngbSolve failed (%d): %s\n", rc, ngbLastError()); }
        ngbBatchDownload(G.B, "x", RHS, (long)(sizeof(double) * ((size_t)G.neq + 1)), (long)(sizeof(double) * ((size_t)G.neq + 1)));
        RHS[0] = 0.0;
        G.solves++;
        return;
    }
    __real_SMPsolve(M, RHS, Spare);
}
