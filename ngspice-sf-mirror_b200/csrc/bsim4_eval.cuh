/* bsim4_eval.cuh -- BSIM4.8.3 instance evaluation for one (instance, sample) thread.
 *
 * What it replaces: the body of BSIM4load / BSIM4LoadOMP
 * (src/spicelib/devices/bsim4/b4ld.c:94-5399) for one instance -- terminal-voltage fetch by
 * INITF mode (:261-407), Newton step limiting (:605-686), S/D junction diodes (:700-1002),
 * threshold/mobility/Vdsat/Ids (:1004-2189), Rg, bias-dependent Rs/Rd (:2191-2306),
 * GIDL/GISL (:2308-2523), gate tunnelling (:2527-2920), intrinsic charges capMod 0/1/2
 * (:3014-3913), junction C-V (:3946-4063), overlap charges + capacitance matrix
 * (:4132-4567), NIintegrate (:4582-4651), equivalent currents (:4689-4743), and the
 * RHS/matrix contributions (:4750-5388) -- the latter emitted as one value per stamp
 * position (bsim4_fields.h) instead of being added through pointers.
 *
 * How it differs structurally: structure-of-arrays operands indexed by thread, phases as
 * separate inlines sharing a register-resident work struct, no linked lists, no pointer
 * stamps.  The arithmetic order of every model equation follows the reference so that
 * results agree to rounding with the CPU path; compile with -fmad=false for the same
 * reason (x86-64 gcc does not contract a*b+c).
 *
 * Not implemented (rejected at upload with NGB_E_UNSUPP, never silently ignored):
 * trnqsMod/acnqsMod != 0 (NQS), MODEINITSMSIG (AC small-signal), CKTbypass != 0.
 */
#ifndef NGB_BSIM4_EVAL_CUH
#define NGB_BSIM4_EVAL_CUH

#include "ngb_common.h"
#include "ngb_types.h"
#include "bsim4_fields.h"
#include "devsup.cuh"

/* Optional CTA-wide barrier at each phase boundary of the evaluation (keeps the warps of a CTA in
 * one instruction-cache window).  Measured on B200: no gain (571 vs 572 us per launch), so it is
 * off unless NGB_B4_PHASE_BARRIERS is defined. */
#if defined(__CUDA_ARCH__) && defined(NGB_B4_PHASE_BARRIERS)
#define NGB_CTA_ALIGN() __syncthreads()
#else
#define NGB_CTA_ALIGN() ((void)0)
#endif

#define B4_MAX_EXPL 2.688117142e+43
#define B4_MIN_EXPL 3.720075976e-44
#define B4_EXPL_THRESHOLD 100.0
#define B4_MAX_EXP 5.834617425e14
#define B4_MIN_EXP 1.713908431e-15
#define B4_EXP_THRESHOLD 34.0
#define B4_EPS0 8.85418e-12
#define B4_EPSSI 1.03594e-10
#define B4_CHARGE_Q 1.60219e-19
#define B4_DELTA_1 0.02
#define B4_DELTA_3 0.02
#define B4_DELTA_4 0.02
#define B4_KCHARGE 1.6021766208e-19   /* CHARGE, src/include/ngspice/const.h:32 */

/* launch-wide operands of the BSIM4 load (device pointers unless noted) */
typedef struct B4Ctx {
    int ninst;                 /* instances per circuit                                  */
    int S;                     /* samples (independent circuits) in the batch            */
    int T;                     /* ninst * S ; thread t = inst * S + sample               */
    const double *mtab;        /* [nrows][B4M_COUNT] model rows                          */
    const double *ptab;        /* [nrows][B4P_COUNT] bin rows                            */
    const int *prow;           /* parameter row per thread ([T]) or per instance ([ninst]) */
    int prow_per_thread;       /* 1: prow[t], 0: prow[inst]                              */
    int overlay;               /* per-thread rows read as an OVERLAY: Mrow / Prow are the rows of the instance's sample 0 and only the
                                * columns that differ between the samples of an instance are read at the thread's own row (below) */
    int mvary[B4M_COUNT];      /* byte pitch of a model row for the columns that differ between samples, 0 for the others */
    int pvary[B4P_COUNT];      /* same for the bin rows                                  */
    unsigned variant;          /* variant key of the batch (bsim4_variants.h), NGB_B4_GENERIC when its instances differ */
    const double *inst;        /* [B4I_COUNT][T]                                         */
    const int *flags;          /* [ninst] packed B4F_*                                   */
    const int *nodes;          /* [B4N_COUNT][ninst] equation numbers (0 = ground)       */
    const int *spos;           /* [B4S_COUNT][ninst] stamp row or -1                     */
    int srow0;                 /* rows are position-major: position k of instance i is row srow0 + k * ninst + i, live or not */
    double *stamp;             /* [nrows_stamp][S]                                       */
    double *state;             /* [NGB_NHIST][B4ST_COUNT][T]                             */
    double *op;                /* [B4O_COUNT][T] operating point (von is read back)      */
    int op_full;               /* 0: keep only von ; 1: export every B4O_* (parity runs) */
    int split;                 /* 1: exact-order stamps (B4X_* rows live), 0: merged      */
    int lte_deferred;          /* 1: BSIM4trunc runs in its own launch after the solve, for converged samples only (b4_lte_thread) */
    const int *nodeconv;       /* [S] node part of NIconvTest (LU kernel), read by b4_lte_thread            */
    const double *x;           /* [2][neq1][S] solution buffers; xsel[s] picks CKTrhsOld  */
    int neq1;                  /* equations + 1 (row 0 is ground)                         */
    NgbCtl ctl;                /* per-sample control block                                */
    /* shared scalars */
    double temp;               /* CKTtemp                                                */
    double vt0;                /* CONSTvt0                                               */
} B4Ctx;

/* values shared between the evaluation phases of one thread.  ~90 of them are written by an early phase and read once
 * by the finish phase; keeping the first 16 / 32 / 48 of those in shared memory ([field][thread], one LDS / STS per
 * access) instead of leaving them to the register allocator was measured on B200 and changes nothing (508 / 507 / 519
 * against 509 us per Newton step; profiles/README.md, round 2), so they are plain members */
typedef struct B4W {
    /* limited terminal voltages (NMOS polarity) */
    double vds, vgs, vbs, vbd, vgd, vgb, vges, vgms, vged, vgmd, vgmb;
    double vdbs, vdbd, vsbs, vses, vdes, qdef, vbs_jct, vbd_jct;
    int Check;
    /* orientation-normalised bias */
    int mode;
    double Vds, Vgs, Vbs, Vdb;
    /* junction */
    double gbs, cbs, gbd, cbd;
    /* core */
    double toxe, epsrox, epssub, Vtm, Vtm0, Leff;
    double Vbseff, dVbseff_dVb, Phis, dPhis_dVb, sqrtPhis, dsqrtPhis_dVb;
    double Vth, dVth_dVb, dVth_dVd, n, dn_dVb, dn_dVd, von;
    double vgs_eff, vgd_eff, dvgs_eff_dvg, dvgd_eff_dvg, Vgs_eff, dVgs_eff_dVg, Vgst;
    double Vgsteff, dVgsteff_dVg, dVgsteff_dVd, dVgsteff_dVb;
    double Vdseff, dVdseff_dVg, dVdseff_dVd, dVdseff_dVb;
    double Abulk0_Q, dAbulk0_Q_dVb;
    double cdrain, gm, gds, gmbs, csub, gbbs, gbgs, gbds, vdsat;
    double beta, dbeta_dVg, dbeta_dVd, dbeta_dVb, Ids, tmp1, tmp2, tmp3;
    /* parasitics */
    double gcrg, gcrgd, gcrgg, gcrgs, gcrgb;
    double gstot, gstotd, gstotg, gstots, gstotb, gdtot, gdtotd, gdtotg, gdtots, gdtotb;
    double Igidl, ggidld, ggidlg, ggidlb, ggidls, Igisl, ggisls, ggislg, ggislb, ggisld;
    double Igcs, gIgcsg, gIgcsd, gIgcsb, gIgcss, Igcd, gIgcdg, gIgcdd, gIgcdb, gIgcds;
    double Igs, gIgsg, gIgss, Igd, gIgdg, gIgdd, Igb, gIgbg, gIgbd, gIgbb, gIgbs;
    /* charges */
    double qgate, qbulk, qdrn, qsrc;
    double cggb, cgsb, cgdb, cdgb, cdsb, cddb, cbgb, cbsb, cbdb;
} B4W;

#include "bsim4_variants.h"
/* Per-sample parameter rows as an overlay (the generic kernel always, a specialised one when its key says so): of the 221
 * model / bin parameters a process-variation draw changes a handful (oxide thickness: 15), so 32 lanes reading 32 whole rows
 * is 32 sectors per load where one would do.  The host finds the columns that differ between the samples of an instance
 * (ngb_host.c: b4_find_vary); Mrow / Prow point at the rows of the instance's sample 0, b4dr is the distance in rows to the
 * thread's own rows, and a load costs one integer multiply-add with a kernel-constant operand: columns that do not vary
 * are read at the warp-uniform address, the others at the lane's row. */
#define B4OVL ((VK == NGB_B4_GENERIC) || B4K_FIELD(VK, rowsO))
#define B4M(f) NGB_LDG(B4OVL ? (const double *)((const char *)&Mrow[B4M_##f] + (ptrdiff_t)b4dr * c->mvary[B4M_##f]) : &Mrow[B4M_##f])
#define B4P(f) NGB_LDG(B4OVL ? (const double *)((const char *)&Prow[B4P_##f] + (ptrdiff_t)b4dr * c->pvary[B4P_##f]) : &Prow[B4P_##f])
#define B4I(f) NGB_LDG(&c->inst[(size_t)B4I_##f * c->T + t])
/* selectors: compile-time constants of the variant key VK in a specialised instantiation, read from the parameter
 * row / instance flags in the generic one (bsim4_variants.h) */
#define B4SEL(f) ((VK == NGB_B4_GENERIC) ? (int)B4M(f) : B4K_FIELD(VK, f))
#define B4SEL_RBODY(fl) ((VK == NGB_B4_GENERIC) ? B4F_RBODY(fl) : B4K_FIELD(VK, rbodyMod))
#define B4SEL_RGATE(fl) ((VK == NGB_B4_GENERIC) ? B4F_RGATE(fl) : B4K_FIELD(VK, rgateMod))
#define B4SEL_GEAR() ((VK == NGB_B4_GENERIC) ? c->ctl.gear : B4K_FIELD(VK, gear))

#ifdef __cplusplus      /* the plain-C host files only need the declarations above */
/* DEXP of b4ld.c:49-60 */
NGB_HD void b4_dexp(double A, double *B, double *C)
{
    if (A > B4_EXP_THRESHOLD) { *B = B4_MAX_EXP * (1.0 + A - B4_EXP_THRESHOLD); *C = B4_MAX_EXP; }
    else if (A < -B4_EXP_THRESHOLD) { *B = B4_MIN_EXP; *C = 0; }
    else { *B = ngb_exp(A); *C = *B; }
}

/* BSIM4polyDepletion, b4ld.c:5402-5433 */
NGB_HD void b4_poly_depletion(double phi, double ngate, double epsgate, double coxe, double Vgs,
                              double *Vgs_eff, double *dVgs_eff_dVg)
{
    if ((ngate > 1.0e18) && (ngate < 1.0e25) && (Vgs > phi) && (epsgate != 0)) {
        double T1 = 1.0e6 * B4_KCHARGE * epsgate * ngate / (coxe * coxe);
        double T8 = Vgs - phi;
        double T4 = sqrt(1.0 + 2.0 * T8 / T1);
        double T2 = 2.0 * T8 / (T4 + 1.0);
        double T3 = 0.5 * T2 * T2 / T1;
        double T7 = 1.12 - T3 - 0.05;
        double T6 = sqrt(T7 * T7 + 0.224);
        double T5 = 1.12 - 0.5 * (T7 + T6);
        *Vgs_eff = Vgs - T5;
        *dVgs_eff_dVg = 1.0 - (0.5 - 0.5 / T4) * (1.0 + T7 / T6);
    } else {
        *Vgs_eff = Vgs;
        *dVgs_eff_dVg = 1.0;
    }
}

/* one S/D junction diode, DC part (b4ld.c:701-800 source side, :802-901 drain side) */
NGB_HD_SHARED void b4_junction_dc(int dioMod, double Isat, double Nvtm, double vj, double gmin,
                           double bv, double xjbv, double XExpBV, double vjmFwd, double vjmRev,
                           double IVjmFwd, double IVjmRev, double slpFwd, double slpRev,
                           double *g, double *cur)
{
    double T0, T1, T2, T3, ev, dev;
    if (Isat <= 0.0) {
        *g = gmin;
        *cur = *g * vj;
        return;
    }
    switch (dioMod) {
    case 0:
        ev = ngb_exp(vj / Nvtm);
        T1 = xjbv * ngb_exp(-(bv + vj) / Nvtm);
        *g = Isat * (ev + T1) / Nvtm + gmin;
        *cur = Isat * (ev + XExpBV - T1 - 1.0) + gmin * vj;
        break;
    case 1:
        T2 = vj / Nvtm;
        if (T2 < -B4_EXP_THRESHOLD) {
            *g = gmin;
            *cur = Isat * (B4_MIN_EXP - 1.0) + gmin * vj;
        } else if (vj <= vjmFwd) {
            ev = ngb_exp(T2);
            *g = Isat * ev / Nvtm + gmin;
            *cur = Isat * (ev - 1.0) + gmin * vj;
        } else {
            T0 = IVjmFwd / Nvtm;
            *g = T0 + gmin;
            *cur = IVjmFwd - Isat + T0 * (vj - vjmFwd) + gmin * vj;
        }
        break;
    case 2:
        if (vj < vjmRev) {
            T0 = vj / Nvtm;
            if (T0 < -B4_EXP_THRESHOLD) { ev = B4_MIN_EXP; dev = 0.0; }
            else { ev = ngb_exp(T0); dev = ev / Nvtm; }
            T1 = ev - 1.0;
            T2 = IVjmRev + slpRev * (vj - vjmRev);
            *g = dev * T2 + T1 * slpRev + gmin;
            *cur = T1 * T2 + gmin * vj;
        } else if (vj <= vjmFwd) {
            T0 = vj / Nvtm;
            if (T0 < -B4_EXP_THRESHOLD) { ev = B4_MIN_EXP; dev = 0.0; }
            else { ev = ngb_exp(T0); dev = ev / Nvtm; }
            T1 = (bv + vj) / Nvtm;
            if (T1 > B4_EXP_THRESHOLD) { T2 = B4_MIN_EXP; T3 = 0.0; }
            else { T2 = ngb_exp(-T1); T3 = -T2 / Nvtm; }
            *g = Isat * (dev - xjbv * T3) + gmin;
            *cur = Isat * (ev + XExpBV - 1.0 - xjbv * T2) + gmin * vj;
        } else {
            *g = slpFwd + gmin;
            *cur = IVjmFwd + slpFwd * (vj - vjmFwd) + gmin * vj;
        }
        break;
    default:
        /* the reference leaves gbs/cbs untouched here; parameter checking forbids it */
        *g = gmin; *cur = gmin * vj;
        break;
    }
}

/* trap-assisted tunnelling factor and its bias derivative (b4ld.c:911-987, six copies) */
NGB_HD_SHARED void b4_tat(double vts, double vj, double Nvtmr, double *Tn, double *dTn_dVb)
{
    double T0, T9, T10, dT0_dVb;
    if ((vts - vj) < (vts * 1e-3)) {
        T9 = 1.0e3;
        T0 = -vj / Nvtmr * T9;
        b4_dexp(T0, Tn, &T10);
        *dTn_dVb = T10 / Nvtmr * T9;
    } else {
        T9 = 1.0 / (vts - vj);
        T0 = -vj / Nvtmr * vts * T9;
        dT0_dVb = vts / Nvtmr * (T9 + vj * T9 * T9);
        b4_dexp(T0, Tn, &T10);
        *dTn_dVb = T10 * dT0_dVb;
    }
}

/* state accessors: ring-rotated history, [hist][state][thread] */
/* the ring position of history h is head + h (mod nhist); the four per-thread base pointers are formed once per
 * function (B4ST_BASES) with a compare instead of an integer remainder per access */
#define B4ST_BASES(head_) \
    const int b4nh_ = c->ctl.nhist; \
    const size_t b4hs_ = (size_t)B4ST_COUNT * c->T; \
    double *const b4st0 = c->state + (size_t)(head_) * b4hs_ + t; \
    double *const b4st1 = c->state + (size_t)(((head_) + 1 >= b4nh_) ? (head_) + 1 - b4nh_ : (head_) + 1) * b4hs_ + t; \
    double *const b4st2 = c->state + (size_t)(((head_) + 2 >= b4nh_) ? (head_) + 2 - b4nh_ : (head_) + 2) * b4hs_ + t; \
    double *const b4st3 = c->state + (size_t)(((head_) + 3 >= b4nh_) ? (head_) + 3 - b4nh_ : (head_) + 3) * b4hs_ + t; \
    (void)b4st1; (void)b4st2; (void)b4st3; (void)b4st0
#if defined(__CUDA_ARCH__) && defined(NGB_B4_STREAM)
/* experiment switch: the states are read once and written once per evaluation and do not come back before the next
 * Newton step (the step's working set is larger than L2), so their loads / stores can carry the streaming (evict-first)
 * policy and leave L1 to the spilled values */
struct B4StRef {
    double *p;
    __device__ __forceinline__ operator double() const { return __ldcs(p); }
    __device__ __forceinline__ double operator=(double v) const { if (NGB_B4_STREAM > 1) __stcs(p, v); else *p = v; return v; }
    __device__ __forceinline__ double operator=(const B4StRef &o) const { const double v = o; *this = v; return v; }
};
#define B4ST(h, k) (B4StRef{ &b4st##h[(size_t)(k) * c->T] })
#else
#define B4ST(h, k) b4st##h[(size_t)(k) * c->T]
#endif

/* Phase A: terminal voltages by INITF mode and Newton step limiting (b4ld.c:257-698). */
template <unsigned VK>
NGB_HD void b4_fetch_limit(const B4Ctx *c, size_t t, int inst, int s, int head, int mode_ckt,
                           const double *Mrow, int b4dr, int flags, B4W *w)
{
    const int rbodyMod = B4SEL_RBODY(flags), rgateMod = B4SEL_RGATE(flags);
    const int off = flags & B4F_OFF;
    const double type = B4M(type);
    const int rdsMod = B4SEL(rdsMod);
    const int S = c->S;
    B4ST_BASES(head);
    double vds, vgs, vbs, vges, vgms, vdbs, vsbs, vses, vdes, qdef;
    double vbd, vgd, vged, vgmd, vdbd;
    int Check = 1, Check1 = 1, Check2 = 1;

    if (mode_ckt & NGB_MODEINITTRAN) {
        vds = B4ST(1, B4ST_vds);   vgs = B4ST(1, B4ST_vgs);   vbs = B4ST(1, B4ST_vbs);
        vges = B4ST(1, B4ST_vges); vgms = B4ST(1, B4ST_vgms); vdbs = B4ST(1, B4ST_vdbs);
        vsbs = B4ST(1, B4ST_vsbs); vses = B4ST(1, B4ST_vses); vdes = B4ST(1, B4ST_vdes);
        qdef = B4ST(1, B4ST_qdef);
    } else if ((mode_ckt & NGB_MODEINITJCT) && !off) {
        vds = type * B4I(icVDS);
        vgs = vges = vgms = type * B4I(icVGS);
        vbs = vdbs = vsbs = type * B4I(icVBS);
        if (vds > 0.0)      { vdes = vds + 0.01; vses = -0.01; }
        else if (vds < 0.0) { vdes = vds - 0.01; vses = 0.01; }
        else vdes = vses = 0.0;
        qdef = 0.0;
        if ((vds == 0.0) && (vgs == 0.0) && (vbs == 0.0) &&
            ((mode_ckt & (NGB_MODETRAN | NGB_MODEAC | NGB_MODEDCOP | NGB_MODEDCTRANCURVE)) ||
             (!(mode_ckt & NGB_MODEUIC)))) {
            vds = 0.1; vdes = 0.11; vses = -0.01;
            vgs = vges = vgms = type * B4I(vth0) + 0.1;
            vbs = vdbs = vsbs = 0.0;
        }
    } else if ((mode_ckt & (NGB_MODEINITJCT | NGB_MODEINITFIX)) && off) {
        vds = vgs = vbs = vges = vgms = 0.0;
        vdbs = vsbs = vdes = vses = qdef = 0.0;
    } else {
        double vgdo, vgedo, vgmdo, von;
        if (mode_ckt & NGB_MODEINITPRED) {
            /* linear extrapolation from the two previous time points; state0 receives the
             * state1 voltages first (b4ld.c:323-372) */
            const double xfact = NGB_LDG(&c->ctl.delta[s]) / NGB_LDG(&c->ctl.delta_old[(size_t)1 * S + s]);
            double s1, s2;
#define B4_PRED(K, V) s1 = B4ST(1, K); s2 = B4ST(2, K); B4ST(0, K) = s1; V = (1.0 + xfact) * s1 - (xfact * s2);
            B4_PRED(B4ST_vds, vds)
            B4_PRED(B4ST_vgs, vgs)
            B4_PRED(B4ST_vges, vges)
            B4_PRED(B4ST_vgms, vgms)
            B4_PRED(B4ST_vbs, vbs)
            B4ST(0, B4ST_vbd) = B4ST(0, B4ST_vbs) - B4ST(0, B4ST_vds);
            B4_PRED(B4ST_vdbs, vdbs)
            B4ST(0, B4ST_vdbd) = B4ST(0, B4ST_vdbs) - B4ST(0, B4ST_vds);
            B4_PRED(B4ST_vsbs, vsbs)
            B4_PRED(B4ST_vses, vses)
            B4_PRED(B4ST_vdes, vdes)
            B4_PRED(B4ST_qdef, qdef)
#undef B4_PRED
        } else {
            /* gather the previous Newton iterate: x[node][sample] */
            const double *xo = c->x + (size_t)NGB_LDG(&c->ctl.xsel[s]) * c->neq1 * S;
#define B4_X(role) NGB_LDG(&xo[(size_t)NGB_LDG(&c->nodes[B4N_##role * c->ninst + inst]) * S + s])
            const double xsp = B4_X(sNodePrime);
            vds  = type * (B4_X(dNodePrime) - xsp);
            vgs  = type * (B4_X(gNodePrime) - xsp);
            vbs  = type * (B4_X(bNodePrime) - xsp);
            vges = type * (B4_X(gNodeExt) - xsp);
            vgms = type * (B4_X(gNodeMid) - xsp);
            vdbs = type * (B4_X(dbNode) - xsp);
            vsbs = type * (B4_X(sbNode) - xsp);
            vses = type * (B4_X(sNode) - xsp);
            vdes = type * (B4_X(dNode) - xsp);
            qdef = type * (B4_X(qNode));
#undef B4_X
        }

        {
            const double o_vds = B4ST(0, B4ST_vds), o_vgs = B4ST(0, B4ST_vgs);
            const double o_vges = B4ST(0, B4ST_vges), o_vgms = B4ST(0, B4ST_vgms);
            vgdo = o_vgs - o_vds;
            vgedo = o_vges - o_vds;
            vgmdo = o_vgms - o_vds;

            vbd = vbs - vds;
            vdbd = vdbs - vds;
            vgd = vgs - vds;
            vged = vges - vds;
            vgmd = vgms - vds;

            /* cdhat/cbhat/... of b4ld.c:442-506 feed only the bypass test and the
             * pre-NEWCONV convergence check; neither exists in this build. */

            von = NGB_LDG(&c->op[(size_t)B4O_von * c->T + t]);
            if (o_vds >= 0.0) {
                vgs = ngb_fetlim(vgs, o_vgs, von);
                vds = vgs - vgd;
                vds = ngb_limvds(vds, o_vds);
                vgd = vgs - vds;
                if (rgateMod == 3) {
                    vges = ngb_fetlim(vges, o_vges, von);
                    vgms = ngb_fetlim(vgms, o_vgms, von);
                    vged = vges - vds;
                    vgmd = vgms - vds;
                } else if ((rgateMod == 1) || (rgateMod == 2)) {
                    vges = ngb_fetlim(vges, o_vges, von);
                    vged = vges - vds;
                }
                if (rdsMod) {
                    vdes = ngb_limvds(vdes, B4ST(0, B4ST_vdes));
                    vses = -ngb_limvds(-vses, -(B4ST(0, B4ST_vses)));
                }
            } else {
                vgd = ngb_fetlim(vgd, vgdo, von);
                vds = vgs - vgd;
                vds = -ngb_limvds(-vds, -o_vds);
                vgs = vgd + vds;
                if (rgateMod == 3) {
                    vged = ngb_fetlim(vged, vgedo, von);
                    vges = vged + vds;
                    vgmd = ngb_fetlim(vgmd, vgmdo, von);
                    vgms = vgmd + vds;
                }
                if ((rgateMod == 1) || (rgateMod == 2)) {
                    vged = ngb_fetlim(vged, vgedo, von);
                    vges = vged + vds;
                }
                if (rdsMod) {
                    vdes = -ngb_limvds(-vdes, -(B4ST(0, B4ST_vdes)));
                    vses = ngb_limvds(vses, B4ST(0, B4ST_vses));
                }
            }

            {
                const double vcrit = B4M(vcrit), vt0 = c->vt0;
                if (vds >= 0.0) {
                    vbs = ngb_pnjlim(vbs, B4ST(0, B4ST_vbs), vt0, vcrit, &Check);
                    vbd = vbs - vds;
                    if (rbodyMod) {
                        vdbs = ngb_pnjlim(vdbs, B4ST(0, B4ST_vdbs), vt0, vcrit, &Check1);
                        vdbd = vdbs - vds;
                        vsbs = ngb_pnjlim(vsbs, B4ST(0, B4ST_vsbs), vt0, vcrit, &Check2);
                        Check = ((Check1 == 0) && (Check2 == 0)) ? 0 : 1;
                    }
                } else {
                    vbd = ngb_pnjlim(vbd, B4ST(0, B4ST_vbd), vt0, vcrit, &Check);
                    vbs = vbd + vds;
                    if (rbodyMod) {
                        double vsbdo, vsbd;
                        vdbd = ngb_pnjlim(vdbd, B4ST(0, B4ST_vdbd), vt0, vcrit, &Check1);
                        vdbs = vdbd + vds;
                        vsbdo = B4ST(0, B4ST_vsbs) - B4ST(0, B4ST_vds);
                        vsbd = vsbs - vds;
                        vsbd = ngb_pnjlim(vsbd, vsbdo, vt0, vcrit, &Check2);
                        vsbs = vsbd + vds;
                        Check = ((Check1 == 0) && (Check2 == 0)) ? 0 : 1;
                    }
                }
            }
        }
    }

    w->vds = vds; w->vgs = vgs; w->vbs = vbs; w->vges = vges; w->vgms = vgms;
    w->vdbs = vdbs; w->vsbs = vsbs; w->vses = vses; w->vdes = vdes; w->qdef = qdef;
    vbd = vbs - vds;
    vdbd = vdbs - vds;
    w->vbd = vbd;
    w->vgd = vgs - vds;
    w->vgb = vgs - vbs;
    w->vged = vges - vds;
    w->vgmd = vgms - vds;
    w->vgmb = vgms - vbs;
    w->vdbd = vdbd;
    w->vbs_jct = (!rbodyMod) ? vbs : vsbs;
    w->vbd_jct = (!rbodyMod) ? vbd : vdbd;
    w->Check = Check;
}

/* Phase B+C: junction diodes, threshold voltage, effective gate drive, mobility, Vdsat,
 * drain current and its output-resistance corrections (b4ld.c:700-2189). */
template <unsigned VK>
NGB_HD void b4_core_dc(const B4Ctx *c, size_t t, int s, const double *Mrow, const double *Prow, int b4dr,
                       int flags, B4W *w)
{
    const double gmin = NGB_LDG(&c->ctl.gmin[s]);
    const double nf = B4I(nf);
    const double vtm = B4M(vtm), vtm0 = B4M(vtm0);
    const int mtrlMod = B4SEL(mtrlMod);
    const double coxe = B4M(coxe);
    double T0, T1, T2, T3, T4, T5, T6, T7, T8, T9, T10, T11, T12, T13, T14;
    double dT0_dVg, dT0_dVd, dT0_dVb, dT1_dVg, dT1_dVd, dT1_dVb, dT2_dVg, dT2_dVd, dT2_dVb;
    double dT3_dVg, dT3_dVd, dT3_dVb, dT4_dVd, dT5_dVg, dT5_dVd, dT5_dVb;
    double dT6_dVg, dT6_dVd, dT6_dVb, dT8_dVg, dT8_dVd, dT8_dVb, dT9_dVg, dT9_dVd, dT9_dVb;
    double dT10_dVg, dT10_dVd, dT10_dVb;
    double tmp, tmp1, tmp2, tmp3, tmp4;

    NGB_CTA_ALIGN();
    /* ---- source/drain junction diodes (DC) ---- */
    {
        const int dioMod = B4SEL(dioMod);
        const double weffCJnf = B4P(weffCJ) * nf;
        double Isat, Nvtm, jg, jc;

        Nvtm = vtm * B4M(SjctEmissionCoeff);
        if ((B4I(Aseff) <= 0.0) && (B4I(Pseff) <= 0.0)) Isat = 0.0;
        else Isat = B4I(Aseff) * B4M(SjctTempSatCurDensity)
                  + B4I(Pseff) * B4M(SjctSidewallTempSatCurDensity)
                  + weffCJnf * B4M(SjctGateSidewallTempSatCurDensity);
        b4_junction_dc(dioMod, Isat, Nvtm, w->vbs_jct, gmin, B4M(bvs), B4M(xjbvs), B4I(XExpBVS),
                       B4I(vjsmFwd), B4I(vjsmRev), B4I(IVjsmFwd), B4I(IVjsmRev),
                       B4I(SslpFwd), B4I(SslpRev), &jg, &jc);
        double gbs = jg, cbs = jc;

        Nvtm = vtm * B4M(DjctEmissionCoeff);
        if ((B4I(Adeff) <= 0.0) && (B4I(Pdeff) <= 0.0)) Isat = 0.0;
        else Isat = B4I(Adeff) * B4M(DjctTempSatCurDensity)
                  + B4I(Pdeff) * B4M(DjctSidewallTempSatCurDensity)
                  + weffCJnf * B4M(DjctGateSidewallTempSatCurDensity);
        b4_junction_dc(dioMod, Isat, Nvtm, w->vbd_jct, gmin, B4M(bvd), B4M(xjbvd), B4I(XExpBVD),
                       B4I(vjdmFwd), B4I(vjdmRev), B4I(IVjdmFwd), B4I(IVjdmRev),
                       B4I(DslpFwd), B4I(DslpRev), &jg, &jc);
        double gbd = jg, cbd = jc;

        /* trap-assisted tunnelling and recombination current for reverse bias */
        {
            double Ts1, Ts3, Ts5, Td2, Td4, Td6, dTs1, dTs3, dTs5, dTd2, dTd4, dTd6;
            b4_tat(B4M(vtss),    w->vbs_jct, vtm0 * B4M(njtsstemp),    &Ts1, &dTs1);
            b4_tat(B4M(vtsd),    w->vbd_jct, vtm0 * B4M(njtsdtemp),    &Td2, &dTd2);
            b4_tat(B4M(vtssws),  w->vbs_jct, vtm0 * B4M(njtsswstemp),  &Ts3, &dTs3);
            b4_tat(B4M(vtsswd),  w->vbd_jct, vtm0 * B4M(njtsswdtemp),  &Td4, &dTd4);
            b4_tat(B4M(vtsswgs), w->vbs_jct, vtm0 * B4M(njtsswgstemp), &Ts5, &dTs5);
            b4_tat(B4M(vtsswgd), w->vbd_jct, vtm0 * B4M(njtsswgdtemp), &Td6, &dTd6);
            gbs += B4I(SjctTempRevSatCur) * dTs1 + B4I(SswTempRevSatCur) * dTs3
                    + B4I(SswgTempRevSatCur) * dTs5;
            cbs -= B4I(SjctTempRevSatCur) * (Ts1 - 1.0) + B4I(SswTempRevSatCur) * (Ts3 - 1.0)
                    + B4I(SswgTempRevSatCur) * (Ts5 - 1.0);
            gbd += B4I(DjctTempRevSatCur) * dTd2 + B4I(DswTempRevSatCur) * dTd4
                    + B4I(DswgTempRevSatCur) * dTd6;
            cbd -= B4I(DjctTempRevSatCur) * (Td2 - 1.0) + B4I(DswTempRevSatCur) * (Td4 - 1.0)
                    + B4I(DswgTempRevSatCur) * (Td6 - 1.0);
        }
        w->gbs = gbs; w->cbs = cbs; w->gbd = gbd; w->cbd = cbd;
    }

    NGB_CTA_ALIGN();
    /* ---- orientation ---- */
    double Vds, Vgs, Vbs, Vdb;
    if (w->vds >= 0.0) { w->mode = 1;  Vds = w->vds;  Vgs = w->vgs; Vbs = w->vbs; Vdb = w->vds - w->vbs; }
    else               { w->mode = -1; Vds = -w->vds; Vgs = w->vgd; Vbs = w->vbd; Vdb = -w->vbs; }
    w->Vds = Vds; w->Vgs = Vgs; w->Vbs = Vbs; w->Vdb = Vdb;

    double epsrox, toxe, epssub;
    if (mtrlMod) { epsrox = 3.9; toxe = B4M(eot); epssub = B4_EPS0 * B4M(epsrsub); }
    else { epsrox = B4M(epsrox); toxe = B4M(toxe); epssub = B4_EPSSI; }
    w->toxe = toxe; w->epsrox = epsrox; w->epssub = epssub;

    NGB_CTA_ALIGN();
    /* ---- effective body bias ---- */
    const double vbsc = B4I(vbsc);
    const double phi = B4P(phi), sqrtPhi = B4P(sqrtPhi);
    double Vbseff, dVbseff_dVb;
    T0 = Vbs - vbsc - 0.001;
    T1 = sqrt(T0 * T0 - 0.004 * vbsc);
    if (T0 >= 0.0) {
        Vbseff = vbsc + 0.5 * (T0 + T1);
        dVbseff_dVb = 0.5 * (1.0 + T0 / T1);
    } else {
        T2 = -0.002 / (T1 - T0);
        Vbseff = vbsc * (1.0 + T2);
        dVbseff_dVb = T2 * vbsc / T1;
    }
    /* correction to forward body bias */
    T9 = 0.95 * phi;
    T0 = T9 - Vbseff - 0.001;
    T1 = sqrt(T0 * T0 + 0.004 * T9);
    Vbseff = T9 - 0.5 * (T0 + T1);
    dVbseff_dVb *= 0.5 * (1.0 + T0 / T1);
    const double Phis = phi - Vbseff;
    const double dPhis_dVb = -1.0;
    const double sqrtPhis = sqrt(Phis);
    const double dsqrtPhis_dVb = -0.5 / sqrtPhis;

    const double Xdep = B4P(Xdep0) * sqrtPhis / sqrtPhi;
    const double dXdep_dVb = (B4P(Xdep0) / sqrtPhi) * dsqrtPhis_dVb;

    const double Leff = B4P(leff);
    const double Vtm = vtm, Vtm0 = vtm0;
    const double weff = B4P(weff);
    const double factor1 = B4M(factor1);

    NGB_CTA_ALIGN();
    /* ---- threshold voltage ---- */
    T3 = sqrt(Xdep);
    const double V0 = B4P(vbi) - phi;

    T0 = B4P(dvt2) * Vbseff;
    if (T0 >= -0.5) { T1 = 1.0 + T0; T2 = B4P(dvt2); }
    else { T4 = 1.0 / (3.0 + 8.0 * T0); T1 = (1.0 + 3.0 * T0) * T4; T2 = B4P(dvt2) * T4 * T4; }
    const double lt1 = factor1 * T3 * T1;
    const double dlt1_dVb = factor1 * (0.5 / T3 * T1 * dXdep_dVb + T3 * T2);

    T0 = B4P(dvt2w) * Vbseff;
    if (T0 >= -0.5) { T1 = 1.0 + T0; T2 = B4P(dvt2w); }
    else { T4 = 1.0 / (3.0 + 8.0 * T0); T1 = (1.0 + 3.0 * T0) * T4; T2 = B4P(dvt2w) * T4 * T4; }
    const double ltw = factor1 * T3 * T1;
    const double dltw_dVb = factor1 * (0.5 / T3 * T1 * dXdep_dVb + T3 * T2);

    double Theta0, dTheta0_dVb;
    T0 = B4P(dvt1) * Leff / lt1;
    if (T0 < B4_EXP_THRESHOLD) {
        T1 = ngb_exp(T0);
        T2 = T1 - 1.0;
        T3 = T2 * T2;
        T4 = T3 + 2.0 * T1 * B4_MIN_EXP;
        Theta0 = T1 / T4;
        dT1_dVb = -T0 * T1 * dlt1_dVb / lt1;
        dTheta0_dVb = dT1_dVb * (T4 - 2.0 * T1 * (T2 + B4_MIN_EXP)) / T4 / T4;
    } else {
        Theta0 = 1.0 / (B4_MAX_EXP - 2.0);
        dTheta0_dVb = 0.0;
    }
    const double thetavth = B4P(dvt0) * Theta0;
    const double Delt_vth = thetavth * V0;
    const double dDelt_vth_dVb = B4P(dvt0) * dTheta0_dVb * V0;

    T0 = B4P(dvt1w) * weff * Leff / ltw;
    if (T0 < B4_EXP_THRESHOLD) {
        T1 = ngb_exp(T0);
        T2 = T1 - 1.0;
        T3 = T2 * T2;
        T4 = T3 + 2.0 * T1 * B4_MIN_EXP;
        T5 = T1 / T4;
        dT1_dVb = -T0 * T1 * dltw_dVb / ltw;
        dT5_dVb = dT1_dVb * (T4 - 2.0 * T1 * (T2 + B4_MIN_EXP)) / T4 / T4;
    } else {
        T5 = 1.0 / (B4_MAX_EXP - 2.0);
        dT5_dVb = 0.0;
    }
    T0 = B4P(dvt0w) * T5;
    T2 = T0 * V0;
    dT2_dVb = B4P(dvt0w) * dT5_dVb * V0;

    const double TempRatio = c->temp / B4M(tnom) - 1.0;
    const double k1ox = B4P(k1ox);
    T0 = sqrt(1.0 + B4P(lpe0) / Leff);
    T1 = k1ox * (T0 - 1.0) * sqrtPhi
       + (B4P(kt1) + B4P(kt1l) / Leff + B4P(kt2) * Vbseff) * TempRatio;
    const double Vth_NarrowW = toxe * phi / (weff + B4P(w0));

    T3 = B4I(eta0) + B4P(etab) * Vbseff;
    if (T3 < 1.0e-4) {
        T9 = 1.0 / (3.0 - 2.0e4 * T3);
        T3 = (2.0e-4 - T3) * T9;
        T4 = T9 * T9;
    } else {
        T4 = 1.0;
    }
    const double dDIBL_Sft_dVd = T3 * B4P(theta0vb0);
    const double DIBL_Sft = dDIBL_Sft_dVd * Vds;

    const double Lpe_Vb = sqrt(1.0 + B4P(lpeb) / Leff);
    const double type = B4M(type);
    const double vth0 = B4I(vth0), k2ox = B4I(k2ox);

    double Vth = type * vth0 + (k1ox * sqrtPhis - B4P(k1) * sqrtPhi) * Lpe_Vb
               - k2ox * Vbseff - Delt_vth - T2 + (B4P(k3) + B4P(k3b) * Vbseff) * Vth_NarrowW
               + T1 - DIBL_Sft;
    double dVth_dVb = Lpe_Vb * k1ox * dsqrtPhis_dVb - k2ox - dDelt_vth_dVb - dT2_dVb
                    + B4P(k3b) * Vth_NarrowW - B4P(etab) * Vds * B4P(theta0vb0) * T4
                    + B4P(kt2) * TempRatio;
    double dVth_dVd = -dDIBL_Sft_dVd;

    NGB_CTA_ALIGN();
    /* ---- subthreshold swing factor n ---- */
    double n, dn_dVb, dn_dVd;
    tmp1 = epssub / Xdep;
    /* nstar (noise only) is not needed on this path */
    tmp2 = B4P(nfactor) * tmp1;
    tmp3 = B4P(cdsc) + B4P(cdscb) * Vbseff + B4P(cdscd) * Vds;
    tmp4 = (tmp2 + tmp3 * Theta0 + B4P(cit)) / coxe;
    if (tmp4 >= -0.5) {
        n = 1.0 + tmp4;
        dn_dVb = (-tmp2 / Xdep * dXdep_dVb + tmp3 * dTheta0_dVb + B4P(cdscb) * Theta0) / coxe;
        dn_dVd = B4P(cdscd) * Theta0 / coxe;
    } else {
        T0 = 1.0 / (3.0 + 8.0 * tmp4);
        n = (1.0 + 3.0 * tmp4) * T0;
        T0 *= T0;
        dn_dVb = (-tmp2 / Xdep * dXdep_dVb + tmp3 * dTheta0_dVb + B4P(cdscb) * Theta0) / coxe * T0;
        dn_dVd = B4P(cdscd) * Theta0 / coxe * T0;
    }

    /* Vth correction for pocket implant */
    const int tempMod = B4SEL(tempMod);
    if (B4P(dvtp0) > 0.0) {
        double dDITS_Sft_dVd, dDITS_Sft_dVb;
        T0 = -B4P(dvtp1) * Vds;
        if (T0 < -B4_EXP_THRESHOLD) { T2 = B4_MIN_EXP; dT2_dVd = 0.0; }
        else { T2 = ngb_exp(T0); dT2_dVd = -B4P(dvtp1) * T2; }
        T3 = Leff + B4P(dvtp0) * (1.0 + T2);
        dT3_dVd = B4P(dvtp0) * dT2_dVd;
        if (tempMod < 2) {
            T4 = Vtm * ngb_log(Leff / T3);
            dT4_dVd = -Vtm * dT3_dVd / T3;
        } else {
            T4 = vtm0 * ngb_log(Leff / T3);
            dT4_dVd = -vtm0 * dT3_dVd / T3;
        }
        dDITS_Sft_dVd = dn_dVd * T4 + n * dT4_dVd;
        dDITS_Sft_dVb = T4 * dn_dVb;
        Vth -= n * T4;
        dVth_dVd -= dDITS_Sft_dVd;
        dVth_dVb -= dDITS_Sft_dVb;
    }

    /* v4.7 DITS_SFT2 */
    if (!((B4P(dvtp4) == 0.0) || (B4P(dvtp2factor) == 0.0))) {
        double DITS_Sft2, dDITS_Sft2_dVd;
        T1 = 2.0 * B4P(dvtp4) * Vds;
        b4_dexp(T1, &T0, &T10);
        DITS_Sft2 = B4P(dvtp2factor) * (T0 - 1) / (T0 + 1);
        dDITS_Sft2_dVd = B4P(dvtp2factor) * B4P(dvtp4) * 4.0 * T10 / ((T0 + 1) * (T0 + 1));
        Vth -= DITS_Sft2;
        dVth_dVd -= dDITS_Sft2_dVd;
    }

    w->von = Vth;

    NGB_CTA_ALIGN();
    /* ---- poly-gate depletion ---- */
    {
        const double vfbphi = B4I(vfb) + phi;
        const double epsg = (mtrlMod == 0) ? B4_EPSSI : B4M(epsrgate) * B4_EPS0;
        const double ngate = B4P(ngate);
        b4_poly_depletion(vfbphi, ngate, epsg, coxe, w->vgs, &w->vgs_eff, &w->dvgs_eff_dvg);
        b4_poly_depletion(vfbphi, ngate, epsg, coxe, w->vgd, &w->vgd_eff, &w->dvgd_eff_dvg);
    }
    double Vgs_eff, dVgs_eff_dVg;
    if (w->mode > 0) { Vgs_eff = w->vgs_eff; dVgs_eff_dVg = w->dvgs_eff_dvg; }
    else             { Vgs_eff = w->vgd_eff; dVgs_eff_dVg = w->dvgd_eff_dvg; }

    const double Vgst = Vgs_eff - Vth;

    NGB_CTA_ALIGN();
    /* ---- Vgsteff ---- */
    const double mstar = B4P(mstar);
    const double cdep0 = B4P(cdep0);
    double ExpVgst;
    T0 = n * Vtm;
    T1 = mstar * Vgst;
    T2 = T1 / T0;
    if (T2 > B4_EXP_THRESHOLD) {
        T10 = T1;
        dT10_dVg = mstar * dVgs_eff_dVg;
        dT10_dVd = -dVth_dVd * mstar;
        dT10_dVb = -dVth_dVb * mstar;
    } else if (T2 < -B4_EXP_THRESHOLD) {
        T10 = Vtm * ngb_log(1.0 + B4_MIN_EXP);
        dT10_dVg = 0.0;
        dT10_dVd = T10 * dn_dVd;
        dT10_dVb = T10 * dn_dVb;
        T10 *= n;
    } else {
        ExpVgst = ngb_exp(T2);
        T3 = Vtm * ngb_log(1.0 + ExpVgst);
        T10 = n * T3;
        dT10_dVg = mstar * ExpVgst / (1.0 + ExpVgst);
        dT10_dVb = T3 * dn_dVb - dT10_dVg * (dVth_dVb + Vgst * dn_dVb / n);
        dT10_dVd = T3 * dn_dVd - dT10_dVg * (dVth_dVd + Vgst * dn_dVd / n);
        dT10_dVg *= dVgs_eff_dVg;
    }

    T1 = B4P(voffcbn) - (1.0 - mstar) * Vgst;
    T2 = T1 / T0;
    if (T2 < -B4_EXP_THRESHOLD) {
        T3 = coxe * B4_MIN_EXP / cdep0;
        T9 = mstar + T3 * n;
        dT9_dVg = 0.0;
        dT9_dVd = dn_dVd * T3;
        dT9_dVb = dn_dVb * T3;
    } else if (T2 > B4_EXP_THRESHOLD) {
        T3 = coxe * B4_MAX_EXP / cdep0;
        T9 = mstar + T3 * n;
        dT9_dVg = 0.0;
        dT9_dVd = dn_dVd * T3;
        dT9_dVb = dn_dVb * T3;
    } else {
        ExpVgst = ngb_exp(T2);
        T3 = coxe / cdep0;
        T4 = T3 * ExpVgst;
        T5 = T1 * T4 / T0;
        T9 = mstar + n * T4;
        dT9_dVg = T3 * (mstar - 1.0) * ExpVgst / Vtm;
        dT9_dVb = T4 * dn_dVb - dT9_dVg * dVth_dVb - T5 * dn_dVb;
        dT9_dVd = T4 * dn_dVd - dT9_dVg * dVth_dVd - T5 * dn_dVd;
        dT9_dVg *= dVgs_eff_dVg;
    }
    const double Vgsteff = T10 / T9;
    T11 = T9 * T9;
    const double dVgsteff_dVg = (T9 * dT10_dVg - T10 * dT9_dVg) / T11;
    const double dVgsteff_dVd = (T9 * dT10_dVd - T10 * dT9_dVd) / T11;
    const double dVgsteff_dVb = (T9 * dT10_dVb - T10 * dT9_dVb) / T11;

    NGB_CTA_ALIGN();
    /* ---- effective channel geometry ---- */
    T9 = sqrtPhis - sqrtPhi;
    double Weff = weff - 2.0 * (B4P(dwg) * Vgsteff + B4P(dwb) * T9);
    double dWeff_dVg = -2.0 * B4P(dwg);
    double dWeff_dVb = -2.0 * B4P(dwb) * dsqrtPhis_dVb;
    if (Weff < 2.0e-8) {
        T0 = 1.0 / (6.0e-8 - 2.0 * Weff);
        Weff = 2.0e-8 * (4.0e-8 - Weff) * T0;
        T0 *= T0 * 4.0e-16;
        dWeff_dVg *= T0;
        dWeff_dVb *= T0;
    }

    const int rdsMod = B4SEL(rdsMod);
    double Rds, dRds_dVg, dRds_dVb;
    if (rdsMod == 1) {
        Rds = dRds_dVg = dRds_dVb = 0.0;
    } else {
        T0 = 1.0 + B4P(prwg) * Vgsteff;
        dT0_dVg = -B4P(prwg) / T0 / T0;
        T1 = B4P(prwb) * T9;
        dT1_dVb = B4P(prwb) * dsqrtPhis_dVb;
        T2 = 1.0 / T0 + T1;
        T3 = T2 + sqrt(T2 * T2 + 0.01);
        dT3_dVg = 1.0 + T2 / (T3 - T2);
        dT3_dVb = dT3_dVg * dT1_dVb;
        dT3_dVg *= dT0_dVg;
        T4 = B4P(rds0) * 0.5;
        Rds = B4P(rdswmin) + T3 * T4;
        dRds_dVg = T4 * dT3_dVg;
        dRds_dVb = T4 * dT3_dVb;
        /* grdsw (noise/ask only) not kept */
    }

    NGB_CTA_ALIGN();
    /* ---- Abulk ---- */
    double Abulk0, dAbulk0_dVb, Abulk, dAbulk_dVg, dAbulk_dVb, Abulk0_Q, dAbulk0_Q_dVb;
    T9 = 0.5 * k1ox * Lpe_Vb / sqrtPhis;
    T1 = T9 + k2ox - B4P(k3b) * Vth_NarrowW;
    dT1_dVb = -T9 / sqrtPhis * dsqrtPhis_dVb;

    T9 = sqrt(B4P(xj) * Xdep);
    tmp1 = Leff + 2.0 * T9;
    T5 = Leff / tmp1;
    tmp2 = B4P(a0) * T5;
    tmp3 = weff + B4P(b1);
    tmp4 = B4P(b0) / tmp3;
    T2 = tmp2 + tmp4;
    dT2_dVb = -T9 / tmp1 / Xdep * dXdep_dVb;
    T6 = T5 * T5;
    T7 = T5 * T6;

    Abulk0 = 1.0 + T1 * T2;
    dAbulk0_dVb = T1 * tmp2 * dT2_dVb + T2 * dT1_dVb;

    T8 = B4P(ags) * B4P(a0) * T7;
    dAbulk_dVg = -T1 * T8;
    Abulk = Abulk0 + dAbulk_dVg * Vgsteff;
    dAbulk_dVb = dAbulk0_dVb - T8 * Vgsteff * (dT1_dVb + 3.0 * T1 * dT2_dVb);

    if (Abulk0 < 0.1) {
        T9 = 1.0 / (3.0 - 20.0 * Abulk0);
        Abulk0 = (0.2 - Abulk0) * T9;
        dAbulk0_dVb *= T9 * T9;
    }
    if (Abulk < 0.1) {
        T9 = 1.0 / (3.0 - 20.0 * Abulk);
        Abulk = (0.2 - Abulk) * T9;
        T10 = T9 * T9;
        dAbulk_dVb *= T10;
        dAbulk_dVg *= T10;
    }

    const double keta = B4P(keta), ketac = B4P(ketac);
    T2 = keta * Vbseff;
    if (T2 >= -0.9) {
        T0 = 1.0 / (1.0 + T2);
        dT0_dVb = -keta * T0 * T0;
    } else {
        T1 = 1.0 / (0.8 + T2);
        T0 = (17.0 + 20.0 * T2) * T1;
        dT0_dVb = -keta * T1 * T1;
    }
    dAbulk_dVg *= T0;
    dAbulk_dVb = dAbulk_dVb * T0 + Abulk * dT0_dVb;
    dAbulk0_Q_dVb = dAbulk0_dVb;
    dAbulk0_dVb = dAbulk0_dVb * T0 + Abulk0 * dT0_dVb;
    Abulk *= T0;
    Abulk0_Q = Abulk0;
    Abulk0 *= T0;

    if (ketac != keta) {
        T2 = ketac * Vbseff;
        if (T2 >= -0.9) {
            T0 = 1.0 / (1.0 + T2);
            dT0_dVb = -ketac * T0 * T0;
        } else {
            T1 = 1.0 / (0.8 + T2);
            T0 = (17.0 + 20.0 * T2) * T1;
            dT0_dVb = -ketac * T1 * T1;
        }
        dAbulk0_Q_dVb = dAbulk0_Q_dVb * T0 + Abulk0_Q * dT0_dVb;
        Abulk0_Q *= T0;
    } else {
        dAbulk0_Q_dVb = dAbulk0_dVb;
        Abulk0_Q = Abulk0;
    }

    NGB_CTA_ALIGN();
    /* ---- mobility ---- */
    const int mobMod = B4SEL(mobMod);
    const double ua = B4P(ua), ub = B4P(ub), uc = B4P(uc), ud = B4P(ud);
    double dDenomi_dVg, dDenomi_dVd, dDenomi_dVb, Denomi;
    if (mtrlMod && (B4SEL(mtrlCompatMod) == 0))
        T14 = 2.0 * type * (B4M(phig) - B4M(easub) - 0.5 * B4M(Eg0) + 0.45);
    else
        T14 = 0.0;

    if (mobMod == 0) {
        T0 = Vgsteff + Vth + Vth - T14;
        T2 = ua + uc * Vbseff;
        T3 = T0 / toxe;
        T12 = sqrt(Vth * Vth + 0.0001);
        T9 = 1.0 / (Vgsteff + 2 * T12);
        T10 = T9 * toxe;
        T8 = ud * T10 * T10 * Vth;
        T6 = T8 * Vth;
        T5 = T3 * (T2 + ub * T3) + T6;
        T7 = -2.0 * T6 * T9;
        T11 = T7 * Vth / T12;
        dDenomi_dVg = (T2 + 2.0 * ub * T3) / toxe;
        T13 = 2.0 * (dDenomi_dVg + T11 + T8);
        dDenomi_dVd = T13 * dVth_dVd;
        dDenomi_dVb = T13 * dVth_dVb + uc * T3;
        dDenomi_dVg += T7;
    } else if (mobMod == 1) {
        T0 = Vgsteff + Vth + Vth - T14;
        T2 = 1.0 + uc * Vbseff;
        T3 = T0 / toxe;
        T4 = T3 * (ua + ub * T3);
        T12 = sqrt(Vth * Vth + 0.0001);
        T9 = 1.0 / (Vgsteff + 2 * T12);
        T10 = T9 * toxe;
        T8 = ud * T10 * T10 * Vth;
        T6 = T8 * Vth;
        T5 = T4 * T2 + T6;
        T7 = -2.0 * T6 * T9;
        T11 = T7 * Vth / T12;
        dDenomi_dVg = (ua + 2.0 * ub * T3) * T2 / toxe;
        T13 = 2.0 * (dDenomi_dVg + T11 + T8);
        dDenomi_dVd = T13 * dVth_dVd;
        dDenomi_dVb = T13 * dVth_dVb + uc * T4;
        dDenomi_dVg += T7;
    } else if (mobMod == 2) {
        T0 = (Vgsteff + B4I(vtfbphi1)) / toxe;
        T1 = ngb_exp(B4P(eu) * ngb_log(T0));
        dT1_dVg = T1 * B4P(eu) / T0 / toxe;
        T2 = ua + uc * Vbseff;
        T12 = sqrt(Vth * Vth + 0.0001);
        T9 = 1.0 / (Vgsteff + 2 * T12);
        T10 = T9 * toxe;
        T8 = ud * T10 * T10 * Vth;
        T6 = T8 * Vth;
        T5 = T1 * T2 + T6;
        T7 = -2.0 * T6 * T9;
        T11 = T7 * Vth / T12;
        dDenomi_dVg = T2 * dT1_dVg + T7;
        T13 = 2.0 * (T11 + T8);
        dDenomi_dVd = T13 * dVth_dVd;
        dDenomi_dVb = T13 * dVth_dVb + T1 * uc;
    } else if (mobMod == 4) {
        const double vtfbphi1 = B4I(vtfbphi1);
        T0 = Vgsteff + vtfbphi1 - T14;
        T2 = ua + uc * Vbseff;
        T3 = T0 / toxe;
        T12 = sqrt(vtfbphi1 * vtfbphi1 + 0.0001);
        T9 = 1.0 / (Vgsteff + 2 * T12);
        T10 = T9 * toxe;
        T8 = ud * T10 * T10 * vtfbphi1;
        T6 = T8 * vtfbphi1;
        T5 = T3 * (T2 + ub * T3) + T6;
        T7 = -2.0 * T6 * T9;
        dDenomi_dVg = (T2 + 2.0 * ub * T3) / toxe;
        dDenomi_dVd = 0.0;
        dDenomi_dVb = uc * T3;
        dDenomi_dVg += T7;
    } else if (mobMod == 5) {
        const double vtfbphi1 = B4I(vtfbphi1);
        T0 = Vgsteff + vtfbphi1 - T14;
        T2 = 1.0 + uc * Vbseff;
        T3 = T0 / toxe;
        T4 = T3 * (ua + ub * T3);
        T12 = sqrt(vtfbphi1 * vtfbphi1 + 0.0001);
        T9 = 1.0 / (Vgsteff + 2 * T12);
        T10 = T9 * toxe;
        T8 = ud * T10 * T10 * vtfbphi1;
        T6 = T8 * vtfbphi1;
        T5 = T4 * T2 + T6;
        T7 = -2.0 * T6 * T9;
        dDenomi_dVg = (ua + 2.0 * ub * T3) * T2 / toxe;
        dDenomi_dVd = 0.0;
        dDenomi_dVb = uc * T4;
        dDenomi_dVg += T7;
    } else if (mobMod == 6) {
        const double vtfbphi1 = B4I(vtfbphi1);
        T0 = (Vgsteff + vtfbphi1) / toxe;
        T1 = ngb_exp(B4P(eu) * ngb_log(T0));
        dT1_dVg = T1 * B4P(eu) / T0 / toxe;
        T2 = ua + uc * Vbseff;
        T12 = sqrt(vtfbphi1 * vtfbphi1 + 0.0001);
        T9 = 1.0 / (Vgsteff + 2 * T12);
        T10 = T9 * toxe;
        T8 = ud * T10 * T10 * vtfbphi1;
        T6 = T8 * vtfbphi1;
        T5 = T1 * T2 + T6;
        T7 = -2.0 * T6 * T9;
        dDenomi_dVg = T2 * dT1_dVg + T7;
        dDenomi_dVd = 0;
        dDenomi_dVb = T1 * uc;
    } else {
        /* high-K mobility: universal + Coulombic */
        const double VgsteffVth = B4P(VgsteffVth);
        double dT11_dVg;
        T0 = (Vgsteff + B4I(vtfbphi1)) * 1.0e-8 / toxe / 6.0;
        T1 = ngb_exp(B4P(eu) * ngb_log(T0));
        dT1_dVg = T1 * B4P(eu) * 1.0e-8 / T0 / toxe / 6.0;
        T2 = ua + uc * Vbseff;
        T10 = ngb_exp(B4P(ucs) * ngb_log(0.5 + 0.5 * Vgsteff / VgsteffVth));
        T11 = ud / T10;
        dT11_dVg = -0.5 * B4P(ucs) * T11 / (0.5 + 0.5 * Vgsteff / VgsteffVth) / VgsteffVth;
        dDenomi_dVg = T2 * dT1_dVg + dT11_dVg;
        dDenomi_dVd = 0.0;
        dDenomi_dVb = T1 * uc;
        T5 = T1 * T2 + T11;
    }

    if (T5 >= -0.8) {
        Denomi = 1.0 + T5;
    } else {
        T9 = 1.0 / (7.0 + 10.0 * T5);
        Denomi = (0.6 + T5) * T9;
        T9 *= T9;
        dDenomi_dVg *= T9;
        dDenomi_dVd *= T9;
        dDenomi_dVb *= T9;
    }

    const double ueff = B4I(u0temp) / Denomi;
    T9 = -ueff / Denomi;
    const double dueff_dVg = T9 * dDenomi_dVg;
    const double dueff_dVd = T9 * dDenomi_dVd;
    const double dueff_dVb = T9 * dDenomi_dVb;

    NGB_CTA_ALIGN();
    /* ---- saturation drain voltage ---- */
    const double vsattemp = B4I(vsattemp);
    const double WVCox = Weff * vsattemp * coxe;
    const double WVCoxRds = WVCox * Rds;

    double Esat = 2.0 * vsattemp / ueff;
    double EsatL = Esat * Leff;
    T0 = -EsatL / ueff;
    double dEsatL_dVg = T0 * dueff_dVg;
    double dEsatL_dVd = T0 * dueff_dVd;
    double dEsatL_dVb = T0 * dueff_dVb;

    const double a1 = B4P(a1), a2 = B4P(a2);
    double Lambda, dLambda_dVg;
    if (a1 == 0.0) {
        Lambda = a2;
        dLambda_dVg = 0.0;
    } else if (a1 > 0.0) {
        T0 = 1.0 - a2;
        T1 = T0 - a1 * Vgsteff - 0.0001;
        T2 = sqrt(T1 * T1 + 0.0004 * T0);
        Lambda = a2 + T0 - 0.5 * (T1 + T2);
        dLambda_dVg = 0.5 * a1 * (1.0 + T1 / T2);
    } else {
        T1 = a2 + a1 * Vgsteff - 0.0001;
        T2 = sqrt(T1 * T1 + 0.0004 * a2);
        Lambda = 0.5 * (T1 + T2);
        dLambda_dVg = 0.5 * a1 * (1.0 + T1 / T2);
    }

    const double Vgst2Vtm = Vgsteff + 2.0 * Vtm;
    if (Rds > 0) {
        tmp2 = dRds_dVg / Rds + dWeff_dVg / Weff;
        tmp3 = dRds_dVb / Rds + dWeff_dVb / Weff;
    } else {
        tmp2 = dWeff_dVg / Weff;
        tmp3 = dWeff_dVb / Weff;
    }
    double Vdsat, dVdsat_dVg, dVdsat_dVd, dVdsat_dVb;
    if ((Rds == 0.0) && (Lambda == 1.0)) {
        T0 = 1.0 / (Abulk * EsatL + Vgst2Vtm);
        tmp1 = 0.0;
        T1 = T0 * T0;
        T2 = Vgst2Vtm * T0;
        T3 = EsatL * Vgst2Vtm;
        Vdsat = T3 * T0;

        dT0_dVg = -(Abulk * dEsatL_dVg + EsatL * dAbulk_dVg + 1.0) * T1;
        dT0_dVd = -(Abulk * dEsatL_dVd) * T1;
        dT0_dVb = -(Abulk * dEsatL_dVb + dAbulk_dVb * EsatL) * T1;

        dVdsat_dVg = T3 * dT0_dVg + T2 * dEsatL_dVg + EsatL * T0;
        dVdsat_dVd = T3 * dT0_dVd + T2 * dEsatL_dVd;
        dVdsat_dVb = T3 * dT0_dVb + T2 * dEsatL_dVb;
    } else {
        tmp1 = dLambda_dVg / (Lambda * Lambda);
        T9 = Abulk * WVCoxRds;
        T8 = Abulk * T9;
        T7 = Vgst2Vtm * T9;
        T6 = Vgst2Vtm * WVCoxRds;
        T0 = 2.0 * Abulk * (T9 - 1.0 + 1.0 / Lambda);
        dT0_dVg = 2.0 * (T8 * tmp2 - Abulk * tmp1 + (2.0 * T9 + 1.0 / Lambda - 1.0) * dAbulk_dVg);
        dT0_dVb = 2.0 * (T8 * (2.0 / Abulk * dAbulk_dVb + tmp3) + (1.0 / Lambda - 1.0) * dAbulk_dVb);
        dT0_dVd = 0.0;
        T1 = Vgst2Vtm * (2.0 / Lambda - 1.0) + Abulk * EsatL + 3.0 * T7;

        dT1_dVg = (2.0 / Lambda - 1.0) - 2.0 * Vgst2Vtm * tmp1 + Abulk * dEsatL_dVg
                + EsatL * dAbulk_dVg + 3.0 * (T9 + T7 * tmp2 + T6 * dAbulk_dVg);
        dT1_dVb = Abulk * dEsatL_dVb + EsatL * dAbulk_dVb + 3.0 * (T6 * dAbulk_dVb + T7 * tmp3);
        dT1_dVd = Abulk * dEsatL_dVd;

        T2 = Vgst2Vtm * (EsatL + 2.0 * T6);
        dT2_dVg = EsatL + Vgst2Vtm * dEsatL_dVg + T6 * (4.0 + 2.0 * Vgst2Vtm * tmp2);
        dT2_dVb = Vgst2Vtm * (dEsatL_dVb + 2.0 * T6 * tmp3);
        dT2_dVd = Vgst2Vtm * dEsatL_dVd;

        T3 = sqrt(T1 * T1 - 2.0 * T0 * T2);
        Vdsat = (T1 - T3) / T0;

        dT3_dVg = (T1 * dT1_dVg - 2.0 * (T0 * dT2_dVg + T2 * dT0_dVg)) / T3;
        dT3_dVd = (T1 * dT1_dVd - 2.0 * (T0 * dT2_dVd + T2 * dT0_dVd)) / T3;
        dT3_dVb = (T1 * dT1_dVb - 2.0 * (T0 * dT2_dVb + T2 * dT0_dVb)) / T3;

        dVdsat_dVg = (dT1_dVg - (T1 * dT1_dVg - dT0_dVg * T2 - T0 * dT2_dVg) / T3
                      - Vdsat * dT0_dVg) / T0;
        dVdsat_dVb = (dT1_dVb - (T1 * dT1_dVb - dT0_dVb * T2 - T0 * dT2_dVb) / T3
                      - Vdsat * dT0_dVb) / T0;
        dVdsat_dVd = (dT1_dVd - (T1 * dT1_dVd - T0 * dT2_dVd) / T3) / T0;
    }
    w->vdsat = Vdsat;

    NGB_CTA_ALIGN();
    /* ---- Vdseff ---- */
    const double delta = B4P(delta);
    double Vdseff, dVdseff_dVg, dVdseff_dVd, dVdseff_dVb;
    T1 = Vdsat - Vds - delta;
    dT1_dVg = dVdsat_dVg;
    dT1_dVd = dVdsat_dVd - 1.0;
    dT1_dVb = dVdsat_dVb;

    T2 = sqrt(T1 * T1 + 4.0 * delta * Vdsat);
    T0 = T1 / T2;
    T9 = 2.0 * delta;
    T3 = T9 / T2;
    dT2_dVg = T0 * dT1_dVg + T3 * dVdsat_dVg;
    dT2_dVd = T0 * dT1_dVd + T3 * dVdsat_dVd;
    dT2_dVb = T0 * dT1_dVb + T3 * dVdsat_dVb;

    if (T1 >= 0.0) {
        Vdseff = Vdsat - 0.5 * (T1 + T2);
        dVdseff_dVg = dVdsat_dVg - 0.5 * (dT1_dVg + dT2_dVg);
        dVdseff_dVd = dVdsat_dVd - 0.5 * (dT1_dVd + dT2_dVd);
        dVdseff_dVb = dVdsat_dVb - 0.5 * (dT1_dVb + dT2_dVb);
    } else {
        T4 = T9 / (T2 - T1);
        T5 = 1.0 - T4;
        T6 = Vdsat * T4 / (T2 - T1);
        Vdseff = Vdsat * T5;
        dVdseff_dVg = dVdsat_dVg * T5 + T6 * (dT2_dVg - dT1_dVg);
        dVdseff_dVd = dVdsat_dVd * T5 + T6 * (dT2_dVd - dT1_dVd);
        dVdseff_dVb = dVdsat_dVb * T5 + T6 * (dT2_dVb - dT1_dVb);
    }
    if (Vds == 0.0) {
        Vdseff = 0.0;
        dVdseff_dVg = 0.0;
        dVdseff_dVb = 0.0;
    }
    if (Vdseff > Vds) Vdseff = Vds;
    const double diffVds = Vds - Vdseff;

    NGB_CTA_ALIGN();
    /* ---- velocity overshoot ---- */
    if (((int)B4M(lambdaGiven)) && (B4M(lambda) > 0.0)) {
        T1 = Leff * ueff;
        T2 = B4P(lambda) / T1;
        T3 = -T2 / T1 * Leff;
        dT2_dVd = T3 * dueff_dVd;
        dT2_dVg = T3 * dueff_dVg;
        dT2_dVb = T3 * dueff_dVb;
        T5 = 1.0 / (Esat * B4P(litl));
        T4 = -T5 / EsatL;
        dT5_dVg = dEsatL_dVg * T4;
        dT5_dVd = dEsatL_dVd * T4;
        dT5_dVb = dEsatL_dVb * T4;
        T6 = 1.0 + diffVds * T5;
        dT6_dVg = dT5_dVg * diffVds - dVdseff_dVg * T5;
        dT6_dVd = dT5_dVd * diffVds + (1.0 - dVdseff_dVd) * T5;
        dT6_dVb = dT5_dVb * diffVds - dVdseff_dVb * T5;
        T7 = 2.0 / (T6 * T6 + 1.0);
        T8 = 1.0 - T7;
        T9 = T6 * T7 * T7;
        dT8_dVg = T9 * dT6_dVg;
        dT8_dVd = T9 * dT6_dVd;
        dT8_dVb = T9 * dT6_dVb;
        T10 = 1.0 + T2 * T8;
        dT10_dVg = dT2_dVg * T8 + T2 * dT8_dVg;
        dT10_dVd = dT2_dVd * T8 + T2 * dT8_dVd;
        dT10_dVb = dT2_dVb * T8 + T2 * dT8_dVb;
        if (T10 == 1.0) dT10_dVg = dT10_dVd = dT10_dVb = 0.0;

        dEsatL_dVg *= T10;
        dEsatL_dVg += EsatL * dT10_dVg;
        dEsatL_dVd *= T10;
        dEsatL_dVd += EsatL * dT10_dVd;
        dEsatL_dVb *= T10;
        dEsatL_dVb += EsatL * dT10_dVb;
        EsatL *= T10;
        Esat = EsatL / Leff;
    }

    NGB_CTA_ALIGN();
    /* ---- Vasat ---- */
    tmp4 = 1.0 - 0.5 * Abulk * Vdsat / Vgst2Vtm;
    T9 = WVCoxRds * Vgsteff;
    T8 = T9 / Vgst2Vtm;
    T0 = EsatL + Vdsat + 2.0 * T9 * tmp4;

    T7 = 2.0 * WVCoxRds * tmp4;
    dT0_dVg = dEsatL_dVg + dVdsat_dVg + T7 * (1.0 + tmp2 * Vgsteff)
            - T8 * (Abulk * dVdsat_dVg - Abulk * Vdsat / Vgst2Vtm + Vdsat * dAbulk_dVg);
    dT0_dVb = dEsatL_dVb + dVdsat_dVb + T7 * tmp3 * Vgsteff
            - T8 * (dAbulk_dVb * Vdsat + Abulk * dVdsat_dVb);
    dT0_dVd = dEsatL_dVd + dVdsat_dVd - T8 * Abulk * dVdsat_dVd;

    T9 = WVCoxRds * Abulk;
    T1 = 2.0 / Lambda - 1.0 + T9;
    dT1_dVg = -2.0 * tmp1 + WVCoxRds * (Abulk * tmp2 + dAbulk_dVg);
    dT1_dVb = dAbulk_dVb * WVCoxRds + T9 * tmp3;

    const double Vasat = T0 / T1;
    const double dVasat_dVg = (dT0_dVg - Vasat * dT1_dVg) / T1;
    const double dVasat_dVb = (dT0_dVb - Vasat * dT1_dVb) / T1;
    const double dVasat_dVd = dT0_dVd / T1;

    NGB_CTA_ALIGN();
    /* ---- Idl ---- */
    double Tcen, dTcen_dVg, Coxeff, dCoxeff_dVg;
    tmp1 = B4I(vtfbphi2);
    tmp2 = 2.0e8 * B4I(toxp);
    dT0_dVg = 1.0 / tmp2;
    T0 = (Vgsteff + tmp1) * dT0_dVg;

    tmp3 = ngb_exp(B4M(bdos) * 0.7 * ngb_log(T0));
    T1 = 1.0 + tmp3;
    T2 = B4M(bdos) * 0.7 * tmp3 / T0;
    Tcen = B4M(ados) * 1.9e-9 / T1;
    dTcen_dVg = -Tcen * T2 * dT0_dVg / T1;

    const double coxp = B4I(coxp);
    Coxeff = epssub * coxp / (epssub + coxp * Tcen);
    dCoxeff_dVg = -Coxeff * Coxeff * dTcen_dVg / epssub;

    const double CoxeffWovL = Coxeff * Weff / Leff;
    const double beta = ueff * CoxeffWovL;
    T3 = ueff / Leff;
    const double dbeta_dVg = CoxeffWovL * dueff_dVg + T3 * (Weff * dCoxeff_dVg + Coxeff * dWeff_dVg);
    const double dbeta_dVd = CoxeffWovL * dueff_dVd;
    const double dbeta_dVb = CoxeffWovL * dueff_dVb + T3 * Coxeff * dWeff_dVb;

    const double AbovVgst2Vtm = Abulk / Vgst2Vtm;
    T0 = 1.0 - 0.5 * Vdseff * AbovVgst2Vtm;
    dT0_dVg = -0.5 * (Abulk * dVdseff_dVg - Abulk * Vdseff / Vgst2Vtm + Vdseff * dAbulk_dVg) / Vgst2Vtm;
    dT0_dVd = -0.5 * Abulk * dVdseff_dVd / Vgst2Vtm;
    dT0_dVb = -0.5 * (Abulk * dVdseff_dVb + dAbulk_dVb * Vdseff) / Vgst2Vtm;

    const double fgche1 = Vgsteff * T0;
    const double dfgche1_dVg = Vgsteff * dT0_dVg + T0;
    const double dfgche1_dVd = Vgsteff * dT0_dVd;
    const double dfgche1_dVb = Vgsteff * dT0_dVb;

    T9 = Vdseff / EsatL;
    const double fgche2 = 1.0 + T9;
    const double dfgche2_dVg = (dVdseff_dVg - T9 * dEsatL_dVg) / EsatL;
    const double dfgche2_dVd = (dVdseff_dVd - T9 * dEsatL_dVd) / EsatL;
    const double dfgche2_dVb = (dVdseff_dVb - T9 * dEsatL_dVb) / EsatL;

    const double gche = beta * fgche1 / fgche2;
    const double dgche_dVg = (beta * dfgche1_dVg + fgche1 * dbeta_dVg - gche * dfgche2_dVg) / fgche2;
    const double dgche_dVd = (beta * dfgche1_dVd + fgche1 * dbeta_dVd - gche * dfgche2_dVd) / fgche2;
    const double dgche_dVb = (beta * dfgche1_dVb + fgche1 * dbeta_dVb - gche * dfgche2_dVb) / fgche2;

    T0 = 1.0 + gche * Rds;
    const double Idl = gche / T0;
    T1 = (1.0 - Idl * Rds) / T0;
    T2 = Idl * Idl;
    const double dIdl_dVg = T1 * dgche_dVg - T2 * dRds_dVg;
    const double dIdl_dVd = T1 * dgche_dVd;
    const double dIdl_dVb = T1 * dgche_dVb - T2 * dRds_dVb;

    NGB_CTA_ALIGN();
    /* ---- degradation factor due to pocket implant ---- */
    double FP, dFP_dVg;
    if (B4P(fprout) <= 0.0) {
        FP = 1.0;
        dFP_dVg = 0.0;
    } else {
        T9 = B4P(fprout) * sqrt(Leff) / Vgst2Vtm;
        FP = 1.0 / (1.0 + T9);
        dFP_dVg = FP * FP * T9 / Vgst2Vtm;
    }

    NGB_CTA_ALIGN();
    /* ---- VACLM ---- */
    double PvagTerm, dPvagTerm_dVg, dPvagTerm_dVb, dPvagTerm_dVd;
    T8 = B4P(pvag) / EsatL;
    T9 = T8 * Vgsteff;
    if (T9 > -0.9) {
        PvagTerm = 1.0 + T9;
        dPvagTerm_dVg = T8 * (1.0 - Vgsteff * dEsatL_dVg / EsatL);
        dPvagTerm_dVb = -T9 * dEsatL_dVb / EsatL;
        dPvagTerm_dVd = -T9 * dEsatL_dVd / EsatL;
    } else {
        T4 = 1.0 / (17.0 + 20.0 * T9);
        PvagTerm = (0.8 + T9) * T4;
        T4 *= T4;
        dPvagTerm_dVg = T8 * (1.0 - Vgsteff * dEsatL_dVg / EsatL) * T4;
        T9 *= T4 / EsatL;
        dPvagTerm_dVb = -T9 * dEsatL_dVb;
        dPvagTerm_dVd = -T9 * dEsatL_dVd;
    }

    double Cclm, dCclm_dVg, dCclm_dVd, dCclm_dVb, VACLM, dVACLM_dVg, dVACLM_dVd, dVACLM_dVb;
    if ((B4P(pclm) > B4_MIN_EXP) && (diffVds > 1.0e-10)) {
        T0 = 1.0 + Rds * Idl;
        dT0_dVg = dRds_dVg * Idl + Rds * dIdl_dVg;
        dT0_dVd = Rds * dIdl_dVd;
        dT0_dVb = dRds_dVb * Idl + Rds * dIdl_dVb;

        T2 = Vdsat / Esat;
        T1 = Leff + T2;
        dT1_dVg = (dVdsat_dVg - T2 * dEsatL_dVg / Leff) / Esat;
        dT1_dVd = (dVdsat_dVd - T2 * dEsatL_dVd / Leff) / Esat;
        dT1_dVb = (dVdsat_dVb - T2 * dEsatL_dVb / Leff) / Esat;

        Cclm = FP * PvagTerm * T0 * T1 / (B4P(pclm) * B4P(litl));
        dCclm_dVg = Cclm * (dFP_dVg / FP + dPvagTerm_dVg / PvagTerm + dT0_dVg / T0 + dT1_dVg / T1);
        dCclm_dVb = Cclm * (dPvagTerm_dVb / PvagTerm + dT0_dVb / T0 + dT1_dVb / T1);
        dCclm_dVd = Cclm * (dPvagTerm_dVd / PvagTerm + dT0_dVd / T0 + dT1_dVd / T1);
        VACLM = Cclm * diffVds;

        dVACLM_dVg = dCclm_dVg * diffVds - dVdseff_dVg * Cclm;
        dVACLM_dVb = dCclm_dVb * diffVds - dVdseff_dVb * Cclm;
        dVACLM_dVd = dCclm_dVd * diffVds + (1.0 - dVdseff_dVd) * Cclm;
    } else {
        VACLM = Cclm = B4_MAX_EXP;
        dVACLM_dVd = dVACLM_dVg = dVACLM_dVb = 0.0;
        dCclm_dVd = dCclm_dVg = dCclm_dVb = 0.0;
    }

    NGB_CTA_ALIGN();
    /* ---- VADIBL ---- */
    double VADIBL, dVADIBL_dVg, dVADIBL_dVd, dVADIBL_dVb;
    if (B4P(thetaRout) > B4_MIN_EXP) {
        T8 = Abulk * Vdsat;
        T0 = Vgst2Vtm * T8;
        dT0_dVg = Vgst2Vtm * Abulk * dVdsat_dVg + T8 + Vgst2Vtm * Vdsat * dAbulk_dVg;
        dT0_dVb = Vgst2Vtm * (dAbulk_dVb * Vdsat + Abulk * dVdsat_dVb);
        dT0_dVd = Vgst2Vtm * Abulk * dVdsat_dVd;

        T1 = Vgst2Vtm + T8;
        dT1_dVg = 1.0 + Abulk * dVdsat_dVg + Vdsat * dAbulk_dVg;
        dT1_dVb = Abulk * dVdsat_dVb + dAbulk_dVb * Vdsat;
        dT1_dVd = Abulk * dVdsat_dVd;

        T9 = T1 * T1;
        T2 = B4P(thetaRout);
        VADIBL = (Vgst2Vtm - T0 / T1) / T2;
        dVADIBL_dVg = (1.0 - dT0_dVg / T1 + T0 * dT1_dVg / T9) / T2;
        dVADIBL_dVb = (-dT0_dVb / T1 + T0 * dT1_dVb / T9) / T2;
        dVADIBL_dVd = (-dT0_dVd / T1 + T0 * dT1_dVd / T9) / T2;

        T7 = B4P(pdiblb) * Vbseff;
        if (T7 >= -0.9) {
            T3 = 1.0 / (1.0 + T7);
            VADIBL *= T3;
            dVADIBL_dVg *= T3;
            dVADIBL_dVb = (dVADIBL_dVb - VADIBL * B4P(pdiblb)) * T3;
            dVADIBL_dVd *= T3;
        } else {
            T4 = 1.0 / (0.8 + T7);
            T3 = (17.0 + 20.0 * T7) * T4;
            dVADIBL_dVg *= T3;
            dVADIBL_dVb = dVADIBL_dVb * T3 - VADIBL * B4P(pdiblb) * T4 * T4;
            dVADIBL_dVd *= T3;
            VADIBL *= T3;
        }

        dVADIBL_dVg = dVADIBL_dVg * PvagTerm + VADIBL * dPvagTerm_dVg;
        dVADIBL_dVb = dVADIBL_dVb * PvagTerm + VADIBL * dPvagTerm_dVb;
        dVADIBL_dVd = dVADIBL_dVd * PvagTerm + VADIBL * dPvagTerm_dVd;
        VADIBL *= PvagTerm;
    } else {
        VADIBL = B4_MAX_EXP;
        dVADIBL_dVd = dVADIBL_dVg = dVADIBL_dVb = 0.0;
    }

    NGB_CTA_ALIGN();
    /* ---- Va ---- */
    const double Va = Vasat + VACLM;
    const double dVa_dVg = dVasat_dVg + dVACLM_dVg;
    const double dVa_dVb = dVasat_dVb + dVACLM_dVb;
    const double dVa_dVd = dVasat_dVd + dVACLM_dVd;

    NGB_CTA_ALIGN();
    /* ---- VADITS ---- */
    double VADITS, dVADITS_dVg, dVADITS_dVd;
    T0 = B4P(pditsd) * Vds;
    if (T0 > B4_EXP_THRESHOLD) { T1 = B4_MAX_EXP; dT1_dVd = 0; }
    else { T1 = ngb_exp(T0); dT1_dVd = T1 * B4P(pditsd); }
    if (B4P(pdits) > B4_MIN_EXP) {
        T2 = 1.0 + B4M(pditsl) * Leff;
        VADITS = (1.0 + T2 * T1) / B4P(pdits);
        dVADITS_dVg = VADITS * dFP_dVg;
        dVADITS_dVd = FP * T2 * dT1_dVd / B4P(pdits);
        VADITS *= FP;
    } else {
        VADITS = B4_MAX_EXP;
        dVADITS_dVg = dVADITS_dVd = 0;
    }

    NGB_CTA_ALIGN();
    /* ---- VASCBE ---- */
    double VASCBE, dVASCBE_dVg, dVASCBE_dVd, dVASCBE_dVb;
    if ((B4P(pscbe2) > 0.0) && (B4P(pscbe1) >= 0.0)) {
        if (diffVds > B4P(pscbe1) * B4P(litl) / B4_EXP_THRESHOLD) {
            T0 = B4P(pscbe1) * B4P(litl) / diffVds;
            VASCBE = Leff * ngb_exp(T0) / B4P(pscbe2);
            T1 = T0 * VASCBE / diffVds;
            dVASCBE_dVg = T1 * dVdseff_dVg;
            dVASCBE_dVd = -T1 * (1.0 - dVdseff_dVd);
            dVASCBE_dVb = T1 * dVdseff_dVb;
        } else {
            VASCBE = B4_MAX_EXP * Leff / B4P(pscbe2);
            dVASCBE_dVg = dVASCBE_dVd = dVASCBE_dVb = 0.0;
        }
    } else {
        VASCBE = B4_MAX_EXP;
        dVASCBE_dVg = dVASCBE_dVd = dVASCBE_dVb = 0.0;
    }

    NGB_CTA_ALIGN();
    /* ---- add DIBL to Ids ---- */
    double Idsa, dIdsa_dVg, dIdsa_dVd, dIdsa_dVb;
    T9 = diffVds / VADIBL;
    T0 = 1.0 + T9;
    Idsa = Idl * T0;
    dIdsa_dVg = T0 * dIdl_dVg - Idl * (dVdseff_dVg + T9 * dVADIBL_dVg) / VADIBL;
    dIdsa_dVd = T0 * dIdl_dVd + Idl * (1.0 - dVdseff_dVd - T9 * dVADIBL_dVd) / VADIBL;
    dIdsa_dVb = T0 * dIdl_dVb - Idl * (dVdseff_dVb + T9 * dVADIBL_dVb) / VADIBL;

    NGB_CTA_ALIGN();
    /* ---- add DITS to Ids ---- */
    T9 = diffVds / VADITS;
    T0 = 1.0 + T9;
    dIdsa_dVg = T0 * dIdsa_dVg - Idsa * (dVdseff_dVg + T9 * dVADITS_dVg) / VADITS;
    dIdsa_dVd = T0 * dIdsa_dVd + Idsa * (1.0 - dVdseff_dVd - T9 * dVADITS_dVd) / VADITS;
    dIdsa_dVb = T0 * dIdsa_dVb - Idsa * dVdseff_dVb / VADITS;
    Idsa *= T0;

    NGB_CTA_ALIGN();
    /* ---- add CLM to Ids ---- */
    T0 = ngb_log(Va / Vasat);
    dT0_dVg = dVa_dVg / Va - dVasat_dVg / Vasat;
    dT0_dVb = dVa_dVb / Va - dVasat_dVb / Vasat;
    dT0_dVd = dVa_dVd / Va - dVasat_dVd / Vasat;
    T1 = T0 / Cclm;
    T9 = 1.0 + T1;
    dT9_dVg = (dT0_dVg - T1 * dCclm_dVg) / Cclm;
    dT9_dVb = (dT0_dVb - T1 * dCclm_dVb) / Cclm;
    dT9_dVd = (dT0_dVd - T1 * dCclm_dVd) / Cclm;

    dIdsa_dVg = dIdsa_dVg * T9 + Idsa * dT9_dVg;
    dIdsa_dVb = dIdsa_dVb * T9 + Idsa * dT9_dVb;
    dIdsa_dVd = dIdsa_dVd * T9 + Idsa * dT9_dVd;
    Idsa *= T9;

    NGB_CTA_ALIGN();
    /* ---- substrate current ---- */
    double Isub, Gbd, Gbb, Gbg;
    tmp = B4P(alpha0) + B4P(alpha1) * Leff;
    if ((tmp <= 0.0) || (B4P(beta0) <= 0.0)) {
        Isub = Gbd = Gbb = Gbg = 0.0;
    } else {
        T2 = tmp / Leff;
        if (diffVds > B4P(beta0) / B4_EXP_THRESHOLD) {
            T0 = -B4P(beta0) / diffVds;
            T1 = T2 * diffVds * ngb_exp(T0);
            T3 = T1 / diffVds * (T0 - 1.0);
            dT1_dVg = T3 * dVdseff_dVg;
            dT1_dVd = T3 * (dVdseff_dVd - 1.0);
            dT1_dVb = T3 * dVdseff_dVb;
        } else {
            T3 = T2 * B4_MIN_EXP;
            T1 = T3 * diffVds;
            dT1_dVg = -T3 * dVdseff_dVg;
            dT1_dVd = T3 * (1.0 - dVdseff_dVd);
            dT1_dVb = -T3 * dVdseff_dVb;
        }
        T4 = Idsa * Vdseff;
        Isub = T1 * T4;
        Gbg = T1 * (dIdsa_dVg * Vdseff + Idsa * dVdseff_dVg) + T4 * dT1_dVg;
        Gbd = T1 * (dIdsa_dVd * Vdseff + Idsa * dVdseff_dVd) + T4 * dT1_dVd;
        Gbb = T1 * (dIdsa_dVb * Vdseff + Idsa * dVdseff_dVb) + T4 * dT1_dVb;

        Gbd += Gbg * dVgsteff_dVd;
        Gbb += Gbg * dVgsteff_dVb;
        Gbg *= dVgsteff_dVg;
        Gbb *= dVbseff_dVb;
    }
    w->csub = Isub; w->gbbs = Gbb; w->gbgs = Gbg; w->gbds = Gbd;

    NGB_CTA_ALIGN();
    /* ---- add SCBE to Ids ---- */
    double Ids, Gm, Gds, Gmb, cdrain;
    T9 = diffVds / VASCBE;
    T0 = 1.0 + T9;
    Ids = Idsa * T0;

    Gm = T0 * dIdsa_dVg - Idsa * (dVdseff_dVg + T9 * dVASCBE_dVg) / VASCBE;
    Gds = T0 * dIdsa_dVd + Idsa * (1.0 - dVdseff_dVd - T9 * dVASCBE_dVd) / VASCBE;
    Gmb = T0 * dIdsa_dVb - Idsa * (dVdseff_dVb + T9 * dVASCBE_dVb) / VASCBE;

    tmp1 = Gds + Gm * dVgsteff_dVd;
    tmp2 = Gmb + Gm * dVgsteff_dVb;
    tmp3 = Gm;

    Gm = (Ids * dVdseff_dVg + Vdseff * tmp3) * dVgsteff_dVg;
    Gds = Ids * (dVdseff_dVd + dVdseff_dVg * dVgsteff_dVd) + Vdseff * tmp1;
    Gmb = (Ids * (dVdseff_dVb + dVdseff_dVg * dVgsteff_dVb) + Vdseff * tmp2) * dVbseff_dVb;

    cdrain = Ids * Vdseff;

    NGB_CTA_ALIGN();
    /* ---- source-end velocity limit ---- */
    if (((int)B4M(vtlGiven)) && (B4M(vtl) > 0.0)) {
        double vs, dvs_dVg, dvs_dVd, dvs_dVb, Fsevl, dFsevl_dVg, dFsevl_dVd, dFsevl_dVb;
        T12 = 1.0 / Leff / CoxeffWovL;
        T11 = T12 / Vgsteff;
        T10 = -T11 / Vgsteff;
        vs = cdrain * T11;
        dvs_dVg = Gm * T11 + cdrain * T10 * dVgsteff_dVg;
        dvs_dVd = Gds * T11 + cdrain * T10 * dVgsteff_dVd;
        dvs_dVb = Gmb * T11 + cdrain * T10 * dVgsteff_dVb;
        T0 = 2 * 3;                                       /* 2 * MM */
        T1 = vs / (B4P(vtl) * B4P(tfactor));
        if (T1 > 0.0) {
            T2 = 1.0 + ngb_exp(T0 * ngb_log(T1));
            T3 = (T2 - 1.0) * T0 / vs;
            Fsevl = 1.0 / ngb_exp(ngb_log(T2) / T0);
            dT2_dVg = T3 * dvs_dVg;
            dT2_dVd = T3 * dvs_dVd;
            dT2_dVb = T3 * dvs_dVb;
            T4 = -1.0 / T0 * Fsevl / T2;
            dFsevl_dVg = T4 * dT2_dVg;
            dFsevl_dVd = T4 * dT2_dVd;
            dFsevl_dVb = T4 * dT2_dVb;
        } else {
            Fsevl = 1.0;
            dFsevl_dVg = 0.0;
            dFsevl_dVd = 0.0;
            dFsevl_dVb = 0.0;
        }
        Gm *= Fsevl;
        Gm += cdrain * dFsevl_dVg;
        Gmb *= Fsevl;
        Gmb += cdrain * dFsevl_dVb;
        Gds *= Fsevl;
        Gds += cdrain * dFsevl_dVd;
        cdrain *= Fsevl;
    }

    /* hand-off to the later phases */
    w->Vtm = Vtm; w->Vtm0 = Vtm0; w->Leff = Leff;
    w->Vbseff = Vbseff; w->dVbseff_dVb = dVbseff_dVb;
    w->Phis = Phis; w->dPhis_dVb = dPhis_dVb; w->sqrtPhis = sqrtPhis; w->dsqrtPhis_dVb = dsqrtPhis_dVb;
    w->Vth = Vth; w->dVth_dVb = dVth_dVb; w->dVth_dVd = dVth_dVd;
    w->n = n; w->dn_dVb = dn_dVb; w->dn_dVd = dn_dVd;
    w->Vgs_eff = Vgs_eff; w->dVgs_eff_dVg = dVgs_eff_dVg; w->Vgst = Vgst;
    w->Vgsteff = Vgsteff; w->dVgsteff_dVg = dVgsteff_dVg; w->dVgsteff_dVd = dVgsteff_dVd;
    w->dVgsteff_dVb = dVgsteff_dVb;
    w->Vdseff = Vdseff; w->dVdseff_dVg = dVdseff_dVg; w->dVdseff_dVd = dVdseff_dVd;
    w->dVdseff_dVb = dVdseff_dVb;
    w->Abulk0_Q = Abulk0_Q; w->dAbulk0_Q_dVb = dAbulk0_Q_dVb;
    w->cdrain = cdrain; w->gm = Gm; w->gds = Gds; w->gmbs = Gmb;
    w->beta = beta; w->dbeta_dVg = dbeta_dVg; w->dbeta_dVd = dbeta_dVd; w->dbeta_dVb = dbeta_dVb;
    w->Ids = Ids; w->tmp1 = tmp1; w->tmp2 = tmp2; w->tmp3 = tmp3;
    (void)dPvagTerm_dVd; (void)dT4_dVd; (void)dT3_dVd; (void)flags; (void)dT5_dVd;
}

/* one GIDL/GISL branch, gidlMod == 0 (b4ld.c:2324-2358 GIDL, :2365-2399 GISL) */
NGB_HD_SHARED void b4_gidl0(double T1, double dvg_eff, double T0den, double agidl, double bgidl, double cgidl,
                     double weffCJ, double vb, double *I, double *Gd, double *Gg, double *Gb)
{
    if ((agidl <= 0.0) || (bgidl <= 0.0) || (T1 <= 0.0) || (cgidl <= 0.0) || (vb > 0.0)) {
        *I = *Gd = *Gg = *Gb = 0.0;
    } else {
        double dT1_dVd = 1.0 / T0den;
        double dT1_dVg = -dvg_eff * dT1_dVd;
        double T2 = bgidl / T1, T3, T4, T5, T6, T7, T8, Ig, Ggd, Ggg;
        if (T2 < 100.0) {
            Ig = agidl * weffCJ * T1 * ngb_exp(-T2);
            T3 = Ig * (1.0 + T2) / T1;
            Ggd = T3 * dT1_dVd;
            Ggg = T3 * dT1_dVg;
        } else {
            Ig = agidl * weffCJ * 3.720075976e-44;
            Ggd = Ig * dT1_dVd;
            Ggg = Ig * dT1_dVg;
            Ig *= T1;
        }
        T4 = vb * vb;
        T5 = -vb * T4;
        T6 = cgidl + T5;
        T7 = T5 / T6;
        T8 = 3.0 * cgidl * T4 / T6 / T6;
        *Gd = Ggd * T7 + Ig * T8;
        *Gg = Ggg * T7;
        *Gb = -Ig * T8;
        *I = Ig * T7;
    }
}

/* one GIDL/GISL branch, gidlMod != 0 (b4ld.c:2409-2459 GISL, :2467-2516 GIDL) */
NGB_HD_SHARED void b4_gidl1(double T1, double dvg_eff, double T0den, double agidl, double bgidl, double cgidl,
                     double rgidl, double kgidl, double fgidl, double gidlclamp, double weffCJ,
                     double vb, double *I, double *Gd, double *Gg, double *Gb)
{
    if ((agidl <= 0.0) || (bgidl <= 0.0) || (T1 <= 0.0) || (cgidl < 0.0)) {
        *I = *Gd = *Gg = *Gb = 0.0;
    } else {
        double dT1_dVd = 1 / T0den;
        double dT1_dVg = -rgidl * dT1_dVd * dvg_eff;
        double T2 = bgidl / T1, T3, T4, T5, T6, Ig, Ggd, Ggg, Ggb;
        if (T2 < B4_EXPL_THRESHOLD) {
            Ig = weffCJ * agidl * T1 * ngb_exp(-T2);
            T3 = Ig / T1 * (T2 + 1);
            Ggd = T3 * dT1_dVd;
            Ggg = T3 * dT1_dVg;
        } else {
            T3 = weffCJ * agidl * B4_MIN_EXPL;
            Ig = T3 * T1;
            Ggd = T3 * dT1_dVd;
            Ggg = T3 * dT1_dVg;
        }
        T4 = vb - fgidl;
        if (T4 > gidlclamp) T4 = gidlclamp;
        if (T4 == 0) T5 = B4_EXPL_THRESHOLD;
        else T5 = kgidl / T4;
        if (T5 < B4_EXPL_THRESHOLD) {
            T6 = ngb_exp(T5);
            Ggb = -Ig * T6 * T5 / T4;
        } else {
            T6 = B4_MAX_EXPL;
            Ggb = 0.0;
        }
        *Gd = Ggd * T6;
        *Gg = Ggg * T6;
        *Gb = Ggb;
        *I = Ig * T6;
    }
}

/* edge (gate-to-S/D overlap) tunnelling current (b4ld.c:2727-2755 source, :2758-2785 drain) */
NGB_HD_SHARED void b4_ig_edge(double vg, double vfbsd_tot, double Aechvb, double BechvbEdge,
                       double aig, double big, double cig, double *Ig, double *dIg_dVg)
{
    double T0 = vg - vfbsd_tot;
    double vg_eff = sqrt(T0 * T0 + 1.0e-4);
    double dvg_eff = T0 / vg_eff;
    double T2 = vg * vg_eff;
    double dT2_dVg = vg * dvg_eff + vg_eff;
    double T3 = aig * cig - big;
    double T4 = big * cig;
    double T5 = BechvbEdge * (aig + T3 * vg_eff - T4 * vg_eff * vg_eff);
    double T6, dT6_dVg;
    if (T5 > B4_EXP_THRESHOLD) { T6 = B4_MAX_EXP; dT6_dVg = 0.0; }
    else if (T5 < -B4_EXP_THRESHOLD) { T6 = B4_MIN_EXP; dT6_dVg = 0.0; }
    else { T6 = ngb_exp(T5); dT6_dVg = T6 * BechvbEdge * (T3 - 2.0 * T4 * vg_eff) * dvg_eff; }
    *Ig = Aechvb * T2 * T6;
    *dIg_dVg = Aechvb * (T2 * dT6_dVg + T6 * dT2_dVg);
}

/* Phase D: gate resistance network, bias-dependent S/D resistance, GIDL/GISL, gate
 * tunnelling, finger scaling (b4ld.c:2191-2976). */
template <unsigned VK>
NGB_HD void b4_parasitics(const B4Ctx *c, size_t t, const double *Mrow, const double *Prow, int b4dr,
                          int flags, B4W *w)
{
    const int rgateMod = B4SEL_RGATE(flags);
    const double nf = B4I(nf);
    const int mtrlMod = B4SEL(mtrlMod);
    const int igcMod = B4SEL(igcMod), igbMod = B4SEL(igbMod);
    const double toxe = w->toxe;
    const double vds = w->vds, vgs = w->vgs, vgd = w->vgd, vbs = w->vbs, vbd = w->vbd;
    const double Vgsteff = w->Vgsteff, dVgsteff_dVg = w->dVgsteff_dVg;
    const double dVgsteff_dVd = w->dVgsteff_dVd, dVgsteff_dVb = w->dVgsteff_dVb;
    const double Vgs_eff = w->Vgs_eff, dVgs_eff_dVg = w->dVgs_eff_dVg;
    const double Vbseff = w->Vbseff, dVbseff_dVb = w->dVbseff_dVb;
    const double Vdseff = w->Vdseff, dVdseff_dVg = w->dVdseff_dVg;
    const double dVdseff_dVd = w->dVdseff_dVd, dVdseff_dVb = w->dVdseff_dVb;
    double T0, T1, T2, T3, T4, T5, T6, T7, T8, T9, T10, T11, T12, T13, T14;
    double dT2_dVg, dT2_dVd, dT2_dVb, dT6_dVg, dT6_dVd, dT6_dVb, dT7_dVg, dT7_dVd, dT7_dVb;
    double dT8_dVg, dT8_dVd, dT8_dVb, dT9_dVg, dT9_dVd, dT9_dVb, dT10_dVg, dT10_dVd, dT10_dVb;

    NGB_CTA_ALIGN();
    /* ---- Rg ---- */
    w->gcrg = w->gcrgd = w->gcrgg = w->gcrgs = w->gcrgb = 0.0;
    if (rgateMod > 1) {      /* trnqsMod/acnqsMod are 0 on this path */
        double dT0_dVd, dT0_dVb, dT0_dVg;
        const double xrcrg1 = B4P(xrcrg1);
        T9 = B4P(xrcrg2) * B4M(vtm);
        T0 = T9 * w->beta;
        dT0_dVd = (w->dbeta_dVd + w->dbeta_dVg * dVgsteff_dVd) * T9;
        dT0_dVb = (w->dbeta_dVb + w->dbeta_dVg * dVgsteff_dVb) * T9;
        dT0_dVg = w->dbeta_dVg * T9;

        w->gcrg = xrcrg1 * (T0 + w->Ids);
        w->gcrgd = xrcrg1 * (dT0_dVd + w->tmp1);
        w->gcrgb = xrcrg1 * (dT0_dVb + w->tmp2) * dVbseff_dVb;
        w->gcrgg = xrcrg1 * (dT0_dVg + w->tmp3) * dVgsteff_dVg;

        if (nf != 1.0) {
            w->gcrg *= nf; w->gcrgg *= nf; w->gcrgd *= nf; w->gcrgb *= nf;
        }
        if (rgateMod == 2) {
            const double grgeltd = B4I(grgeltd);
            T10 = grgeltd * grgeltd;
            T11 = grgeltd + w->gcrg;
            w->gcrg = grgeltd * w->gcrg / T11;
            T12 = T10 / T11 / T11;
            w->gcrgg *= T12; w->gcrgd *= T12; w->gcrgb *= T12;
        }
        w->gcrgs = -(w->gcrgg + w->gcrgd + w->gcrgb);
    }

    NGB_CTA_ALIGN();
    /* ---- bias-dependent external S/D resistance ---- */
    if (B4SEL(rdsMod)) {
        double vgs_eff, dvgs_eff_dvg, vgd_eff, dvgd_eff_dvg, dT0_dvg, dT1_dvb, dT3_dvg, dT3_dvb;
        double Rs, dRs_dvg, dRs_dvb, Rd, dRd_dvg, dRd_dvb;
        double dgstot_dvd, dgstot_dvg, dgstot_dvb, dgstot_dvs;
        double dgdtot_dvd, dgdtot_dvg, dgdtot_dvb, dgdtot_dvs;
        const double vfbsd = B4P(vfbsd), prwg = B4P(prwg), prwb = B4P(prwb);
        /* Rs(V) */
        T0 = vgs - vfbsd;
        T1 = sqrt(T0 * T0 + 1.0e-4);
        vgs_eff = 0.5 * (T0 + T1);
        dvgs_eff_dvg = vgs_eff / T1;

        T0 = 1.0 + prwg * vgs_eff;
        dT0_dvg = -prwg / T0 / T0 * dvgs_eff_dvg;
        T1 = -prwb * vbs;
        dT1_dvb = -prwb;

        T2 = 1.0 / T0 + T1;
        T3 = T2 + sqrt(T2 * T2 + 0.01);
        dT3_dvg = T3 / (T3 - T2);
        dT3_dvb = dT3_dvg * dT1_dvb;
        dT3_dvg *= dT0_dvg;

        T4 = B4P(rs0) * 0.5;
        Rs = B4P(rswmin) + T3 * T4;
        dRs_dvg = T4 * dT3_dvg;
        dRs_dvb = T4 * dT3_dvb;

        T0 = 1.0 + B4I(sourceConductance) * Rs;
        w->gstot = B4I(sourceConductance) / T0;
        T0 = -w->gstot * w->gstot;
        dgstot_dvd = 0.0;
        dgstot_dvg = T0 * dRs_dvg;
        dgstot_dvb = T0 * dRs_dvb;
        dgstot_dvs = -(dgstot_dvg + dgstot_dvb + dgstot_dvd);

        /* Rd(V) */
        T0 = vgd - vfbsd;
        T1 = sqrt(T0 * T0 + 1.0e-4);
        vgd_eff = 0.5 * (T0 + T1);
        dvgd_eff_dvg = vgd_eff / T1;

        T0 = 1.0 + prwg * vgd_eff;
        dT0_dvg = -prwg / T0 / T0 * dvgd_eff_dvg;
        T1 = -prwb * vbd;
        dT1_dvb = -prwb;

        T2 = 1.0 / T0 + T1;
        T3 = T2 + sqrt(T2 * T2 + 0.01);
        dT3_dvg = T3 / (T3 - T2);
        dT3_dvb = dT3_dvg * dT1_dvb;
        dT3_dvg *= dT0_dvg;

        T4 = B4P(rd0) * 0.5;
        Rd = B4P(rdwmin) + T3 * T4;
        dRd_dvg = T4 * dT3_dvg;
        dRd_dvb = T4 * dT3_dvb;

        T0 = 1.0 + B4I(drainConductance) * Rd;
        w->gdtot = B4I(drainConductance) / T0;
        T0 = -w->gdtot * w->gdtot;
        dgdtot_dvs = 0.0;
        dgdtot_dvg = T0 * dRd_dvg;
        dgdtot_dvb = T0 * dRd_dvb;
        dgdtot_dvd = -(dgdtot_dvg + dgdtot_dvb + dgdtot_dvs);

        w->gstotd = w->vses * dgstot_dvd;
        w->gstotg = w->vses * dgstot_dvg;
        w->gstots = w->vses * dgstot_dvs;
        w->gstotb = w->vses * dgstot_dvb;

        T2 = w->vdes - vds;
        w->gdtotd = T2 * dgdtot_dvd;
        w->gdtotg = T2 * dgdtot_dvg;
        w->gdtots = T2 * dgdtot_dvs;
        w->gdtotb = T2 * dgdtot_dvb;
    } else {
        w->gstot = w->gstotd = w->gstotg = w->gstots = w->gstotb = 0.0;
        w->gdtot = w->gdtotd = w->gdtotg = w->gdtots = w->gdtotb = 0.0;
    }

    NGB_CTA_ALIGN();
    /* ---- GIDL / GISL ---- */
    {
        const double weffCJ = B4P(weffCJ);
        const double vfbsd_add = (mtrlMod == 0) ? 0.0 : B4P(vfbsd);
        double gI, gGd, gGg, gGb;
        if (mtrlMod == 0) T0 = 3.0 * toxe;
        else T0 = B4M(epsrsub) * toxe / w->epsrox;

        if (B4SEL(gidlMod) == 0) {
            if (mtrlMod == 0) T1 = (vds - w->vgs_eff - B4P(egidl)) / T0;
            else T1 = (vds - w->vgs_eff - B4P(egidl) + vfbsd_add) / T0;
            b4_gidl0(T1, w->dvgs_eff_dvg, T0, B4P(agidl), B4P(bgidl), B4P(cgidl), weffCJ, vbd,
                     &gI, &gGd, &gGg, &gGb);
            w->Igidl = gI; w->ggidld = gGd; w->ggidlg = gGg; w->ggidlb = gGb;
            if (mtrlMod == 0) T1 = (-vds - w->vgd_eff - B4P(egisl)) / T0;
            else T1 = (-vds - w->vgd_eff - B4P(egisl) + vfbsd_add) / T0;
            b4_gidl0(T1, w->dvgd_eff_dvg, T0, B4P(agisl), B4P(bgisl), B4P(cgisl), weffCJ, vbs,
                     &gI, &gGd, &gGg, &gGb);
            w->Igisl = gI; w->ggisls = gGd; w->ggislg = gGg; w->ggislb = gGb;
        } else {
            const double gidlclamp = B4M(gidlclamp);
            if (mtrlMod == 0) T1 = (-vds - B4P(rgisl) * w->vgd_eff - B4P(egisl)) / T0;
            else T1 = (-vds - B4P(rgisl) * w->vgd_eff - B4P(egisl) + vfbsd_add) / T0;
            b4_gidl1(T1, w->dvgd_eff_dvg, T0, B4P(agisl), B4P(bgisl), B4P(cgisl), B4P(rgisl),
                     B4P(kgisl), B4P(fgisl), gidlclamp, weffCJ, vbs,
                     &gI, &gGd, &gGg, &gGb);
            w->Igisl = gI; w->ggisls = gGd; w->ggislg = gGg; w->ggislb = gGb;
            if (mtrlMod == 0) T1 = (vds - B4P(rgidl) * w->vgs_eff - B4P(egidl)) / T0;
            else T1 = (vds - B4P(rgidl) * w->vgs_eff - B4P(egidl) + vfbsd_add) / T0;
            b4_gidl1(T1, w->dvgs_eff_dvg, T0, B4P(agidl), B4P(bgidl), B4P(cgidl), B4P(rgidl),
                     B4P(kgidl), B4P(fgidl), gidlclamp, weffCJ, vbd,
                     &gI, &gGd, &gGg, &gGb);
            w->Igidl = gI; w->ggidld = gGd; w->ggidlg = gGg; w->ggidlb = gGb;
        }
    }

    NGB_CTA_ALIGN();
    /* ---- gate tunnelling current ---- */
    double Vfb = 0.0, Voxacc = 0.0, dVoxacc_dVg = 0.0, dVoxacc_dVb = 0.0;
    double Voxdepinv = 0.0, dVoxdepinv_dVg = 0.0, dVoxdepinv_dVd = 0.0, dVoxdepinv_dVb = 0.0;
    double VxNVt = 0.0, ExpVxNVt, Vaux = 0.0, dVaux_dVg = 0.0, dVaux_dVd = 0.0, dVaux_dVb = 0.0;
    const double k1ox = B4P(k1ox);
    if ((igcMod != 0) || (igbMod != 0)) {
        double V3, Vfbeff, dVfbeff_dVg, dVfbeff_dVb;
        Vfb = B4I(vfbzb);
        V3 = Vfb - Vgs_eff + Vbseff - B4_DELTA_3;
        if (Vfb <= 0.0) T0 = sqrt(V3 * V3 - 4.0 * B4_DELTA_3 * Vfb);
        else T0 = sqrt(V3 * V3 + 4.0 * B4_DELTA_3 * Vfb);
        T1 = 0.5 * (1.0 + V3 / T0);
        Vfbeff = Vfb - 0.5 * (V3 + T0);
        dVfbeff_dVg = T1 * dVgs_eff_dVg;
        dVfbeff_dVb = -T1;

        Voxacc = Vfb - Vfbeff;
        dVoxacc_dVg = -dVfbeff_dVg;
        dVoxacc_dVb = -dVfbeff_dVb;
        if (Voxacc < 0.0) Voxacc = dVoxacc_dVg = dVoxacc_dVb = 0.0;

        T0 = 0.5 * k1ox;
        T3 = Vgs_eff - Vfbeff - Vbseff - Vgsteff;
        if (k1ox == 0.0) {
            Voxdepinv = dVoxdepinv_dVg = dVoxdepinv_dVd = dVoxdepinv_dVb = 0.0;
        } else if (T3 < 0.0) {
            Voxdepinv = -T3;
            dVoxdepinv_dVg = -dVgs_eff_dVg + dVfbeff_dVg + dVgsteff_dVg;
            dVoxdepinv_dVd = dVgsteff_dVd;
            dVoxdepinv_dVb = dVfbeff_dVb + 1.0 + dVgsteff_dVb;
        } else {
            T1 = sqrt(T0 * T0 + T3);
            T2 = T0 / T1;
            Voxdepinv = k1ox * (T1 - T0);
            dVoxdepinv_dVg = T2 * (dVgs_eff_dVg - dVfbeff_dVg - dVgsteff_dVg);
            dVoxdepinv_dVd = -T2 * dVgsteff_dVd;
            dVoxdepinv_dVb = -T2 * (dVfbeff_dVb + 1.0 + dVgsteff_dVb);
        }
        Voxdepinv += Vgsteff;
        dVoxdepinv_dVg += dVgsteff_dVg;
        dVoxdepinv_dVd += dVgsteff_dVd;
        dVoxdepinv_dVb += dVgsteff_dVb;
    }

    const double tmpV = (B4SEL(tempMod) < 2) ? w->Vtm : w->Vtm0;
    if (igcMod) {
        const double type = B4M(type), vth0 = B4I(vth0);
        double Igc, dIgc_dVg, dIgc_dVd, dIgc_dVb, Pigcd, dPigcd_dVg, dPigcd_dVd, dPigcd_dVb;
        double Igcs, dIgcs_dVg, dIgcs_dVd, dIgcs_dVb, Igcd, dIgcd_dVg, dIgcd_dVd, dIgcd_dVb;
        T0 = tmpV * B4P(nigc);
        if (igcMod == 1) {
            VxNVt = (Vgs_eff - type * vth0) / T0;
            if (VxNVt > B4_EXP_THRESHOLD) {
                Vaux = Vgs_eff - type * vth0;
                dVaux_dVg = dVgs_eff_dVg;
                dVaux_dVd = 0.0;
                dVaux_dVb = 0.0;
            }
        } else if (igcMod == 2) {
            VxNVt = (Vgs_eff - w->von) / T0;
            if (VxNVt > B4_EXP_THRESHOLD) {
                Vaux = Vgs_eff - w->von;
                dVaux_dVg = dVgs_eff_dVg;
                dVaux_dVd = -w->dVth_dVd;
                dVaux_dVb = -w->dVth_dVb;
            }
        }
        if (VxNVt < -B4_EXP_THRESHOLD) {
            Vaux = T0 * ngb_log(1.0 + B4_MIN_EXP);
            dVaux_dVg = dVaux_dVd = dVaux_dVb = 0.0;
        } else if ((VxNVt >= -B4_EXP_THRESHOLD) && (VxNVt <= B4_EXP_THRESHOLD)) {
            ExpVxNVt = ngb_exp(VxNVt);
            Vaux = T0 * ngb_log(1.0 + ExpVxNVt);
            dVaux_dVg = ExpVxNVt / (1.0 + ExpVxNVt);
            if (igcMod == 1) {
                dVaux_dVd = 0.0;
                dVaux_dVb = 0.0;
            } else if (igcMod == 2) {
                dVaux_dVd = -dVaux_dVg * w->dVth_dVd;
                dVaux_dVb = -dVaux_dVg * w->dVth_dVb;
            }
            dVaux_dVg *= dVgs_eff_dVg;
        }

        T2 = w->Vgs * Vaux;
        dT2_dVg = Vaux + w->Vgs * dVaux_dVg;
        dT2_dVd = w->Vgs * dVaux_dVd;
        dT2_dVb = w->Vgs * dVaux_dVb;

        T11 = B4P(Aechvb);
        T12 = B4P(Bechvb);
        T3 = B4P(aigc) * B4P(cigc) - B4P(bigc);
        T4 = B4P(bigc) * B4P(cigc);
        T5 = T12 * (B4P(aigc) + T3 * Voxdepinv - T4 * Voxdepinv * Voxdepinv);

        if (T5 > B4_EXP_THRESHOLD) {
            T6 = B4_MAX_EXP;
            dT6_dVg = dT6_dVd = dT6_dVb = 0.0;
        } else if (T5 < -B4_EXP_THRESHOLD) {
            T6 = B4_MIN_EXP;
            dT6_dVg = dT6_dVd = dT6_dVb = 0.0;
        } else {
            T6 = ngb_exp(T5);
            dT6_dVg = T6 * T12 * (T3 - 2.0 * T4 * Voxdepinv);
            dT6_dVd = dT6_dVg * dVoxdepinv_dVd;
            dT6_dVb = dT6_dVg * dVoxdepinv_dVb;
            dT6_dVg *= dVoxdepinv_dVg;
        }

        Igc = T11 * T2 * T6;
        dIgc_dVg = T11 * (T2 * dT6_dVg + T6 * dT2_dVg);
        dIgc_dVd = T11 * (T2 * dT6_dVd + T6 * dT2_dVd);
        dIgc_dVb = T11 * (T2 * dT6_dVb + T6 * dT2_dVb);

        if ((int)B4M(pigcdGiven)) {
            Pigcd = B4P(pigcd);
            dPigcd_dVg = dPigcd_dVd = dPigcd_dVb = 0.0;
        } else {
            T11 = -B4P(Bechvb);
            T12 = Vgsteff + 1.0e-20;
            T13 = T11 / T12 / T12;
            T14 = -T13 / T12;
            Pigcd = T13 * (1.0 - 0.5 * Vdseff / T12);
            dPigcd_dVg = T14 * (2.0 + 0.5 * (dVdseff_dVg - 3.0 * Vdseff / T12));
            dPigcd_dVd = 0.5 * T14 * dVdseff_dVd;
            dPigcd_dVb = 0.5 * T14 * dVdseff_dVb;
        }

        T7 = -Pigcd * Vdseff;
        dT7_dVg = -Vdseff * dPigcd_dVg - Pigcd * dVdseff_dVg;
        dT7_dVd = -Vdseff * dPigcd_dVd - Pigcd * dVdseff_dVd + dT7_dVg * dVgsteff_dVd;
        dT7_dVb = -Vdseff * dPigcd_dVb - Pigcd * dVdseff_dVb + dT7_dVg * dVgsteff_dVb;
        dT7_dVg *= dVgsteff_dVg;
        T8 = T7 * T7 + 2.0e-4;
        dT8_dVg = 2.0 * T7;
        dT8_dVd = dT8_dVg * dT7_dVd;
        dT8_dVb = dT8_dVg * dT7_dVb;
        dT8_dVg *= dT7_dVg;

        if (T7 > B4_EXP_THRESHOLD) {
            T9 = B4_MAX_EXP;
            dT9_dVg = dT9_dVd = dT9_dVb = 0.0;
        } else if (T7 < -B4_EXP_THRESHOLD) {
            T9 = B4_MIN_EXP;
            dT9_dVg = dT9_dVd = dT9_dVb = 0.0;
        } else {
            T9 = ngb_exp(T7);
            dT9_dVg = T9 * dT7_dVg;
            dT9_dVd = T9 * dT7_dVd;
            dT9_dVb = T9 * dT7_dVb;
        }

        T0 = T8 * T8;
        T1 = T9 - 1.0 + 1.0e-4;
        T10 = (T1 - T7) / T8;
        dT10_dVg = (dT9_dVg - dT7_dVg - T10 * dT8_dVg) / T8;
        dT10_dVd = (dT9_dVd - dT7_dVd - T10 * dT8_dVd) / T8;
        dT10_dVb = (dT9_dVb - dT7_dVb - T10 * dT8_dVb) / T8;

        Igcs = Igc * T10;
        dIgcs_dVg = dIgc_dVg * T10 + Igc * dT10_dVg;
        dIgcs_dVd = dIgc_dVd * T10 + Igc * dT10_dVd;
        dIgcs_dVb = dIgc_dVb * T10 + Igc * dT10_dVb;

        T1 = T9 - 1.0 - 1.0e-4;
        T10 = (T7 * T9 - T1) / T8;
        dT10_dVg = (dT7_dVg * T9 + (T7 - 1.0) * dT9_dVg - T10 * dT8_dVg) / T8;
        dT10_dVd = (dT7_dVd * T9 + (T7 - 1.0) * dT9_dVd - T10 * dT8_dVd) / T8;
        dT10_dVb = (dT7_dVb * T9 + (T7 - 1.0) * dT9_dVb - T10 * dT8_dVb) / T8;
        Igcd = Igc * T10;
        dIgcd_dVg = dIgc_dVg * T10 + Igc * dT10_dVg;
        dIgcd_dVd = dIgc_dVd * T10 + Igc * dT10_dVd;
        dIgcd_dVb = dIgc_dVb * T10 + Igc * dT10_dVb;

        w->Igcs = Igcs;
        w->gIgcsg = dIgcs_dVg;
        w->gIgcsd = dIgcs_dVd;
        w->gIgcsb = dIgcs_dVb * dVbseff_dVb;
        w->Igcd = Igcd;
        w->gIgcdg = dIgcd_dVg;
        w->gIgcdd = dIgcd_dVd;
        w->gIgcdb = dIgcd_dVb * dVbseff_dVb;

        {
            const double vfbsd_tot = B4P(vfbsd) + B4P(vfbsdoff);
            const double BechvbEdge = B4P(BechvbEdge);
            double eI, eG;
            b4_ig_edge(vgs, vfbsd_tot, B4P(AechvbEdgeS), BechvbEdge, B4P(aigs), B4P(bigs), B4P(cigs),
                       &eI, &eG);
            w->Igs = eI; w->gIgsg = eG;
            w->gIgss = -eG;
            b4_ig_edge(vgd, vfbsd_tot, B4P(AechvbEdgeD), BechvbEdge, B4P(aigd), B4P(bigd), B4P(cigd),
                       &eI, &eG);
            w->Igd = eI; w->gIgdg = eG;
            w->gIgdd = -eG;
        }
        (void)T0;
    } else {
        w->Igcs = w->gIgcsg = w->gIgcsd = w->gIgcsb = 0.0;
        w->Igcd = w->gIgcdg = w->gIgcdd = w->gIgcdb = 0.0;
        w->Igs = w->gIgsg = w->gIgss = 0.0;
        w->Igd = w->gIgdg = w->gIgdd = 0.0;
    }

    if (igbMod) {
        const double Vgs = w->Vgs, Vbs = w->Vbs;
        double Igbacc, dIgbacc_dVg, dIgbacc_dVb, Igbinv, dIgbinv_dVg, dIgbinv_dVd, dIgbinv_dVb;
        T0 = tmpV * B4P(nigbacc);
        T1 = -Vgs_eff + Vbseff + Vfb;
        VxNVt = T1 / T0;
        if (VxNVt > B4_EXP_THRESHOLD) {
            Vaux = T1;
            dVaux_dVg = -dVgs_eff_dVg;
            dVaux_dVb = 1.0;
        } else if (VxNVt < -B4_EXP_THRESHOLD) {
            Vaux = T0 * ngb_log(1.0 + B4_MIN_EXP);
            dVaux_dVg = dVaux_dVb = 0.0;
        } else {
            ExpVxNVt = ngb_exp(VxNVt);
            Vaux = T0 * ngb_log(1.0 + ExpVxNVt);
            dVaux_dVb = ExpVxNVt / (1.0 + ExpVxNVt);
            dVaux_dVg = -dVaux_dVb * dVgs_eff_dVg;
        }

        T2 = (Vgs - Vbs) * Vaux;
        dT2_dVg = Vaux + (Vgs - Vbs) * dVaux_dVg;
        dT2_dVb = -Vaux + (Vgs - Vbs) * dVaux_dVb;

        T11 = 4.97232e-7 * B4P(weff) * B4P(leff) * B4P(ToxRatio);
        T12 = -7.45669e11 * toxe;
        T3 = B4P(aigbacc) * B4P(cigbacc) - B4P(bigbacc);
        T4 = B4P(bigbacc) * B4P(cigbacc);
        T5 = T12 * (B4P(aigbacc) + T3 * Voxacc - T4 * Voxacc * Voxacc);

        if (T5 > B4_EXP_THRESHOLD) {
            T6 = B4_MAX_EXP;
            dT6_dVg = dT6_dVb = 0.0;
        } else if (T5 < -B4_EXP_THRESHOLD) {
            T6 = B4_MIN_EXP;
            dT6_dVg = dT6_dVb = 0.0;
        } else {
            T6 = ngb_exp(T5);
            dT6_dVg = T6 * T12 * (T3 - 2.0 * T4 * Voxacc);
            dT6_dVb = dT6_dVg * dVoxacc_dVb;
            dT6_dVg *= dVoxacc_dVg;
        }

        Igbacc = T11 * T2 * T6;
        dIgbacc_dVg = T11 * (T2 * dT6_dVg + T6 * dT2_dVg);
        dIgbacc_dVb = T11 * (T2 * dT6_dVb + T6 * dT2_dVb);

        T0 = tmpV * B4P(nigbinv);
        T1 = Voxdepinv - B4P(eigbinv);
        VxNVt = T1 / T0;
        if (VxNVt > B4_EXP_THRESHOLD) {
            Vaux = T1;
            dVaux_dVg = dVoxdepinv_dVg;
            dVaux_dVd = dVoxdepinv_dVd;
            dVaux_dVb = dVoxdepinv_dVb;
        } else if (VxNVt < -B4_EXP_THRESHOLD) {
            Vaux = T0 * ngb_log(1.0 + B4_MIN_EXP);
            dVaux_dVg = dVaux_dVd = dVaux_dVb = 0.0;
        } else {
            ExpVxNVt = ngb_exp(VxNVt);
            Vaux = T0 * ngb_log(1.0 + ExpVxNVt);
            dVaux_dVg = ExpVxNVt / (1.0 + ExpVxNVt);
            dVaux_dVd = dVaux_dVg * dVoxdepinv_dVd;
            dVaux_dVb = dVaux_dVg * dVoxdepinv_dVb;
            dVaux_dVg *= dVoxdepinv_dVg;
        }

        T2 = (Vgs - Vbs) * Vaux;
        dT2_dVg = Vaux + (Vgs - Vbs) * dVaux_dVg;
        dT2_dVd = (Vgs - Vbs) * dVaux_dVd;
        dT2_dVb = -Vaux + (Vgs - Vbs) * dVaux_dVb;

        T11 *= 0.75610;
        T12 *= 1.31724;
        T3 = B4P(aigbinv) * B4P(cigbinv) - B4P(bigbinv);
        T4 = B4P(bigbinv) * B4P(cigbinv);
        T5 = T12 * (B4P(aigbinv) + T3 * Voxdepinv - T4 * Voxdepinv * Voxdepinv);

        if (T5 > B4_EXP_THRESHOLD) {
            T6 = B4_MAX_EXP;
            dT6_dVg = dT6_dVd = dT6_dVb = 0.0;
        } else if (T5 < -B4_EXP_THRESHOLD) {
            T6 = B4_MIN_EXP;
            dT6_dVg = dT6_dVd = dT6_dVb = 0.0;
        } else {
            T6 = ngb_exp(T5);
            dT6_dVg = T6 * T12 * (T3 - 2.0 * T4 * Voxdepinv);
            dT6_dVd = dT6_dVg * dVoxdepinv_dVd;
            dT6_dVb = dT6_dVg * dVoxdepinv_dVb;
            dT6_dVg *= dVoxdepinv_dVg;
        }

        Igbinv = T11 * T2 * T6;
        dIgbinv_dVg = T11 * (T2 * dT6_dVg + T6 * dT2_dVg);
        dIgbinv_dVd = T11 * (T2 * dT6_dVd + T6 * dT2_dVd);
        dIgbinv_dVb = T11 * (T2 * dT6_dVb + T6 * dT2_dVb);

        w->Igb = Igbinv + Igbacc;
        w->gIgbg = dIgbinv_dVg + dIgbacc_dVg;
        w->gIgbd = dIgbinv_dVd;
        w->gIgbb = (dIgbinv_dVb + dIgbacc_dVb) * dVbseff_dVb;
    } else {
        w->Igb = w->gIgbg = w->gIgbd = w->gIgbs = w->gIgbb = 0.0;
    }

    NGB_CTA_ALIGN();
    /* ---- finger multiplication ---- */
    if (nf != 1.0) {
        w->cdrain *= nf; w->gds *= nf; w->gm *= nf; w->gmbs *= nf;
        w->gbbs *= nf; w->gbgs *= nf; w->gbds *= nf; w->csub *= nf;
        w->Igidl *= nf; w->ggidld *= nf; w->ggidlg *= nf; w->ggidlb *= nf;
        w->Igisl *= nf; w->ggisls *= nf; w->ggislg *= nf; w->ggislb *= nf;
        w->Igcs *= nf; w->gIgcsg *= nf; w->gIgcsd *= nf; w->gIgcsb *= nf;
        w->Igcd *= nf; w->gIgcdg *= nf; w->gIgcdd *= nf; w->gIgcdb *= nf;
        w->Igs *= nf; w->gIgsg *= nf; w->gIgss *= nf;
        w->Igd *= nf; w->gIgdg *= nf; w->gIgdd *= nf;
        w->Igb *= nf; w->gIgbg *= nf; w->gIgbd *= nf; w->gIgbb *= nf;
    }

    w->ggidls = -(w->ggidld + w->ggidlg + w->ggidlb);
    w->ggisld = -(w->ggisls + w->ggislg + w->ggislb);
    w->gIgbs = -(w->gIgbg + w->gIgbd + w->gIgbb);
    w->gIgcss = -(w->gIgcsg + w->gIgcsd + w->gIgcsb);
    w->gIgcds = -(w->gIgcdg + w->gIgcdd + w->gIgcdb);
    (void)vgd; (void)dT2_dVd; (void)VxNVt;
}

/* VgsteffCV selection shared by capMod 1 and 2 (b4ld.c:3351-3457) */
template <unsigned VK>
NGB_HD void b4_vgsteff_cv(const B4Ctx *c, const double *Mrow, const double *Prow, int b4dr, const B4W *w,
                          double *pVgsteff, double *pdVg, double *pdVd, double *pdVb)
{
    const double n = w->n, dn_dVd = w->dn_dVd, dn_dVb = w->dn_dVb, Vtm = w->Vtm, Vgst = w->Vgst;
    const double dVgs_eff_dVg = w->dVgs_eff_dVg, dVth_dVd = w->dVth_dVd, dVth_dVb = w->dVth_dVb;
    double Vgsteff, dVgsteff_dVg, dVgsteff_dVd, dVgsteff_dVb, T0, T1, T2, T3, T4, T5, T9, T10, T11;
    double dT9_dVg, dT9_dVd, dT9_dVb, dT10_dVg, dT10_dVd, dT10_dVb, ExpVgst;
    if (B4SEL(cvchargeMod) == 0) {
        const double noff = n * B4P(noff);
        const double dnoff_dVd = B4P(noff) * dn_dVd;
        const double dnoff_dVb = B4P(noff) * dn_dVb;
        const double voffcv = B4P(voffcv);
        double VgstNVt;
        T0 = Vtm * noff;
        VgstNVt = (Vgst - voffcv) / T0;
        if (VgstNVt > B4_EXP_THRESHOLD) {
            Vgsteff = Vgst - voffcv;
            dVgsteff_dVg = dVgs_eff_dVg;
            dVgsteff_dVd = -dVth_dVd;
            dVgsteff_dVb = -dVth_dVb;
        } else if (VgstNVt < -B4_EXP_THRESHOLD) {
            Vgsteff = T0 * ngb_log(1.0 + B4_MIN_EXP);
            dVgsteff_dVg = 0.0;
            dVgsteff_dVd = Vgsteff / noff;
            dVgsteff_dVb = dVgsteff_dVd * dnoff_dVb;
            dVgsteff_dVd *= dnoff_dVd;
        } else {
            ExpVgst = ngb_exp(VgstNVt);
            Vgsteff = T0 * ngb_log(1.0 + ExpVgst);
            dVgsteff_dVg = ExpVgst / (1.0 + ExpVgst);
            dVgsteff_dVd = -dVgsteff_dVg * (dVth_dVd + (Vgst - voffcv) / noff * dnoff_dVd)
                         + Vgsteff / noff * dnoff_dVd;
            dVgsteff_dVb = -dVgsteff_dVg * (dVth_dVb + (Vgst - voffcv) / noff * dnoff_dVb)
                         + Vgsteff / noff * dnoff_dVb;
            dVgsteff_dVg *= dVgs_eff_dVg;
        }
    } else {
        const double mstarcv = B4P(mstarcv), coxe = B4M(coxe), cdep0 = B4P(cdep0);
        T0 = n * Vtm;
        T1 = mstarcv * Vgst;
        T2 = T1 / T0;
        if (T2 > B4_EXP_THRESHOLD) {
            T10 = T1;
            dT10_dVg = mstarcv * dVgs_eff_dVg;
            dT10_dVd = -dVth_dVd * mstarcv;
            dT10_dVb = -dVth_dVb * mstarcv;
        } else if (T2 < -B4_EXP_THRESHOLD) {
            T10 = Vtm * ngb_log(1.0 + B4_MIN_EXP);
            dT10_dVg = 0.0;
            dT10_dVd = T10 * dn_dVd;
            dT10_dVb = T10 * dn_dVb;
            T10 *= n;
        } else {
            ExpVgst = ngb_exp(T2);
            T3 = Vtm * ngb_log(1.0 + ExpVgst);
            T10 = n * T3;
            dT10_dVg = mstarcv * ExpVgst / (1.0 + ExpVgst);
            dT10_dVb = T3 * dn_dVb - dT10_dVg * (dVth_dVb + Vgst * dn_dVb / n);
            dT10_dVd = T3 * dn_dVd - dT10_dVg * (dVth_dVd + Vgst * dn_dVd / n);
            dT10_dVg *= dVgs_eff_dVg;
        }

        T1 = B4P(voffcbncv) - (1.0 - mstarcv) * Vgst;
        T2 = T1 / T0;
        if (T2 < -B4_EXP_THRESHOLD) {
            T3 = coxe * B4_MIN_EXP / cdep0;
            T9 = mstarcv + T3 * n;
            dT9_dVg = 0.0;
            dT9_dVd = dn_dVd * T3;
            dT9_dVb = dn_dVb * T3;
        } else if (T2 > B4_EXP_THRESHOLD) {
            T3 = coxe * B4_MAX_EXP / cdep0;
            T9 = mstarcv + T3 * n;
            dT9_dVg = 0.0;
            dT9_dVd = dn_dVd * T3;
            dT9_dVb = dn_dVb * T3;
        } else {
            ExpVgst = ngb_exp(T2);
            T3 = coxe / cdep0;
            T4 = T3 * ExpVgst;
            T5 = T1 * T4 / T0;
            T9 = mstarcv + n * T4;
            dT9_dVg = T3 * (mstarcv - 1.0) * ExpVgst / Vtm;
            dT9_dVb = T4 * dn_dVb - dT9_dVg * dVth_dVb - T5 * dn_dVb;
            dT9_dVd = T4 * dn_dVd - dT9_dVg * dVth_dVd - T5 * dn_dVd;
            dT9_dVg *= dVgs_eff_dVg;
        }
        Vgsteff = T10 / T9;
        T11 = T9 * T9;
        dVgsteff_dVg = (T9 * dT10_dVg - T10 * dT9_dVg) / T11;
        dVgsteff_dVd = (T9 * dT10_dVd - T10 * dT9_dVd) / T11;
        dVgsteff_dVb = (T9 * dT10_dVb - T10 * dT9_dVb) / T11;
    }
    *pVgsteff = Vgsteff; *pdVg = dVgsteff_dVg; *pdVd = dVgsteff_dVd; *pdVb = dVgsteff_dVb;
}

/* Phase E: intrinsic terminal charges and trans-capacitances (b4ld.c:3014-3913).
 * Returns 0 when charges are not computed (xpart<0 or no charge computation). */
template <unsigned VK>
NGB_HD int b4_charges(const B4Ctx *c, size_t t, const double *Mrow, const double *Prow, int b4dr,
                      int ChargeComputationNeeded, B4W *w)
{
    const double xpart = B4M(xpart);
    const int capMod = B4SEL(capMod);
    const double nf = B4I(nf);
    const double coxe = B4M(coxe);
    const double phi = B4P(phi), k1ox = B4P(k1ox);
    const double Vds = w->Vds, Vbs = w->Vbs;
    const double Vgs_eff = w->Vgs_eff, dVgs_eff_dVg = w->dVgs_eff_dVg;
    const double Vbseff = w->Vbseff, dVbseff_dVb = w->dVbseff_dVb;
    const double Phis = w->Phis, dPhis_dVb = w->dPhis_dVb, sqrtPhis = w->sqrtPhis;
    const double dsqrtPhis_dVb = w->dsqrtPhis_dVb;
    const double Abulk0_Q = w->Abulk0_Q, dAbulk0_Q_dVb = w->dAbulk0_Q_dVb;
    double qgate, qbulk, qdrn, qsrc;
    double cggb, cgsb, cgdb, cdgb, cdsb, cddb, cbgb, cbsb, cbdb;
    double T0, T1, T2, T3, T4, T5, T6, T7, T8, T9, T10, T11, T12, tmp, tmp1;

    if ((xpart < 0) || (!ChargeComputationNeeded)) {
        w->qgate = w->qdrn = w->qsrc = w->qbulk = 0.0;
        w->cggb = w->cgsb = w->cgdb = 0.0;
        w->cdgb = w->cdsb = w->cddb = 0.0;
        w->cbgb = w->cbsb = w->cbdb = 0.0;
        return 0;
    }

    const double CoxWL = coxe * B4P(weffCV) * B4P(leffCV) * nf;

    if (capMod == 0) {
        double VbseffCV, dVbseffCV_dVb, Vfb, Vth, Vgst, dVth_dVb, Arg1;
        if (Vbseff < 0.0) { VbseffCV = Vbs; dVbseffCV_dVb = 1.0; }
        else { VbseffCV = phi - Phis; dVbseffCV_dVb = -dPhis_dVb * dVbseff_dVb; }

        Vfb = B4P(vfbcv);
        Vth = Vfb + phi + k1ox * sqrtPhis;
        Vgst = Vgs_eff - Vth;
        dVth_dVb = k1ox * dsqrtPhis_dVb * dVbseff_dVb;
        Arg1 = Vgs_eff - VbseffCV - Vfb;

        if (Arg1 <= 0.0) {
            qgate = CoxWL * Arg1;
            qbulk = -qgate;
            qdrn = 0.0;
            cggb = CoxWL * dVgs_eff_dVg;
            cgdb = 0.0;
            cgsb = CoxWL * (dVbseffCV_dVb - dVgs_eff_dVg);
            cdgb = 0.0; cddb = 0.0; cdsb = 0.0;
            cbgb = -CoxWL * dVgs_eff_dVg;
            cbdb = 0.0;
            cbsb = -cgsb;
        } else if (Vgst <= 0.0) {
            T1 = 0.5 * k1ox;
            T2 = sqrt(T1 * T1 + Arg1);
            qgate = CoxWL * k1ox * (T2 - T1);
            qbulk = -qgate;
            qdrn = 0.0;
            T0 = CoxWL * T1 / T2;
            cggb = T0 * dVgs_eff_dVg;
            cgdb = 0.0;
            cgsb = T0 * (dVbseffCV_dVb - dVgs_eff_dVg);
            cdgb = 0.0; cddb = 0.0; cdsb = 0.0;
            cbgb = -cggb;
            cbdb = 0.0;
            cbsb = -cgsb;
        } else {
            const double One_Third_CoxWL = CoxWL / 3.0;
            const double Two_Third_CoxWL = 2.0 * One_Third_CoxWL;
            const double AbulkCV = Abulk0_Q * B4P(abulkCVfactor);
            const double dAbulkCV_dVb = B4P(abulkCVfactor) * dAbulk0_Q_dVb * dVbseff_dVb;
            const double dVdsat_dVg = 1.0 / AbulkCV;
            const double Vdsat = Vgst * dVdsat_dVg;
            const double dVdsat_dVb = -(Vdsat * dAbulkCV_dVb + dVth_dVb) * dVdsat_dVg;
            double Alphaz, dAlphaz_dVg, dAlphaz_dVb;

            if (xpart > 0.5) {
                /* 0/100 partition */
                if (Vdsat <= Vds) {
                    T1 = Vdsat / 3.0;
                    qgate = CoxWL * (Vgs_eff - Vfb - phi - T1);
                    T2 = -Two_Third_CoxWL * Vgst;
                    qbulk = -(qgate + T2);
                    qdrn = 0.0;
                    cggb = One_Third_CoxWL * (3.0 - dVdsat_dVg) * dVgs_eff_dVg;
                    T2 = -One_Third_CoxWL * dVdsat_dVb;
                    cgsb = -(cggb + T2);
                    cgdb = 0.0;
                    cdgb = 0.0; cddb = 0.0; cdsb = 0.0;
                    cbgb = -(cggb - Two_Third_CoxWL * dVgs_eff_dVg);
                    T3 = -(T2 + Two_Third_CoxWL * dVth_dVb);
                    cbsb = -(cbgb + T3);
                    cbdb = 0.0;
                } else {
                    Alphaz = Vgst / Vdsat;
                    T1 = 2.0 * Vdsat - Vds;
                    T2 = Vds / (3.0 * T1);
                    T3 = T2 * Vds;
                    T9 = 0.25 * CoxWL;
                    T4 = T9 * Alphaz;
                    T7 = 2.0 * Vds - T1 - 3.0 * T3;
                    T8 = T3 - T1 - 2.0 * Vds;
                    qgate = CoxWL * (Vgs_eff - Vfb - phi - 0.5 * (Vds - T3));
                    T10 = T4 * T8;
                    qdrn = T4 * T7;
                    qbulk = -(qgate + qdrn + T10);

                    T5 = T3 / T1;
                    cggb = CoxWL * (1.0 - T5 * dVdsat_dVg) * dVgs_eff_dVg;
                    T11 = -CoxWL * T5 * dVdsat_dVb;
                    cgdb = CoxWL * (T2 - 0.5 + 0.5 * T5);
                    cgsb = -(cggb + T11 + cgdb);
                    T6 = 1.0 / Vdsat;
                    dAlphaz_dVg = T6 * (1.0 - Alphaz * dVdsat_dVg);
                    dAlphaz_dVb = -T6 * (dVth_dVb + Alphaz * dVdsat_dVb);
                    T7 = T9 * T7;
                    T8 = T9 * T8;
                    T9 = 2.0 * T4 * (1.0 - 3.0 * T5);
                    cdgb = (T7 * dAlphaz_dVg - T9 * dVdsat_dVg) * dVgs_eff_dVg;
                    T12 = T7 * dAlphaz_dVb - T9 * dVdsat_dVb;
                    cddb = T4 * (3.0 - 6.0 * T2 - 3.0 * T5);
                    cdsb = -(cdgb + T12 + cddb);

                    T9 = 2.0 * T4 * (1.0 + T5);
                    T10 = (T8 * dAlphaz_dVg - T9 * dVdsat_dVg) * dVgs_eff_dVg;
                    T11 = T8 * dAlphaz_dVb - T9 * dVdsat_dVb;
                    T12 = T4 * (2.0 * T2 + T5 - 1.0);
                    T0 = -(T10 + T11 + T12);

                    cbgb = -(cggb + cdgb + T10);
                    cbdb = -(cgdb + cddb + T12);
                    cbsb = -(cgsb + cdsb + T0);
                }
            } else if (xpart < 0.5) {
                /* 40/60 partition */
                if (Vds >= Vdsat) {
                    T1 = Vdsat / 3.0;
                    qgate = CoxWL * (Vgs_eff - Vfb - phi - T1);
                    T2 = -Two_Third_CoxWL * Vgst;
                    qbulk = -(qgate + T2);
                    qdrn = 0.4 * T2;
                    cggb = One_Third_CoxWL * (3.0 - dVdsat_dVg) * dVgs_eff_dVg;
                    T2 = -One_Third_CoxWL * dVdsat_dVb;
                    cgsb = -(cggb + T2);
                    cgdb = 0.0;
                    T3 = 0.4 * Two_Third_CoxWL;
                    cdgb = -T3 * dVgs_eff_dVg;
                    cddb = 0.0;
                    T4 = T3 * dVth_dVb;
                    cdsb = -(T4 + cdgb);
                    cbgb = -(cggb - Two_Third_CoxWL * dVgs_eff_dVg);
                    T3 = -(T2 + Two_Third_CoxWL * dVth_dVb);
                    cbsb = -(cbgb + T3);
                    cbdb = 0.0;
                } else {
                    Alphaz = Vgst / Vdsat;
                    T1 = 2.0 * Vdsat - Vds;
                    T2 = Vds / (3.0 * T1);
                    T3 = T2 * Vds;
                    T9 = 0.25 * CoxWL;
                    T4 = T9 * Alphaz;
                    qgate = CoxWL * (Vgs_eff - Vfb - phi - 0.5 * (Vds - T3));

                    T5 = T3 / T1;
                    cggb = CoxWL * (1.0 - T5 * dVdsat_dVg) * dVgs_eff_dVg;
                    tmp = -CoxWL * T5 * dVdsat_dVb;
                    cgdb = CoxWL * (T2 - 0.5 + 0.5 * T5);
                    cgsb = -(cggb + cgdb + tmp);

                    T6 = 1.0 / Vdsat;
                    dAlphaz_dVg = T6 * (1.0 - Alphaz * dVdsat_dVg);
                    dAlphaz_dVb = -T6 * (dVth_dVb + Alphaz * dVdsat_dVb);

                    T6 = 8.0 * Vdsat * Vdsat - 6.0 * Vdsat * Vds + 1.2 * Vds * Vds;
                    T8 = T2 / T1;
                    T7 = Vds - T1 - T8 * T6;
                    qdrn = T4 * T7;
                    T7 *= T9;
                    tmp = T8 / T1;
                    tmp1 = T4 * (2.0 - 4.0 * tmp * T6 + T8 * (16.0 * Vdsat - 6.0 * Vds));

                    cdgb = (T7 * dAlphaz_dVg - tmp1 * dVdsat_dVg) * dVgs_eff_dVg;
                    T10 = T7 * dAlphaz_dVb - tmp1 * dVdsat_dVb;
                    cddb = T4 * (2.0 - (1.0 / (3.0 * T1 * T1) + 2.0 * tmp) * T6
                                 + T8 * (6.0 * Vdsat - 2.4 * Vds));
                    cdsb = -(cdgb + T10 + cddb);

                    T7 = 2.0 * (T1 + T3);
                    qbulk = -(qgate - T4 * T7);
                    T7 *= T9;
                    T0 = 4.0 * T4 * (1.0 - T5);
                    T12 = (-T7 * dAlphaz_dVg - T0 * dVdsat_dVg) * dVgs_eff_dVg - cdgb;
                    T11 = -T7 * dAlphaz_dVb - T10 - T0 * dVdsat_dVb;
                    T10 = -4.0 * T4 * (T2 - 0.5 + 0.5 * T5) - cddb;
                    tmp = -(T10 + T11 + T12);

                    cbgb = -(cggb + cdgb + T12);
                    cbdb = -(cgdb + cddb + T10);
                    cbsb = -(cgsb + cdsb + tmp);
                }
            } else {
                /* 50/50 partition */
                if (Vds >= Vdsat) {
                    T1 = Vdsat / 3.0;
                    qgate = CoxWL * (Vgs_eff - Vfb - phi - T1);
                    T2 = -Two_Third_CoxWL * Vgst;
                    qbulk = -(qgate + T2);
                    qdrn = 0.5 * T2;
                    cggb = One_Third_CoxWL * (3.0 - dVdsat_dVg) * dVgs_eff_dVg;
                    T2 = -One_Third_CoxWL * dVdsat_dVb;
                    cgsb = -(cggb + T2);
                    cgdb = 0.0;
                    cdgb = -One_Third_CoxWL * dVgs_eff_dVg;
                    cddb = 0.0;
                    T4 = One_Third_CoxWL * dVth_dVb;
                    cdsb = -(T4 + cdgb);
                    cbgb = -(cggb - Two_Third_CoxWL * dVgs_eff_dVg);
                    T3 = -(T2 + Two_Third_CoxWL * dVth_dVb);
                    cbsb = -(cbgb + T3);
                    cbdb = 0.0;
                } else {
                    Alphaz = Vgst / Vdsat;
                    T1 = 2.0 * Vdsat - Vds;
                    T2 = Vds / (3.0 * T1);
                    T3 = T2 * Vds;
                    T9 = 0.25 * CoxWL;
                    T4 = T9 * Alphaz;
                    qgate = CoxWL * (Vgs_eff - Vfb - phi - 0.5 * (Vds - T3));

                    T5 = T3 / T1;
                    cggb = CoxWL * (1.0 - T5 * dVdsat_dVg) * dVgs_eff_dVg;
                    tmp = -CoxWL * T5 * dVdsat_dVb;
                    cgdb = CoxWL * (T2 - 0.5 + 0.5 * T5);
                    cgsb = -(cggb + cgdb + tmp);

                    T6 = 1.0 / Vdsat;
                    dAlphaz_dVg = T6 * (1.0 - Alphaz * dVdsat_dVg);
                    dAlphaz_dVb = -T6 * (dVth_dVb + Alphaz * dVdsat_dVb);

                    T7 = T1 + T3;
                    qdrn = -T4 * T7;
                    qbulk = -(qgate + qdrn + qdrn);
                    T7 *= T9;
                    T0 = T4 * (2.0 * T5 - 2.0);

                    cdgb = (T0 * dVdsat_dVg - T7 * dAlphaz_dVg) * dVgs_eff_dVg;
                    T12 = T0 * dVdsat_dVb - T7 * dAlphaz_dVb;
                    cddb = T4 * (1.0 - 2.0 * T2 - T5);
                    cdsb = -(cdgb + T12 + cddb);

                    cbgb = -(cggb + 2.0 * cdgb);
                    cbdb = -(cgdb + 2.0 * cddb);
                    cbsb = -(cgsb + 2.0 * cdsb);
                }
            }
        }
    } else {
        double VbseffCV, dVbseffCV_dVb, Vgsteff, dVgsteff_dVg, dVgsteff_dVd, dVgsteff_dVb;
        double Vfbeff, dVfbeff_dVg, dVfbeff_dVb, V3, Qac0, dQac0_dVg, dQac0_dVb;
        double Qsub0, dQsub0_dVg, dQsub0_dVd, dQsub0_dVb, AbulkCV, dAbulkCV_dVb, VdsatCV;
        double dT0_dVg, dT0_dVb, dT1_dVg, dT1_dVd, dT1_dVb;
        double VdseffCV, dVdseffCV_dVg, dVdseffCV_dVd, dVdseffCV_dVb;
        double Cgg1, Cgd1, Cgb1, Cbg1, Cbd1, Cbb1, Csg, Csd, Csb, Cgg, Cgd, Cgb, Cbg, Cbd, Cbb;
        const double vfbzb = B4I(vfbzb);

        if (Vbseff < 0.0) { VbseffCV = Vbseff; dVbseffCV_dVb = 1.0; }
        else { VbseffCV = phi - Phis; dVbseffCV_dVb = -dPhis_dVb; }

        b4_vgsteff_cv<VK>(c, Mrow, Prow, b4dr, w, &Vgsteff, &dVgsteff_dVg, &dVgsteff_dVd, &dVgsteff_dVb);

        if (capMod == 1) {
            const double Vfb = vfbzb;
            V3 = Vfb - Vgs_eff + VbseffCV - B4_DELTA_3;
            if (Vfb <= 0.0) T0 = sqrt(V3 * V3 - 4.0 * B4_DELTA_3 * Vfb);
            else T0 = sqrt(V3 * V3 + 4.0 * B4_DELTA_3 * Vfb);

            T1 = 0.5 * (1.0 + V3 / T0);
            Vfbeff = Vfb - 0.5 * (V3 + T0);
            dVfbeff_dVg = T1 * dVgs_eff_dVg;
            dVfbeff_dVb = -T1 * dVbseffCV_dVb;
            Qac0 = CoxWL * (Vfbeff - Vfb);
            dQac0_dVg = CoxWL * dVfbeff_dVg;
            dQac0_dVb = CoxWL * dVfbeff_dVb;

            T0 = 0.5 * k1ox;
            T3 = Vgs_eff - Vfbeff - VbseffCV - Vgsteff;
            if (k1ox == 0.0) { T1 = 0.0; T2 = 0.0; }
            else if (T3 < 0.0) { T1 = T0 + T3 / k1ox; T2 = CoxWL; }
            else { T1 = sqrt(T0 * T0 + T3); T2 = CoxWL * T0 / T1; }

            Qsub0 = CoxWL * k1ox * (T1 - T0);
            dQsub0_dVg = T2 * (dVgs_eff_dVg - dVfbeff_dVg - dVgsteff_dVg);
            dQsub0_dVd = -T2 * dVgsteff_dVd;
            dQsub0_dVb = -T2 * (dVfbeff_dVb + dVbseffCV_dVb + dVgsteff_dVb);

            AbulkCV = Abulk0_Q * B4P(abulkCVfactor);
            dAbulkCV_dVb = B4P(abulkCVfactor) * dAbulk0_Q_dVb;
            VdsatCV = Vgsteff / AbulkCV;

            T0 = VdsatCV - Vds - B4_DELTA_4;
            dT0_dVg = 1.0 / AbulkCV;
            dT0_dVb = -VdsatCV * dAbulkCV_dVb / AbulkCV;
            T1 = sqrt(T0 * T0 + 4.0 * B4_DELTA_4 * VdsatCV);
            dT1_dVg = (T0 + B4_DELTA_4 + B4_DELTA_4) / T1;
            dT1_dVd = -T0 / T1;
            dT1_dVb = dT1_dVg * dT0_dVb;
            dT1_dVg *= dT0_dVg;
            if (T0 >= 0.0) {
                VdseffCV = VdsatCV - 0.5 * (T0 + T1);
                dVdseffCV_dVg = 0.5 * (dT0_dVg - dT1_dVg);
                dVdseffCV_dVd = 0.5 * (1.0 - dT1_dVd);
                dVdseffCV_dVb = 0.5 * (dT0_dVb - dT1_dVb);
            } else {
                T3 = (B4_DELTA_4 + B4_DELTA_4) / (T1 - T0);
                T4 = 1.0 - T3;
                T5 = VdsatCV * T3 / (T1 - T0);
                VdseffCV = VdsatCV * T4;
                dVdseffCV_dVg = dT0_dVg * T4 + T5 * (dT1_dVg - dT0_dVg);
                dVdseffCV_dVd = T5 * (dT1_dVd + 1.0);
                dVdseffCV_dVb = dT0_dVb * (T4 - T5) + T5 * dT1_dVb;
            }
            if (Vds == 0.0) {
                VdseffCV = 0.0;
                dVdseffCV_dVg = 0.0;
                dVdseffCV_dVb = 0.0;
            }

            T0 = AbulkCV * VdseffCV;
            T1 = 12.0 * (Vgsteff - 0.5 * T0 + 1.0e-20);
            T2 = VdseffCV / T1;
            T3 = T0 * T2;

            T4 = (1.0 - 12.0 * T2 * T2 * AbulkCV);
            T5 = (6.0 * T0 * (4.0 * Vgsteff - T0) / (T1 * T1) - 0.5);
            T6 = 12.0 * T2 * T2 * Vgsteff;

            qgate = CoxWL * (Vgsteff - 0.5 * VdseffCV + T3);
            Cgg1 = CoxWL * (T4 + T5 * dVdseffCV_dVg);
            Cgd1 = CoxWL * T5 * dVdseffCV_dVd + Cgg1 * dVgsteff_dVd;
            Cgb1 = CoxWL * (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Cgg1 * dVgsteff_dVb;
            Cgg1 *= dVgsteff_dVg;

            T7 = 1.0 - AbulkCV;
            qbulk = CoxWL * T7 * (0.5 * VdseffCV - T3);
            T4 = -T7 * (T4 - 1.0);
            T5 = -T7 * T5;
            T6 = -(T7 * T6 + (0.5 * VdseffCV - T3));
            Cbg1 = CoxWL * (T4 + T5 * dVdseffCV_dVg);
            Cbd1 = CoxWL * T5 * dVdseffCV_dVd + Cbg1 * dVgsteff_dVd;
            Cbb1 = CoxWL * (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Cbg1 * dVgsteff_dVb;
            Cbg1 *= dVgsteff_dVg;

            if (xpart > 0.5) {
                T1 = T1 + T1;
                qsrc = -CoxWL * (0.5 * Vgsteff + 0.25 * T0 - T0 * T0 / T1);
                T7 = (4.0 * Vgsteff - T0) / (T1 * T1);
                T4 = -(0.5 + 24.0 * T0 * T0 / (T1 * T1));
                T5 = -(0.25 * AbulkCV - 12.0 * AbulkCV * T0 * T7);
                T6 = -(0.25 * VdseffCV - 12.0 * T0 * VdseffCV * T7);
                Csg = CoxWL * (T4 + T5 * dVdseffCV_dVg);
                Csd = CoxWL * T5 * dVdseffCV_dVd + Csg * dVgsteff_dVd;
                Csb = CoxWL * (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Csg * dVgsteff_dVb;
                Csg *= dVgsteff_dVg;
            } else if (xpart < 0.5) {
                T1 = T1 / 12.0;
                T2 = 0.5 * CoxWL / (T1 * T1);
                T3 = Vgsteff * (2.0 * T0 * T0 / 3.0 + Vgsteff * (Vgsteff - 4.0 * T0 / 3.0))
                   - 2.0 * T0 * T0 * T0 / 15.0;
                qsrc = -T2 * T3;
                T7 = 4.0 / 3.0 * Vgsteff * (Vgsteff - T0) + 0.4 * T0 * T0;
                T4 = -2.0 * qsrc / T1 - T2 * (Vgsteff * (3.0 * Vgsteff - 8.0 * T0 / 3.0)
                                              + 2.0 * T0 * T0 / 3.0);
                T5 = (qsrc / T1 + T2 * T7) * AbulkCV;
                T6 = (qsrc / T1 * VdseffCV + T2 * T7 * VdseffCV);
                Csg = (T4 + T5 * dVdseffCV_dVg);
                Csd = T5 * dVdseffCV_dVd + Csg * dVgsteff_dVd;
                Csb = (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Csg * dVgsteff_dVb;
                Csg *= dVgsteff_dVg;
            } else {
                qsrc = -0.5 * (qgate + qbulk);
                Csg = -0.5 * (Cgg1 + Cbg1);
                Csb = -0.5 * (Cgb1 + Cbb1);
                Csd = -0.5 * (Cgd1 + Cbd1);
            }

            qgate += Qac0 + Qsub0;
            qbulk -= (Qac0 + Qsub0);
            qdrn = -(qgate + qbulk + qsrc);

            Cgg = dQac0_dVg + dQsub0_dVg + Cgg1;
            Cgd = dQsub0_dVd + Cgd1;
            Cgb = dQac0_dVb + dQsub0_dVb + Cgb1;

            Cbg = Cbg1 - dQac0_dVg - dQsub0_dVg;
            Cbd = Cbd1 - dQsub0_dVd;
            Cbb = Cbb1 - dQac0_dVb - dQsub0_dVb;

            Cgb *= dVbseff_dVb;
            Cbb *= dVbseff_dVb;
            Csb *= dVbseff_dVb;
        } else {
            /* capMod == 2: charge-thickness model */
            const double Vtm = w->Vtm, epssub = w->epssub;
            const double toxp = B4I(toxp), Cox = B4I(coxp);
            const double ldeb = B4P(ldeb), sqrtPhi = B4P(sqrtPhi);
            double Tox, Tcen, dTcen_dVg, dTcen_dVd, dTcen_dVb, LINK, V4, Ccen, Coxeff;
            double dCoxeff_dVg, dCoxeff_dVd, dCoxeff_dVb, CoxWLcen, QovCox;
            double Denomi, DeltaPhi, dDeltaPhi_dVg, VgDP, dVgDP_dVg;

            V3 = vfbzb - Vgs_eff + VbseffCV - B4_DELTA_3;
            if (vfbzb <= 0.0) T0 = sqrt(V3 * V3 - 4.0 * B4_DELTA_3 * vfbzb);
            else T0 = sqrt(V3 * V3 + 4.0 * B4_DELTA_3 * vfbzb);

            T1 = 0.5 * (1.0 + V3 / T0);
            Vfbeff = vfbzb - 0.5 * (V3 + T0);
            dVfbeff_dVg = T1 * dVgs_eff_dVg;
            dVfbeff_dVb = -T1 * dVbseffCV_dVb;

            Tox = 1.0e8 * toxp;
            T0 = (Vgs_eff - VbseffCV - vfbzb) / Tox;
            dT0_dVg = dVgs_eff_dVg / Tox;
            dT0_dVb = -dVbseffCV_dVb / Tox;

            tmp = T0 * B4P(acde);
            if ((-B4_EXP_THRESHOLD < tmp) && (tmp < B4_EXP_THRESHOLD)) {
                Tcen = ldeb * ngb_exp(tmp);
                dTcen_dVg = B4P(acde) * Tcen;
                dTcen_dVb = dTcen_dVg * dT0_dVb;
                dTcen_dVg *= dT0_dVg;
            } else if (tmp <= -B4_EXP_THRESHOLD) {
                Tcen = ldeb * B4_MIN_EXP;
                dTcen_dVg = dTcen_dVb = 0.0;
            } else {
                Tcen = ldeb * B4_MAX_EXP;
                dTcen_dVg = dTcen_dVb = 0.0;
            }

            LINK = 1.0e-3 * toxp;
            V3 = ldeb - Tcen - LINK;
            V4 = sqrt(V3 * V3 + 4.0 * LINK * ldeb);
            Tcen = ldeb - 0.5 * (V3 + V4);
            T1 = 0.5 * (1.0 + V3 / V4);
            dTcen_dVg *= T1;
            dTcen_dVb *= T1;

            Ccen = epssub / Tcen;
            T2 = Cox / (Cox + Ccen);
            Coxeff = T2 * Ccen;
            T3 = -Ccen / Tcen;
            dCoxeff_dVg = T2 * T2 * T3;
            dCoxeff_dVb = dCoxeff_dVg * dTcen_dVb;
            dCoxeff_dVg *= dTcen_dVg;
            CoxWLcen = CoxWL * Coxeff / coxe;

            Qac0 = CoxWLcen * (Vfbeff - vfbzb);
            QovCox = Qac0 / Coxeff;
            dQac0_dVg = CoxWLcen * dVfbeff_dVg + QovCox * dCoxeff_dVg;
            dQac0_dVb = CoxWLcen * dVfbeff_dVb + QovCox * dCoxeff_dVb;

            T0 = 0.5 * k1ox;
            T3 = Vgs_eff - Vfbeff - VbseffCV - Vgsteff;
            if (k1ox == 0.0) { T1 = 0.0; T2 = 0.0; }
            else if (T3 < 0.0) { T1 = T0 + T3 / k1ox; T2 = CoxWLcen; }
            else { T1 = sqrt(T0 * T0 + T3); T2 = CoxWLcen * T0 / T1; }

            Qsub0 = CoxWLcen * k1ox * (T1 - T0);
            QovCox = Qsub0 / Coxeff;
            dQsub0_dVg = T2 * (dVgs_eff_dVg - dVfbeff_dVg - dVgsteff_dVg) + QovCox * dCoxeff_dVg;
            dQsub0_dVd = -T2 * dVgsteff_dVd;
            dQsub0_dVb = -T2 * (dVfbeff_dVb + dVbseffCV_dVb + dVgsteff_dVb) + QovCox * dCoxeff_dVb;

            /* gate-bias dependent delta Phis */
            if (k1ox <= 0.0) {
                Denomi = 0.25 * B4P(moin) * Vtm;
                T0 = 0.5 * sqrtPhi;
            } else {
                Denomi = B4P(moin) * Vtm * k1ox * k1ox;
                T0 = k1ox * sqrtPhi;
            }
            T1 = 2.0 * T0 + Vgsteff;

            DeltaPhi = Vtm * ngb_log(1.0 + T1 * Vgsteff / Denomi);
            dDeltaPhi_dVg = 2.0 * Vtm * (T1 - T0) / (Denomi + T1 * Vgsteff);

            /* VgDP = Vgsteff - DeltaPhi */
            T0 = Vgsteff - DeltaPhi - 0.001;
            dT0_dVg = 1.0 - dDeltaPhi_dVg;
            T1 = sqrt(T0 * T0 + Vgsteff * 0.004);
            VgDP = 0.5 * (T0 + T1);
            dVgDP_dVg = 0.5 * (dT0_dVg + (T0 * dT0_dVg + 0.002) / T1);

            Tox += Tox;
            T0 = (Vgsteff + B4I(vtfbphi2)) / Tox;
            tmp = ngb_exp(B4M(bdos) * 0.7 * ngb_log(T0));
            T1 = 1.0 + tmp;
            T2 = B4M(bdos) * 0.7 * tmp / (T0 * Tox);
            Tcen = B4M(ados) * 1.9e-9 / T1;
            dTcen_dVg = -Tcen * T2 / T1;
            dTcen_dVd = dTcen_dVg * dVgsteff_dVd;
            dTcen_dVb = dTcen_dVg * dVgsteff_dVb;
            dTcen_dVg *= dVgsteff_dVg;

            Ccen = epssub / Tcen;
            T0 = Cox / (Cox + Ccen);
            Coxeff = T0 * Ccen;
            T1 = -Ccen / Tcen;
            dCoxeff_dVg = T0 * T0 * T1;
            dCoxeff_dVd = dCoxeff_dVg * dTcen_dVd;
            dCoxeff_dVb = dCoxeff_dVg * dTcen_dVb;
            dCoxeff_dVg *= dTcen_dVg;
            CoxWLcen = CoxWL * Coxeff / coxe;

            AbulkCV = Abulk0_Q * B4P(abulkCVfactor);
            dAbulkCV_dVb = B4P(abulkCVfactor) * dAbulk0_Q_dVb;
            VdsatCV = VgDP / AbulkCV;

            T0 = VdsatCV - Vds - B4_DELTA_4;
            dT0_dVg = dVgDP_dVg / AbulkCV;
            dT0_dVb = -VdsatCV * dAbulkCV_dVb / AbulkCV;
            T1 = sqrt(T0 * T0 + 4.0 * B4_DELTA_4 * VdsatCV);
            dT1_dVg = (T0 + B4_DELTA_4 + B4_DELTA_4) / T1;
            dT1_dVd = -T0 / T1;
            dT1_dVb = dT1_dVg * dT0_dVb;
            dT1_dVg *= dT0_dVg;
            if (T0 >= 0.0) {
                VdseffCV = VdsatCV - 0.5 * (T0 + T1);
                dVdseffCV_dVg = 0.5 * (dT0_dVg - dT1_dVg);
                dVdseffCV_dVd = 0.5 * (1.0 - dT1_dVd);
                dVdseffCV_dVb = 0.5 * (dT0_dVb - dT1_dVb);
            } else {
                T3 = (B4_DELTA_4 + B4_DELTA_4) / (T1 - T0);
                T4 = 1.0 - T3;
                T5 = VdsatCV * T3 / (T1 - T0);
                VdseffCV = VdsatCV * T4;
                dVdseffCV_dVg = dT0_dVg * T4 + T5 * (dT1_dVg - dT0_dVg);
                dVdseffCV_dVd = T5 * (dT1_dVd + 1.0);
                dVdseffCV_dVb = dT0_dVb * (T4 - T5) + T5 * dT1_dVb;
            }
            if (Vds == 0.0) {
                VdseffCV = 0.0;
                dVdseffCV_dVg = 0.0;
                dVdseffCV_dVb = 0.0;
            }

            T0 = AbulkCV * VdseffCV;
            T1 = VgDP;
            T2 = 12.0 * (T1 - 0.5 * T0 + 1.0e-20);
            T3 = T0 / T2;
            T4 = 1.0 - 12.0 * T3 * T3;
            T5 = AbulkCV * (6.0 * T0 * (4.0 * T1 - T0) / (T2 * T2) - 0.5);
            T6 = T5 * VdseffCV / AbulkCV;

            qgate = CoxWLcen * (T1 - T0 * (0.5 - T3));
            QovCox = qgate / Coxeff;
            Cgg1 = CoxWLcen * (T4 * dVgDP_dVg + T5 * dVdseffCV_dVg);
            Cgd1 = CoxWLcen * T5 * dVdseffCV_dVd + Cgg1 * dVgsteff_dVd + QovCox * dCoxeff_dVd;
            Cgb1 = CoxWLcen * (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Cgg1 * dVgsteff_dVb
                 + QovCox * dCoxeff_dVb;
            Cgg1 = Cgg1 * dVgsteff_dVg + QovCox * dCoxeff_dVg;

            T7 = 1.0 - AbulkCV;
            T8 = T2 * T2;
            T9 = 12.0 * T7 * T0 * T0 / (T8 * AbulkCV);
            T10 = T9 * dVgDP_dVg;
            T11 = -T7 * T5 / AbulkCV;
            T12 = -(T9 * T1 / AbulkCV + VdseffCV * (0.5 - T0 / T2));

            qbulk = CoxWLcen * T7 * (0.5 * VdseffCV - T0 * VdseffCV / T2);
            QovCox = qbulk / Coxeff;
            Cbg1 = CoxWLcen * (T10 + T11 * dVdseffCV_dVg);
            Cbd1 = CoxWLcen * T11 * dVdseffCV_dVd + Cbg1 * dVgsteff_dVd + QovCox * dCoxeff_dVd;
            Cbb1 = CoxWLcen * (T11 * dVdseffCV_dVb + T12 * dAbulkCV_dVb) + Cbg1 * dVgsteff_dVb
                 + QovCox * dCoxeff_dVb;
            Cbg1 = Cbg1 * dVgsteff_dVg + QovCox * dCoxeff_dVg;

            if (xpart > 0.5) {
                qsrc = -CoxWLcen * (T1 / 2.0 + T0 / 4.0 - 0.5 * T0 * T0 / T2);
                QovCox = qsrc / Coxeff;
                T2 += T2;
                T3 = T2 * T2;
                T7 = -(0.25 - 12.0 * T0 * (4.0 * T1 - T0) / T3);
                T4 = -(0.5 + 24.0 * T0 * T0 / T3) * dVgDP_dVg;
                T5 = T7 * AbulkCV;
                T6 = T7 * VdseffCV;

                Csg = CoxWLcen * (T4 + T5 * dVdseffCV_dVg);
                Csd = CoxWLcen * T5 * dVdseffCV_dVd + Csg * dVgsteff_dVd + QovCox * dCoxeff_dVd;
                Csb = CoxWLcen * (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Csg * dVgsteff_dVb
                    + QovCox * dCoxeff_dVb;
                Csg = Csg * dVgsteff_dVg + QovCox * dCoxeff_dVg;
            } else if (xpart < 0.5) {
                T2 = T2 / 12.0;
                T3 = 0.5 * CoxWLcen / (T2 * T2);
                T4 = T1 * (2.0 * T0 * T0 / 3.0 + T1 * (T1 - 4.0 * T0 / 3.0))
                   - 2.0 * T0 * T0 * T0 / 15.0;
                qsrc = -T3 * T4;
                QovCox = qsrc / Coxeff;
                T8 = 4.0 / 3.0 * T1 * (T1 - T0) + 0.4 * T0 * T0;
                T5 = -2.0 * qsrc / T2 - T3 * (T1 * (3.0 * T1 - 8.0 * T0 / 3.0) + 2.0 * T0 * T0 / 3.0);
                T6 = AbulkCV * (qsrc / T2 + T3 * T8);
                T7 = T6 * VdseffCV / AbulkCV;

                Csg = T5 * dVgDP_dVg + T6 * dVdseffCV_dVg;
                Csd = Csg * dVgsteff_dVd + T6 * dVdseffCV_dVd + QovCox * dCoxeff_dVd;
                Csb = Csg * dVgsteff_dVb + T6 * dVdseffCV_dVb + T7 * dAbulkCV_dVb
                    + QovCox * dCoxeff_dVb;
                Csg = Csg * dVgsteff_dVg + QovCox * dCoxeff_dVg;
            } else {
                qsrc = -0.5 * qgate;
                Csg = -0.5 * Cgg1;
                Csd = -0.5 * Cgd1;
                Csb = -0.5 * Cgb1;
            }

            qgate += Qac0 + Qsub0 - qbulk;
            qbulk -= (Qac0 + Qsub0);
            qdrn = -(qgate + qbulk + qsrc);

            Cbg = Cbg1 - dQac0_dVg - dQsub0_dVg;
            Cbd = Cbd1 - dQsub0_dVd;
            Cbb = Cbb1 - dQac0_dVb - dQsub0_dVb;

            Cgg = Cgg1 - Cbg;
            Cgd = Cgd1 - Cbd;
            Cgb = Cgb1 - Cbb;

            Cgb *= dVbseff_dVb;
            Cbb *= dVbseff_dVb;
            Csb *= dVbseff_dVb;
        }
        cggb = Cgg;
        cgsb = -(Cgg + Cgd + Cgb);
        cgdb = Cgd;
        cdgb = -(Cgg + Cbg + Csg);
        cdsb = (Cgg + Cgd + Cgb + Cbg + Cbd + Cbb + Csg + Csd + Csb);
        cddb = -(Cgd + Cbd + Csd);
        cbgb = Cbg;
        cbsb = -(Cbg + Cbd + Cbb);
        cbdb = Cbd;
    }

    w->qgate = qgate; w->qbulk = qbulk; w->qdrn = qdrn;
    w->qsrc = -(qgate + qbulk + qdrn);
    w->cggb = cggb; w->cgsb = cgsb; w->cgdb = cgdb;
    w->cdgb = cdgb; w->cdsb = cdsb; w->cddb = cddb;
    w->cbgb = cbgb; w->cbsb = cbsb; w->cbdb = cbdb;
    (void)qsrc;
    return 1;
}

/* one junction's depletion charge and capacitance (b4ld.c:3966-4013 source, :4016-4062 drain) */
NGB_HD_SHARED void b4_junction_cv(double vj, double cz, double czsw, double czswg, double MJ, double MJSW,
                           double MJSWG, double PhiB, double PhiBSW, double PhiBSWG,
                           double *q, double *cap)
{
    double arg, sarg, T0, T1;
    if (vj == 0.0) {
        *q = 0.0;
        *cap = cz + czsw + czswg;
    } else if (vj < 0.0) {
        if (cz > 0.0) {
            arg = 1.0 - vj / PhiB;
            if (MJ == 0.5) sarg = 1.0 / sqrt(arg);
            else sarg = ngb_exp(-MJ * ngb_log(arg));
            *q = PhiB * cz * (1.0 - arg * sarg) / (1.0 - MJ);
            *cap = cz * sarg;
        } else {
            *q = 0.0;
            *cap = 0.0;
        }
        if (czsw > 0.0) {
            arg = 1.0 - vj / PhiBSW;
            if (MJSW == 0.5) sarg = 1.0 / sqrt(arg);
            else sarg = ngb_exp(-MJSW * ngb_log(arg));
            *q += PhiBSW * czsw * (1.0 - arg * sarg) / (1.0 - MJSW);
            *cap += czsw * sarg;
        }
        if (czswg > 0.0) {
            arg = 1.0 - vj / PhiBSWG;
            if (MJSWG == 0.5) sarg = 1.0 / sqrt(arg);
            else sarg = ngb_exp(-MJSWG * ngb_log(arg));
            *q += PhiBSWG * czswg * (1.0 - arg * sarg) / (1.0 - MJSWG);
            *cap += czswg * sarg;
        }
    } else {
        T0 = cz + czsw + czswg;
        T1 = vj * (cz * MJ / PhiB + czsw * MJSW / PhiBSW + czswg * MJSWG / PhiBSWG);
        *q = vj * (T0 + 0.5 * T1);
        *cap = T0 + T1;
    }
}

/* store one stamp value if the position is live for this instance */
/* every (position, instance) owns a row whether the assembly reads it or not (ground nodes, absent internal
 * nodes), so a stamp is one store at a computed address: no row lookup, no branch.  NGB_B4_STAMP_LOOKUP restores
 * the lookup (dead rows are then not written). */
#ifdef NGB_B4_STAMP_LOOKUP
#define B4_STAMP(K, V) do { int r_ = NGB_LDG(&c->spos[(K) * c->ninst + inst]); \
                            if (r_ >= 0) c->stamp[(size_t)r_ * c->S + s] = (V); } while (0)
#else
#define B4_STAMP(K, V) c->stamp[(size_t)(c->srow0 + (K) * c->ninst + inst) * c->S + s] = (V)
#endif

/* The whole load for thread t = inst * S + s.  Returns NGB_OK or an NGB_E_* code. */
/* what every phase of one evaluation starts from (cheap to recompute, so the split kernels do) */
typedef struct B4Pro {
    int inst, s, head, mode_ckt, flags, charge;
    int dr;                    /* overlay: rows from the instance's sample-0 rows (Mrow / Prow) to the thread's own */
    const double *Mrow, *Prow;
} B4Pro;

/* returns 1 when the thread has work; 0 with *err set otherwise.  `first` also applies DCtran's deferred
 * whole-state copies, which must happen exactly once per load */
NGB_HD int b4_prologue(const B4Ctx *c, size_t t, int first, B4Pro *p, int *err)
{
    const int S = c->S;
    const int inst = (int)(t / (size_t)S);
    const int s = (int)(t - (size_t)inst * S);
    *err = NGB_OK;
    if (!NGB_LDG(&c->ctl.active[s])) return 0;

    const int mode_ckt = NGB_LDG(&c->ctl.mode[s]);
    const int head = NGB_LDG(&c->ctl.head[s]);
    const int prow = c->prow_per_thread ? NGB_LDG(&c->prow[t]) : NGB_LDG(&c->prow[inst]);
    p->inst = inst; p->s = s; p->head = head; p->mode_ckt = mode_ckt;
    p->flags = NGB_LDG(&c->flags[inst]);
    {
        const int prow0 = c->overlay ? NGB_LDG(&c->prow[t - (size_t)s]) : prow;
        p->dr = prow - prow0;
        p->Mrow = c->mtab + (size_t)prow0 * B4M_COUNT;
        p->Prow = c->ptab + (size_t)prow0 * B4P_COUNT;
    }

    if (mode_ckt & NGB_MODEINITSMSIG) { *err = NGB_E_UNSUPP; return 0; }

    /* deferred whole-state copies of DCtran: this thread owns its 29 states */
    if (first) {
        const int sop = NGB_LDG(&c->ctl.stateop[s]);
        if (sop) {
            B4ST_BASES(head);
            for (int k = 0; k < B4ST_COUNT; k++) {
                if (sop & NGB_OP_COPY01) B4ST(1, k) = B4ST(0, k);
                if (sop & NGB_OP_COPY1_23) { const double v = B4ST(1, k); B4ST(2, k) = v; if (c->ctl.nhist > 3) B4ST(3, k) = v; }
                if (sop & NGB_OP_COPY23) { const double v = B4ST(2, k); B4ST(0, k) = v; if (c->ctl.nhist > 3) B4ST(3, k) = v; }
            }
        }
    }

    p->charge =
        ((mode_ckt & (NGB_MODEDCTRANCURVE | NGB_MODEAC | NGB_MODETRAN | NGB_MODEINITSMSIG)) ||
         ((mode_ckt & NGB_MODETRANOP) && (mode_ckt & NGB_MODEUIC))) ? 1 : 0;
    return 1;
}

template <unsigned VK> NGB_HD int b4_finish(const B4Ctx *c, size_t t, const B4Pro *pro, const B4W *wp);

template <unsigned VK>
NGB_HD int b4_load_thread(const B4Ctx *c, size_t t)
{
    B4Pro p;
    B4W w;
    int err;
    if (!b4_prologue(c, t, 1, &p, &err)) return err;
#if defined(__CUDA_ARCH__) && defined(NGB_B4_PREFETCH)
    {   /* experiment switch: the per-thread columns this evaluation will read (instance parameters, the two newest
         * state planes) are asked for at once, so that the loads scattered over the 14 k instructions behind find them
         * in L2 instead of waiting for DRAM one after the other */
        B4ST_BASES(p.head);
#pragma unroll
        for (int k = 0; k < B4ST_COUNT; k++) {
            asm volatile("prefetch.global.L2 [%0];" :: "l"(&b4st0[(size_t)k * c->T]));
            asm volatile("prefetch.global.L2 [%0];" :: "l"(&b4st1[(size_t)k * c->T]));
        }
#pragma unroll
        for (int f = 0; f < B4I_COUNT; f++)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(&c->inst[(size_t)f * c->T + t]));
    }
#endif
    b4_fetch_limit<VK>(c, t, p.inst, p.s, p.head, p.mode_ckt, p.Mrow, p.dr, p.flags, &w);
    b4_core_dc<VK>(c, t, p.s, p.Mrow, p.Prow, p.dr, p.flags, &w);
    /* the parasitics and the intrinsic charges only read what the core phase left; the charges first (12 values for the
     * finish phase alive across the parasitics instead of 52 across the charges) was measured 6 % SLOWER on B200 */
    b4_parasitics<VK>(c, t, p.Mrow, p.Prow, p.dr, p.flags, &w);
    b4_charges<VK>(c, t, p.Mrow, p.Prow, p.dr, p.charge, &w);
    return b4_finish<VK>(c, t, &p, &w);
}

#define B4FIN_NAME b4_finish
#define B4FIN_W const B4W w = *wp;
#define BW(f) w.f
#include "bsim4_finish.inc"
#undef B4FIN_NAME
#undef B4FIN_W
#undef BW

#ifndef __CUDACC__
/* host build (tests/hostsim): the same dispatch the CUDA launcher does over its kernel instantiations */
static inline int b4_load_thread_variant(const B4Ctx *c, size_t t)
{
#define X(k) if (c->variant == (k)) return b4_load_thread<(k)>(c, t);
    NGB_B4_VARIANT_KEYS(X)
#undef X
    return b4_load_thread<NGB_B4_GENERIC>(c, t);
}
#endif
/* ---- BSIM4trunc out of the load ------------------------------------------------------------------
 * The reference calls DEVtrunc once per converged time point (CKTtrunc, dctran.c:794); evaluated inside every
 * load it was a fifth of the kernel's instructions (10 copies of CKTterr, 70 of the 313 divisions).  With
 * lte_deferred the load skips it and this thread runs after SMPsolve + NIconvTest, only for samples whose
 * iteration can be the converged one (transient, MODEINITFLOAT, CKTnoncon == 0, node test passed -- a superset
 * of NIiter's own decision, which the controller takes afterwards).  It reads the same states the load has just
 * written, so the bounds are the same bits.  On the device a warp takes ONE sample (lanes = instances): samples
 * converge at different steps, and with the load's sample-fastest threads nearly every warp would hold a
 * converged lane and walk the whole evaluation. */
/* can this sample's iteration be the converged one?  (uniform over the instances of a sample) */
NGB_HD int b4_lte_wanted(const B4Ctx *c, int s)
{
    if (!c->ctl.lte || !NGB_LDG(&c->ctl.active[s])) return 0;
    const int mode_ckt = NGB_LDG(&c->ctl.mode[s]);
    if (!(mode_ckt & NGB_MODETRAN) || !(mode_ckt & NGB_MODEINITFLOAT)) return 0;
    if (NGB_LDG(&c->ctl.noncon[s]) != 0 || NGB_LDG(&c->nodeconv[s]) != 0) return 0;
    return 1;
}
/* bounds of one instance, folded into *m1 / *m2 (the caller reduces over its instances and writes once) */
NGB_HD void b4_lte_inst(const B4Ctx *c, int inst, int s, double *m1, double *m2)
{
    const size_t t = (size_t)inst * c->S + s;
    const int head = NGB_LDG(&c->ctl.head[s]);
    const int order = NGB_LDG(&c->ctl.order[s]);
    const int flags = NGB_LDG(&c->flags[inst]);
    const int rbodyMod = B4F_RBODY(flags), rgateMod = B4F_RGATE(flags);
    double d1, d2;
    if (order != 1 && order != 2) return;
#define B4_LTE1(KQ) do { ngb_lte_values(&c->ctl, s, c->state, B4ST_COUNT, (size_t)c->T, t, head, KQ, order, &d1, &d2); \
                         if (d1 < *m1) { *m1 = d1; } if (d2 < *m2) { *m2 = d2; } } while (0)
    B4_LTE1(B4ST_qb);
    B4_LTE1(B4ST_qg);
    B4_LTE1(B4ST_qd);
    if (rbodyMod) { B4_LTE1(B4ST_qbs); B4_LTE1(B4ST_qbd); }
    if (rgateMod == 3) B4_LTE1(B4ST_qgmid);
#undef B4_LTE1
}

#endif /* __cplusplus */

#endif
