/* ngb_types.h -- operand blocks shared by the host code (C) and the kernels (CUDA).
 *
 * Layout rules (DESIGN.md "Data layout in HBM"):
 *   - a batch holds S independent samples (Monte-Carlo draws / sweep points) of ONE circuit
 *     topology; the sample index is always the fastest-varying one, so a warp that works on
 *     32 consecutive samples of the same instance/node issues fully coalesced accesses;
 *   - solution vectors  x[buf][eq][S]      (buf 0/1; xsel[s] says which one is CKTrhsOld);
 *   - device states     state[hist][k][T]  (T = instances * S), ring-rotated per sample by head[s];
 *   - stamp buffer      stamp[row][S]      one row per live (instance, stamp position);
 *   - matrices          Ax[S][nnz]         sample-major: one warp owns one sample's matrix.
 */
#ifndef NGB_TYPES_H
#define NGB_TYPES_H

#include "ngb_common.h"

#ifdef __cplusplus
extern "C" {
#endif

/* per-sample control block (the per-circuit scalars of CKTcircuit, cktdefs.h:66-331), SoA [S] */
typedef struct NgbCtl {
    int S;
    int *mode;            /* CKTmode                                                    */
    int *active;          /* 1: sample takes part in this Newton step                   */
    int *head;            /* ring position of CKTstate0 (CKTstates rotation, dctran.c:659) */
    int *order;           /* CKTorder                                                   */
    int *noncon;          /* CKTnoncon                                                  */
    int *xsel;            /* which x buffer is CKTrhsOld (pointer SWAP of niiter.c:362)  */
    int *err;             /* sticky NGB_E_* per sample                                  */
    double *ag0, *ag1;    /* CKTag[0..1]                                                */
    double *ag2;          /* CKTag[2]: GEAR order 2 (nicomcof.c:52-121)                  */
    double *delta;        /* CKTdelta                                                   */
    double *delta_old;    /* CKTdeltaOld[7][S]                                          */
    double *time;         /* CKTtime                                                    */
    double *gmin;         /* CKTgmin                                                    */
    double *diag_gmin;    /* CKTdiagGmin                                                */
    double *srcfact;      /* CKTsrcFact                                                 */
    /* local-truncation-error estimates written by the load kernels (CKTtrunc/CKTterr):
     * lte for the current CKTorder, lte2 for a trial with order 2 (dctran.c:794, 823) */
    double *lte, *lte2;
    int *stateop;         /* pending whole-state copies, applied by the next load (NGB_OP_*) */
    int *lusel;           /* which LU pattern set the sample is on: 0 = first pivoting factor (DC op),
                             1 = the re-pivoting of the first transient iteration (niiter.c:107-111)   */
    /* shared scalars */
    int nhist;            /* state vectors in the ring: CKTmaxOrder + 2 (cktsetup.c:192)  */
    int gear;             /* CKTintegrateMethod == GEAR                                   */
    double reltol, abstol, chgtol, trtol;
} NgbCtl;

/* deferred state copies of dctran.c (each device thread owns its own state slice) */
#define NGB_OP_COPY01   1   /* memcpy(CKTstate1, CKTstate0)           dctran.c:319-322 */
#define NGB_OP_COPY1_23 2   /* state2 = state1; state3 = state1       dctran.c:711-716 */
#define NGB_OP_COPY23   4   /* the same copy seen after the ring rotated once: state3 = state2 and state0 (the old state3) = state2 */

/* source waveform codes (vsrcdefs.h:146-160) */
#define NGB_FN_PULSE 1
#define NGB_FN_SINE  2
#define NGB_FN_EXP   3
#define NGB_FN_SFFM  4
#define NGB_FN_PWL   5
#define NGB_FN_AM    6

/* circuit-wide scalars */
typedef struct NgbOpts {
    double reltol, abstol, vntol, chgtol, trtol, temp, vt0, xmu;
    double tstep, tstop, tmax, tstart, delmin, minbreak, gmin;
    int method, maxorder, itl4, itl1, uic;
    int no_op_iter;                            /* CKTnoOpIter: skip the plain NIiter of CKTop, start with the fallbacks */
    int num_gmin_steps, num_src_steps, itl2;   /* CKTnumGminSteps, CKTnumSrcSteps, CKTdcTrcvMaxIter (cktntask.c:117-122) */
    double gmin_factor;                        /* CKTgminFactor */
    double gshunt;                             /* CKTgshunt (`.option gshunt`): what the gmin ladders of CKTop end on */
} NgbOpts;

/* two-terminal linear elements and sources */
typedef struct NgbCapCtx {
    int ninst, S, T;
    const int *nodes;       /* [2][ninst] pos, neg                                      */
    const double *par;      /* [3][T]  capacitance, m, initial condition                */
    const int *spos;        /* [6][ninst] stamp rows: pp nn pn np rhs_pos rhs_neg        */
    double *state;          /* [NGB_NHIST][2][T]  q, i                                   */
    double *stamp;
    const double *x; int neq1;
    NgbCtl ctl;
} NgbCapCtx;

typedef struct NgbSrcCtx {
    int ninst, S, T;
    int is_current;         /* 0: VSRC (rhs row = branch), 1: ISRC (rhs rows = pos,neg)  */
    const int *fn;          /* [3][ninst] function type, order, dcGiven                  */
    const double *par;      /* VSRC [9][T]: dc, coeffs[8] ; ISRC [10][T]: dc, m, coeffs[8] */
    const int *spos;        /* VSRC [1][ninst] rhs_branch ; ISRC [2][ninst] rhs_pos rhs_neg */
    double *stamp;
    double tstep, tstop;    /* CKTstep, CKTfinalTime (PULSE/SINE defaults)               */
    /* PWL sources: corner lists t0 v0 t1 v1 ... (shared by the samples), pwl_ptr [ninst+1] into pwl;
     * pwl_rdelay [ninst] = VSRCrdelay, pwl_rep [ninst] = VSRCrBreakpt when the list repeats, else -1 */
    const int *pwl_ptr, *pwl_rep;
    const double *pwl, *pwl_rdelay;
    NgbCtl ctl;
} NgbSrcCtx;

/* assembly of Ax / rhs from the stamp buffer: target t sums rows tgt_rows[tgt_ptr[t]..tgt_ptr[t+1])
 * in that (reference load) order; targets [0,nnz) are CSC slots, [nnz, nnz+neq+1) are rhs rows */
/* pivoting events of one run (niiter.c:107-111, 333-349): 0 the MODEINITJCT iteration, 1 the iteration after it (rest of the
 * operating point), 2 the first iteration under MODEINITTRAN, 3 the iteration after it (rest of the transient) */
#define NGB_LU_EVENTS 4
#define NGB_LU_SETS 16     /* 0..3: the pivoting events of a run; the rest: per-sample re-pivots after a zero pivot (ngb_tran.c) */
#define NGB_ASM_LONG 4096
#define NGB_ASM_CHUNK 256
typedef struct NgbAsmCtx {
    int S, nnz, neq1;
    const int *tgt_ptr, *tgt_rows;
    const int *slot_diag;   /* [nnz] 1 if the slot is a diagonal entry (LoadGmin_CSC)    */
    /* targets with more than NGB_ASM_LONG contributions (supply rails of a large flat circuit) are left to a tree of
     * launches: chunks of NGB_ASM_CHUNK consecutive contributions summed in order by one thread each, then chunks of chunk
     * totals, ... -- deterministic, every level fully parallel (a rail of 10^6 contributions: 3 907 + 16 + 1 threads) */
    const int *long_tgt; int nlong;
    const int *long_len_host;   /* HOST pointer [nlong]: contributions per long target (launch geometry) */
    double *long_part;          /* [2][long_cap] chunk totals, ping-pong between levels */
    int long_cap;
    const double *stamp;
    double *Ax;             /* [S][nnz]                                                  */
    double *x;              /* rhs is assembled into x[1 - xsel]                         */
    int add_diag_gmin;
    /* .nodeset / .ic row overrides of CKTload (cktload.c:118-172): nov rows in reference order (nodesets, then
     * initial conditions); ov_zptr/ov_zslot list the voltage-column entries ZeroNoncurRow clears, ov_cur says
     * whether the row keeps an entry in a current column, ov_val [nov][S] is the forced voltage */
    int nov;
    const int *ov_eq, *ov_kind, *ov_cur, *ov_diag, *ov_zptr, *ov_zslot;
    const double *ov_val;
    NgbCtl ctl;
} NgbAsmCtx;

/* numeric LU on a fixed pattern + pivot order, as element tasks (DESIGN.md "LU") */
typedef struct NgbLuSched {
    int n;                  /* matrix order (after node collapsing)                      */
    int nnz;                /* entries of A                                              */
    int nV;                 /* LU values: L, U, Udiag, Offx entries                      */
    int nlev;               /* factor levels                                             */
    const int *lev_ptr;     /* [nlev+1] into lev_ent                                     */
    const int *lev_ent;     /* [nV] entry ids by level                                   */
    const int *e_aslot;     /* [nV] A slot feeding the entry or -1                       */
    const int *e_arow;      /* [nV] original row (index of Rs) of the entry              */
    const int *e_div;       /* [nV] value id of the pivot to divide by (L entries) or -1 */
    const int *e_pptr;      /* [nV+1] into pair_l / pair_u                               */
    const int *pair_l, *pair_u;
    const int *diag_v;      /* [n] value id of Udiag[k]                                  */
    const int *vchk;        /* [nV] or NULL: what KLU's pivot rule (lpivot, klu_kernel.c:370-470) asks of an L entry for this
                             * pivot order to be the one klu_factor would choose: 0 nothing, 1 column with its diagonal as pivot
                             * (|l| * tol <= 1), 2 column with another pivot (|l| < 1), 3 the displaced diagonal (|l| < tol) */
    const int *row_ptr;     /* [n+1] CSR view of A for the row scale factors             */
    const int *row_slot;
    /* triangular solves as 2n row tasks */
    int ntask, nslev;
    const int *slev_ptr;    /* [nslev+1]                                                 */
    const int *slev_task;   /* [ntask]                                                   */
    const int *t_kind;      /* 0: y (forward), 1: x (backward, divides by pivot)         */
    const int *t_init;      /* y: original row of b ; x: task id of y                    */
    const int *t_div;       /* x: value id of the pivot                                  */
    const int *t_pptr;      /* [ntask+1]                                                 */
    const int *t_val, *t_src;
    const int *b_eq;        /* [n] equation number feeding original row i (node collapsing) */
    const int *out_task;    /* [n] x-task whose result is the solution of column q       */
    const int *out_eq;      /* [n] equation number receiving it                          */
} NgbLuSched;

/* the same schedule renumbered so that every level is a contiguous index range, with all
 * per-level data as 16-bit indices in one blob that a CTA copies to shared memory once */
typedef struct NgbLuPacked {
    int ok;                 /* usable: every index fits 16 bits                           */
    int n, nnz, nV, nlev, ntask, nslev, npairs, nsp;
    int blob_u16;           /* length of blob in 16-bit words (even)                      */
    int o_lev_ptr, o_div, o_pptr, o_pl, o_pu, o_diag;              /* factor, offsets in blob */
    int o_slev_ptr, o_kind, o_init, o_tdiv, o_tpptr, o_tval, o_tsrc, o_out;   /* solve      */
    int maxlp;                                  /* most pairs any one factor/solve level has     */
    int o_aslot, o_arow, o_rowptr, o_rowslot;   /* A -> value map and CSR of A (0xFFFF = none)   */
    const unsigned short *blob;
    const int *aslot, *arow, *ext;      /* [nV] internal order: A slot, original row, external id */
    const int *row_ptr, *row_slot;      /* CSR view of A                                   */
    const int *b_eq, *out_eq;           /* [n]                                             */
    /* second packing (ngb_lu_sample_pk2): one 8-byte record per level / value / task and one 4-byte
     * word per product, so that the dependent index loads of a level collapse into one load each;
     * levels whose entries need no arithmetic (level 0) are not visited at all */
    int ok2;
    int blob2_u16;          /* length of blob2 in 16-bit words (multiple of 4)            */
    int lev0, e0;           /* first factor level with work, first value of that level    */
    int slev0;              /* first solve level with work                                */
    int o2_levd;            /* [nlev]  {first item, last item, pbase, pend}               */
    int o2_emeta;           /* [items] {p0, p1, div or 0xFFFF, target value}              */
    int o2_pair;            /* [np]    {l, u}                                             */
    int o2_diag;            /* [n]                                                        */
    int o2_slotmap;         /* [nnz]   {value, row} of A slot j                           */
    int o2_rowptr, o2_rowv; /* [n+1], [nnz] values of row i (row scale factors)           */
    int o2_slevd;           /* [nslev] {lo, hi, pbase, pend}                              */
    int o2_tmeta;           /* [items] {p0, p1, start task, div or 0xFFFF}                */
    int o2_ttgt;            /* [items] target task                                        */
    int o2_tpair;           /* [nsp]   {value, source task}                               */
    int o2_yinit;           /* [n]     {task, row, equation, 0} of the forward solve      */
    int o2_eqtask;          /* [neq1]  task holding the solution of equation i, 0xFFFF none */
    const unsigned short *blob2;
} NgbLuPacked;

typedef struct NgbLuCtx {
    NgbLuSched sch;
    NgbLuPacked pk;
    int S, neq1;
    const double *Ax;       /* [S][nnz]                                                  */
    double *V;              /* [S][nV] LU values (kept for a later solve / parity checks) */
    double *Rs;             /* [S][n]                                                    */
    double *x;              /* rhs in x[1 - xsel], overwritten by the solution           */
    int do_factor, do_solve;
    int which;              /* pattern set of this launch: samples with ctl.lusel != which are skipped */
    /* node convergence test of NIconvTest fused after the solve */
    const int *node_type;   /* [neq1] SP_VOLTAGE(3) / SP_CURRENT(4)                      */
    double reltol, abstol, vntol;
    int *nodeconv;          /* [S] 1 if some node failed                                 */
    int *singular_col;      /* [S] -1 or first zero pivot column                         */
    const int *verify;      /* [S] or NULL: 1 = a pivoting event (SMPreorder) is due for the sample in this iteration: after the
                             * refactor on this set's order, check that the order is what the pivoting factor would choose
                             * (sch.vchk); if not, report E_SINGULAR with singular_col = -2 and the host factors with pivoting */
    double pivtol;
    /* work arrays of the grid-wide LU (a circuit too large for one CTA's shared memory): values, scale factors, solve tasks */
    double *gV, *gRs, *gZ;  /* [S][nV], [S][n], [S][ntask]                               */
    NgbCtl ctl;
} NgbLuCtx;

#ifdef __cplusplus
}
#endif
#endif
