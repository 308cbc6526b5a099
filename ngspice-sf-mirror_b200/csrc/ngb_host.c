/* ngb_host.c -- host side of the hot path (C): circuit flattening results, CSC pattern and
 * slot map, stamp/contribution lists, LU task schedule, batch memory and the C-ABI calls that
 * launch the kernels.  Everything device-side goes through ngb_dev.h.
 *
 * Reference code whose role this file plays:
 *   SMPmakeElt / SMPconvertCOOtoCSC / DEVbindCSC   src/maths/KLU/klusmp.c:137-323,417-440,
 *                                                  src/spicelib/devices/bsim4/b4set.c:2587-2676
 *   CKTload driver                                 src/spicelib/analysis/cktload.c:32-180
 *   SMPluFac / SMPsolve wrappers                   src/maths/KLU/klusmp.c:573-624, 950-1011
 *   klu_refactor column loop (turned into a task schedule here)  src/maths/KLU/klu_refactor.c:285-426
 *   klu_solve / KLU_lsolve / KLU_usolve (ditto)    src/maths/KLU/klu_solve.c, klu.c:201-345
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include "ngb_dev.h"
#include "ngb_host.h"
#include "../../include/ngb200.h"

static __thread char g_err[512] = "";      /* per host thread, like the launch stream */
void ngb_set_error(const char *fmt, ...)
{
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
}
const char *ngbLastError(void) { return g_err; }
const char *ngbBackend(void) { return ngb_dev_backend(); }
long ngbLaunchCount(void) { return ngb_dev_launch_count(); }
int ngbInit(int device) { return ngb_dev_init(device); }
int ngbSetStream(void *cuda_stream) { return ngb_dev_set_stream(cuda_stream); }
int ngbSync(void) { return ngb_dev_sync(); }
void ngbProfile(int enable, int every) { ngb_dev_profile(enable, every); }
int ngbProfileRead(double *ms_sum, long *count) { return ngb_dev_profile_read(ms_sum, count); }
int ngbProfileStages(double ms[8]) { return ngb_dev_stage_read(ms); }
static unsigned b4_batch_key(const ngb_circuit *c, const double *mtab, int nrows, const int *prow_inst);
int ngbMeasureFp64Peak(double out[3]) { return ngb_dev_fp64_peak(out); }

/* ------------------------------------------------------------------ field names */
#define X(n) #n,
static const char *b4_model_names[] = { NGB_B4_MODEL_FIELDS(X) NULL };
static const char *b4_bin_names[] = { NGB_B4_BIN_FIELDS(X) NULL };
static const char *b4_inst_names[] = { NGB_B4_INST_FIELDS(X) NULL };
static const char *b4_node_names[] = { NGB_B4_NODE_FIELDS(X) NULL };
static const char *b4_stamp_names[] = { NGB_B4_MAT_FIELDS(X) NGB_B4_RHS_FIELDS(X) NGB_B4_EXTRA_FIELDS(X) NULL };
static const char *b4_op_names[] = { NGB_B4_OP_FIELDS(X) NULL };
#undef X

void ngbBsim4Layout(int out[8])
{
    out[0] = B4M_COUNT; out[1] = B4P_COUNT; out[2] = B4I_COUNT; out[3] = B4N_COUNT;
    out[4] = B4S_MAT_COUNT; out[5] = B4S_COUNT; out[6] = B4ST_COUNT; out[7] = B4O_COUNT;
}
static const char *b3_model_names[B3M_COUNT], *b3_bin_names[B3P_COUNT], *b3_inst_names[B3I_COUNT];
static const char *dio_par_names[] = {
#define X(n) #n,
    NGB_DIO_INST_FIELDS(X)
    NGB_DIO_MODEL_FIELDS(X)
    NGB_DIO_RAW_INST_FIELDS(X)
    NGB_DIO_RAW_MODEL_FIELDS(X)
#undef X
};
void ngbDioLayout(int out[3]) { out[0] = DIOP_COUNT; out[1] = DIOST_COUNT; out[2] = DIOS_COUNT; }
const char *ngbBsim4FieldName(int list, int index)
{
    const char **t; int n;
    switch (list) {
    case 0: t = b4_model_names; n = B4M_COUNT; break;
    case 1: t = b4_bin_names; n = B4P_COUNT; break;
    case 2: t = b4_inst_names; n = B4I_COUNT; break;
    case 3: t = b4_node_names; n = B4N_COUNT; break;
    case 4: case 5: t = b4_stamp_names; n = B4S_COUNT; break;
    case 7: t = b4_op_names; n = B4O_COUNT; break;
    case 8: t = dio_par_names; n = DIOP_COUNT; break;
    case 9: t = b3_model_names; n = B3M_COUNT; break;
    case 10: t = b3_bin_names; n = B3P_COUNT; break;
    case 11: t = b3_inst_names; n = B3I_COUNT; break;
    default: return NULL;
    }
    return (index >= 0 && index < n) ? t[index] : NULL;
}

/* (row role, column role) of each BSIM4 matrix stamp position: the TSTALLOC table of
 * b4set.c:2587-2676 re-expressed over node roles */
static int b4_role_of(const char *tok, int len)
{
    static const struct { const char *t; int role; } map[] = {
        { "dp", B4N_dNodePrime }, { "gp", B4N_gNodePrime }, { "gm", B4N_gNodeMid },
        { "ge", B4N_gNodeExt }, { "sp", B4N_sNodePrime }, { "bp", B4N_bNodePrime },
        { "db", B4N_dbNode }, { "sb", B4N_sbNode }, { "d", B4N_dNode }, { "s", B4N_sNode },
        { "b", B4N_bNode }, { "q", B4N_qNode } };
    char low[4]; int i;
    for (i = 0; i < len && i < 3; i++) low[i] = (char)(tok[i] | 0x20);
    low[i] = 0;
    for (i = 0; i < (int)(sizeof map / sizeof map[0]); i++)
        if (!strcmp(map[i].t, low)) return map[i].role;
    return -1;
}
static void b4_stamp_roles(int k, int *rrow, int *rcol)
{
    const char *nm = b4_stamp_names[k];
    int up = 0;
    while (nm[up] >= 'A' && nm[up] <= 'Z') up++;
    *rrow = b4_role_of(nm, up);
    *rcol = b4_role_of(nm + up, (int)strlen(nm + up));
}
static int b4_rhs_role(int k)     /* k in [B4S_MAT_COUNT, B4S_COUNT) */
{
    const char *nm = b4_stamp_names[k];
    return b4_role_of(nm, (int)strlen(nm));
}

/* is the matrix position allocated by BSIM4setup for these selectors (b4set.c:2587-2676)? */
static int b4_pos_allocated(int k, int rgateMod, int rbodyMod, int rdsMod)
{
    if (k >= B4S_GEge && k <= B4S_BPgm) return rgateMod != 0 || (k >= B4S_GPgp && k <= B4S_GPbp);
    if (k >= B4S_Dgp && k <= B4S_Sbp) return rdsMod != 0;
    if (k >= B4S_DPdb && k <= B4S_Bb) return rbodyMod == 1 || rbodyMod == 2;
    return 1;
}
/* does the load write the position (b4ld.c:5235-5388, trnqsMod == 0)? */
static int b4_pos_written(int k, int rgateMod, int rbodyMod, int rdsMod)
{
    if (k >= B4S_MAT_COUNT) {
        switch (k) {
        case B4R_dp: case B4R_gp: case B4R_bp: case B4R_sp: return 1;
        case B4R_ge: return rgateMod == 2;
        case B4R_gm: return rgateMod == 3;
        case B4R_db: case B4R_sb: return rbodyMod != 0;
        case B4R_d: case B4R_s: return rdsMod != 0;
        default: return 0;   /* q */
        }
    }
    if (k >= B4S_GPgp && k <= B4S_GPbp) return 1;
    switch (k) {
    case B4S_GEge: case B4S_GPge: return rgateMod == 1 || rgateMod == 2 || (k == B4S_GEge && rgateMod == 3);
    case B4S_GEgp: return rgateMod == 1 || rgateMod == 2;
    case B4S_GEdp: case B4S_GEsp: case B4S_GEbp: return rgateMod == 2;
    case B4S_GEgm: case B4S_GMge: case B4S_GMgm: case B4S_GMdp: case B4S_GMgp: case B4S_GMsp:
    case B4S_GMbp: case B4S_DPgm: case B4S_GPgm: case B4S_SPgm: case B4S_BPgm: return rgateMod == 3;
    default: break;
    }
    if (k >= B4S_Dgp && k <= B4S_Sbp) return rdsMod != 0;
    if (k >= B4S_DPdb && k <= B4S_Bb) return rbodyMod != 0;
    if (k >= B4S_Qq && k <= B4S_GPq) return 0;
    return 1;
}

/* ------------------------------------------------------------------ small containers */
typedef struct { int *v; int n, cap; } ivec;
static void iv_push(ivec *a, int x)
{
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 256; a->v = (int *)realloc(a->v, sizeof(int) * (size_t)a->cap); }
    a->v[a->n++] = x;
}
static void *xcalloc(size_t n, size_t sz) { void *p = calloc(n ? n : 1, sz); if (!p) { fprintf(stderr, "ngb: out of memory\n"); abort(); } return p; }
static void *xdup(const void *src, size_t bytes) { void *p = xcalloc(bytes ? bytes : 1, 1); if (bytes) memcpy(p, src, bytes); return p; }

/* ------------------------------------------------------------------ circuit */
ngb_circuit *ngbCircuitCreate(int neq, const int *node_type)
{
    ngb_circuit *c = (ngb_circuit *)xcalloc(1, sizeof *c);
    int i;
    c->neq = neq; c->pivot_mode = -1;
    c->node_type = (int *)xcalloc((size_t)neq + 1, sizeof(int));
    for (i = 0; i <= neq; i++) c->node_type[i] = node_type ? node_type[i] : 3;
    /* defaults of cktntask.c:95-146 */
    c->opt.reltol = 1e-3; c->opt.abstol = 1e-12; c->opt.vntol = 1e-6; c->opt.chgtol = 1e-14;
    c->opt.trtol = 7; c->opt.temp = 300.15; c->opt.vt0 = 1.38064852e-23 * (27.0 + 273.15) / 1.6021766208e-19;
    c->opt.xmu = 0.5; c->opt.gmin = 1e-12; c->opt.method = NGB_TRAPEZOIDAL; c->opt.maxorder = 2;
    c->opt.itl4 = 10; c->opt.itl1 = 100;
    c->opt.num_gmin_steps = 1; c->opt.num_src_steps = 1; c->opt.itl2 = 50; c->opt.gmin_factor = 10;
    c->exact_order = 1;
    return c;
}

static void free_packed(NgbLuPacked *p);
static void free_sched(NgbLuSched *h)
{
#define F(p) free((void *)h->p)
    F(lev_ptr); F(lev_ent); F(e_aslot); F(e_arow); F(e_div); F(e_pptr); F(pair_l); F(pair_u);
    F(diag_v); F(row_ptr); F(row_slot); F(slev_ptr); F(slev_task); F(t_kind); F(t_init); F(t_div);
    F(t_pptr); F(t_val); F(t_src); F(b_eq); F(out_task); F(out_eq); F(vchk);
#undef F
    memset(h, 0, sizeof *h);
}

void ngbCircuitDestroy(ngb_circuit *c)
{
    if (!c) return;
    free(c->node_type);
    free(c->b4_nodes); free(c->b4_flags); free(c->b4_prow); free(c->b4_inst); free(c->b4_mtab); free(c->b4_ptab);
    free(c->b4_spos); free(c->b4_slots);
    free(c->res_nodes); free(c->res_g); free(c->res_spos);
    free(c->cap_nodes); free(c->cap_par); free(c->cap_spos);
    free(c->b3_nodes); free(c->b3_flags); free(c->b3_prow); free(c->b3_inst); free(c->b3_mtab); free(c->b3_ptab); free(c->b3_spos);
    free(c->vb_nodes); free(c->vb_flags); free(c->vb_par); free(c->vb_aux); free(c->vb_spos);
    free(c->dio_nodes); free(c->dio_flags); free(c->dio_par); free(c->dio_spos);
    free(c->vs_nodes); free(c->vs_fn); free(c->vs_par); free(c->vs_spos); free(c->vs_cspos);
    free(c->vs_pwl_ptr); free(c->vs_pwl_rep); free(c->vs_pwl_rdelay); free(c->vs_pwl_len); free(c->vs_pwl);
    free(c->is_pwl_ptr); free(c->is_pwl_len); free(c->is_pwl);
    free(c->is_nodes); free(c->is_fn); free(c->is_par); free(c->is_spos);
    free(c->Ap); free(c->Ai); free(c->eq2col); free(c->col2eq); free(c->slot_diag); free(c->diag_slot);
    free(c->ov_eq); free(c->ov_kind); free(c->ov_cur); free(c->ov_diag); free(c->ov_zptr); free(c->ov_zslot); free(c->ov_val);
    free(c->long_tgt); free(c->tgt_ptr); free(c->tgt_rows); free(c->const_row); free(c->const_val);
    free(c->klu_Q); free(c->klu_R); free(c->klu_Pnum); free(c->klu_P);
    free(c->pat_Lp); free(c->pat_Li); free(c->pat_Up); free(c->pat_Ui); free(c->pat_Offp); free(c->pat_Offi);
    free_sched(&c->sch);
    free_packed(&c->pk);
    { int w; for (w = 0; w < NGB_LU_SETS; w++) if (c->lu[w].valid) { free_sched(&c->lu[w].sch); free_packed(&c->lu[w].pk); } }
    free(c);
}

int ngbCircuitSetOptions(ngb_circuit *c, const double d[15], const int i[5])
{
    NgbOpts *o = &c->opt;
    o->reltol = d[0]; o->abstol = d[1]; o->vntol = d[2]; o->chgtol = d[3]; o->trtol = d[4];
    o->temp = d[5]; o->vt0 = d[6]; o->xmu = d[7]; o->tstep = d[8]; o->tstop = d[9]; o->tmax = d[10];
    o->tstart = d[11]; o->delmin = d[12]; o->minbreak = d[13]; o->gmin = d[14];
    o->method = i[0]; o->maxorder = i[1]; o->itl4 = i[2]; o->itl1 = i[3]; o->uic = i[4];
    if (o->method != NGB_TRAPEZOIDAL && o->method != NGB_GEAR) { ngb_set_error("integration method %d unknown (1 TRAPEZOIDAL, 2 GEAR)", o->method); return NGB_E_METHOD; }
    /* DCtran never raises CKTorder above 2 (dctran.c:794-826), so maxord > 2 only lengthens the state ring (cktsetup.c:192) */
    if (o->maxorder > 2) { ngb_set_error("maxord %d not supported (the state ring holds maxord + 2 <= 4 vectors)", o->maxorder); return NGB_E_ORDER; }
    return NGB_OK;
}

/* the operating-point fallbacks of CKTop (cktop.c:62-96): CKTnumGminSteps and CKTnumSrcSteps (0 skips the route, 1 is
 * dynamic_gmin + new_gmin / gillespie_src; larger counts select spice3_gmin / spice3_src, which are not built),
 * CKTdcTrcvMaxIter (itl2), CKTgminFactor, and CKTnoOpIter (`.option noopiter`: no plain NIiter, straight to the fallbacks) */
int ngbCircuitSetOpFallbacks(ngb_circuit *c, int num_gmin_steps, int num_src_steps, int itl2, double gmin_factor, int no_op_iter, double gshunt)
{
    if (num_gmin_steps < 0 || num_gmin_steps > 1) { ngb_set_error("gminsteps=%d selects spice3_gmin, which is not on this path (0 or 1)", num_gmin_steps); return NGB_E_UNSUPP; }
    if (num_src_steps < 0 || num_src_steps > 1) { ngb_set_error("srcsteps=%d selects spice3_src, which is not on this path (0 or 1)", num_src_steps); return NGB_E_UNSUPP; }
    if (itl2 < 1 || !(gmin_factor > 1.0)) { ngb_set_error("itl2=%d / gminfactor=%g out of range", itl2, gmin_factor); return NGB_E_PANIC; }
    c->opt.num_gmin_steps = num_gmin_steps; c->opt.num_src_steps = num_src_steps; c->opt.itl2 = itl2; c->opt.gmin_factor = gmin_factor;
    c->opt.no_op_iter = no_op_iter ? 1 : 0;
    c->opt.gshunt = gshunt > 0 ? gshunt : 0.0;
    return NGB_OK;
}

/* 1 (default): every `+=` of the reference load gets its own stamp row, so Ax/rhs are summed in
 * exactly the reference order; 0: addends to the same pointer are pre-summed in the kernel
 * (17 fewer stamp rows per BSIM4 instance, results equal to summation-order rounding) */
int ngbCircuitSetExactOrder(ngb_circuit *c, int on)
{
    if (c->finalized) { ngb_set_error("stamp mode must be chosen before ngbCircuitFinalize"); return NGB_E_PANIC; }
    c->exact_order = on ? 1 : 0;
    return NGB_OK;
}

int ngbCircuitAddBsim4(ngb_circuit *c, int n, const int *nodes, const int *flags, const int *prow,
                       const double *inst, int nrows, const double *mtab, const double *ptab)
{
    int i;
    if (c->finalized || c->b4_n) { ngb_set_error("BSIM4 table already set"); return NGB_E_PANIC; }
    for (i = 0; i < n; i++) {
        if (flags[i] & 0x300) { ngb_set_error("BSIM4 instance %d: trnqsMod/acnqsMod != 0 is not supported on this path", i); return NGB_E_UNSUPP; }
        if (prow[i] < 0 || prow[i] >= nrows) { ngb_set_error("BSIM4 instance %d: bad parameter row", i); return NGB_E_PANIC; }
    }
    c->b4_n = n; c->b4_nrows = nrows;
    c->b4_nodes = (int *)xdup(nodes, sizeof(int) * (size_t)n * B4N_COUNT);
    c->b4_flags = (int *)xdup(flags, sizeof(int) * (size_t)n);
    c->b4_prow = (int *)xdup(prow, sizeof(int) * (size_t)n);
    c->b4_inst = (double *)xdup(inst, sizeof(double) * (size_t)n * B4I_COUNT);
    c->b4_mtab = (double *)xdup(mtab, sizeof(double) * (size_t)nrows * B4M_COUNT);
    c->b4_ptab = (double *)xdup(ptab, sizeof(double) * (size_t)nrows * B4P_COUNT);
    return NGB_OK;
}
int ngbCircuitAddResistors(ngb_circuit *c, int n, const int *nodes, const double *g)
{
    if (c->finalized || c->res_n) return NGB_E_PANIC;
    c->res_n = n; c->res_nodes = (int *)xdup(nodes, sizeof(int) * 2 * (size_t)n);
    c->res_g = (double *)xdup(g, sizeof(double) * (size_t)n);
    return NGB_OK;
}
int ngbCircuitAddCapacitors(ngb_circuit *c, int n, const int *nodes, const double *par)
{
    if (c->finalized || c->cap_n) return NGB_E_PANIC;
    c->cap_n = n; c->cap_nodes = (int *)xdup(nodes, sizeof(int) * 2 * (size_t)n);
    c->cap_par = (double *)xdup(par, sizeof(double) * 3 * (size_t)n);
    return NGB_OK;
}
/* (row role, column role) of every BSIM3 stamp: the TSTALLOC table of b3set.c:1090-1111 */
static const int b3_stamp_r[B3S_COUNT] = {
#define d B3N_d
#define g B3N_g
#define s B3N_s
#define b B3N_b
#define dp B3N_dp
#define sp B3N_sp
#define X(n, r, cc) r,
    NGB_B3_STAMPS(X)
#undef X
};
static const int b3_stamp_c[B3S_COUNT] = {
#define X(n, r, cc) cc,
    NGB_B3_STAMPS(X)
#undef X
#undef d
#undef g
#undef s
#undef b
#undef dp
#undef sp
};
static const char *b3_model_names[] = {
#define X(n) #n,
    NGB_B3_MODEL_FIELDS(X)
#undef X
};
static const char *b3_bin_names[] = {
#define X(n) #n,
    NGB_B3_BIN_FIELDS(X)
#undef X
};
static const char *b3_inst_names[] = {
#define X(n) #n,
    NGB_B3_INST_FIELDS(X)
#undef X
};
void ngbBsim3Layout(int out[6])
{ out[0] = B3M_COUNT; out[1] = B3P_COUNT; out[2] = B3I_COUNT; out[3] = B3N_COUNT; out[4] = B3S_COUNT; out[5] = B3ST_COUNT; }

int ngbCircuitAddBsim3(ngb_circuit *c, int ninst, const int *nodes, const int *flags, const int *prow,
                       const double *inst, int nrows, const double *mtab, const double *ptab)
{
    int i;
    if (c->finalized || c->b3_n) return NGB_E_PANIC;
    for (i = 0; i < ninst; i++) {
        const double *mr;
        if (prow[i] < 0 || prow[i] >= nrows) { ngb_set_error("BSIM3 instance %d: parameter row %d out of range", i, prow[i]); return NGB_E_PANIC; }
        mr = mtab + (size_t)prow[i] * B3M_COUNT;
        if (flags[i] & B3F_NQS) { ngb_set_error("BSIM3 instance %d: nqsMod/acnqsMod not supported", i); return NGB_E_UNSUPP; }
        if ((int)mr[B3M_acmMod] != 0) { ngb_set_error("BSIM3 instance %d: acmMod=%d not supported (0 only)", i, (int)mr[B3M_acmMod]); return NGB_E_UNSUPP; }
        if ((int)mr[B3M_capMod] < 0 || (int)mr[B3M_capMod] > 3) {
            ngb_set_error("BSIM3 instance %d: capMod=%d out of range", i, (int)mr[B3M_capMod]); return NGB_E_UNSUPP; }
    }
    c->b3_n = ninst; c->b3_nrows = nrows;
    c->b3_nodes = (int *)xdup(nodes, sizeof(int) * B3N_COUNT * (size_t)ninst);
    c->b3_flags = (int *)xdup(flags, sizeof(int) * (size_t)ninst);
    c->b3_prow = (int *)xdup(prow, sizeof(int) * (size_t)ninst);
    c->b3_inst = (double *)xdup(inst, sizeof(double) * B3I_COUNT * (size_t)ninst);
    c->b3_mtab = (double *)xdup(mtab, sizeof(double) * B3M_COUNT * (size_t)nrows);
    c->b3_ptab = (double *)xdup(ptab, sizeof(double) * B3P_COUNT * (size_t)nrows);
    return NGB_OK;
}
/* VBIC: structural entries and stamp statements as (row role, column role) tables */
#define coll VBN_coll
#define base VBN_base
#define emit VBN_emit
#define subs VBN_subs
#define cx VBN_cx
#define ci VBN_ci
#define bx VBN_bx
#define bi VBN_bi
#define ei VBN_ei
#define bp VBN_bp
#define si VBN_si
#define temp VBN_temp
#define xf1 VBN_xf1
#define xf2 VBN_xf2
static const int vb_struct_r[] = {
#define T(r, cc) r,
    NGB_VBIC_STRUCT(T)
#undef T
};
static const int vb_struct_c[] = {
#define T(r, cc) cc,
    NGB_VBIC_STRUCT(T)
#undef T
};
static const int vb_stamp_r[] = {
#define SR(n, v) n,
#define SM(r, cc, v) r,
    NGB_VBIC_STAMPS(SR, SM, SR, SM, SR, SM, SR, SM)
#undef SR
#undef SM
};
static const int vb_stamp_c[] = {           /* -1: right-hand side */
#define SR(n, v) -1,
#define SM(r, cc, v) cc,
    NGB_VBIC_STAMPS(SR, SM, SR, SM, SR, SM, SR, SM)
#undef SR
#undef SM
};
static const int vb_stamp_need[] = {        /* instance flags a stamp statement exists for (0: every instance) */
#define F0(a, v) 0,
#define F0M(a, b, v) 0,
#define FX(a, v) VBF_EXCESS,
#define FXM(a, b, v) VBF_EXCESS,
#define FS(a, v) VBF_SELFHEAT,
#define FSM(a, b, v) VBF_SELFHEAT,
#define FSX(a, v) (VBF_SELFHEAT | VBF_EXCESS),
#define FSXM(a, b, v) (VBF_SELFHEAT | VBF_EXCESS),
    NGB_VBIC_STAMPS(F0, F0M, FX, FXM, FS, FSM, FSX, FSXM)
#undef F0
#undef F0M
#undef FX
#undef FXM
#undef FS
#undef FSM
#undef FSX
#undef FSXM
};
#undef coll
#undef base
#undef emit
#undef subs
#undef cx
#undef ci
#undef bx
#undef bi
#undef ei
#undef bp
#undef si
#undef temp
#undef xf1
#undef xf2
void ngbVbicLayout(int out[5]) { out[0] = VBIC_NP; out[1] = VBA_COUNT; out[2] = VBN_COUNT; out[3] = VBS_COUNT; out[4] = VBIC_NSTAMPS; }

int ngbCircuitAddVbic(ngb_circuit *c, int n, const int *nodes, const int *flags, const double *par, const double *aux)
{
    int i;
    if (c->finalized || c->vb_n) return NGB_E_PANIC;
    for (i = 0; i < n; i++) {
        /* the thermal node and the filter nodes exist exactly for the instances that carry the flag (vbicsetup.c:470-523) */
        const int tn = nodes[VBN_temp * n + i], x1 = nodes[VBN_xf1 * n + i], x2 = nodes[VBN_xf2 * n + i];
        if (((flags[i] & VBF_SELFHEAT) != 0) != (tn > 0) || ((flags[i] & VBF_EXCESS) != 0) != (x1 > 0 && x2 > 0)) {
            ngb_set_error("VBIC instance %d: flags 0x%x do not match its thermal / excess-phase nodes (%d, %d, %d)", i, flags[i], tn, x1, x2);
            return NGB_E_PANIC;
        }
    }
    c->vb_n = n;
    c->vb_nodes = (int *)xdup(nodes, sizeof(int) * VBN_COUNT * (size_t)n);
    c->vb_flags = (int *)xdup(flags, sizeof(int) * (size_t)n);
    c->vb_par = (double *)xdup(par, sizeof(double) * VBIC_NP * (size_t)n);
    c->vb_aux = (double *)xdup(aux, sizeof(double) * VBA_COUNT * (size_t)n);
    return NGB_OK;
}
int ngbCircuitAddDiodes(ngb_circuit *c, int n, const int *nodes, const int *flags, const double *par)
{
    int i;
    if (c->finalized || c->dio_n) return NGB_E_PANIC;
    for (i = 0; i < n; i++)
        if (((flags[i] & DIOF_SELFHEAT) != 0) != (nodes[4 * n + i] > 0) || ((flags[i] & DIOF_REVREC) != 0 && nodes[5 * n + i] <= 0)) {
            ngb_set_error("diode %d: flags 0x%x do not match its thermal / qp nodes (%d, %d)", i, flags[i], nodes[4 * n + i], nodes[5 * n + i]);
            return NGB_E_PANIC;
        }
    c->dio_n = n; c->dio_nodes = (int *)xdup(nodes, sizeof(int) * DION_COUNT * (size_t)n);
    c->dio_flags = (int *)xdup(flags, sizeof(int) * (size_t)n);
    c->dio_par = (double *)xdup(par, sizeof(double) * DIOP_COUNT * (size_t)n);
    return NGB_OK;
}
int ngbCircuitAddVsources(ngb_circuit *c, int n, const int *nodes, const int *fn, const double *par)
{
    int i;
    if (c->finalized || c->vs_n) return NGB_E_PANIC;
    for (i = 0; i < n; i++)
        if (fn[i] < 0 || fn[i] > NGB_FN_AM) {
            ngb_set_error("voltage source %d: waveform type %d not supported (DC, PULSE, SIN, EXP, SFFM, PWL, AM)", i, fn[i]);
            return NGB_E_UNSUPP;
        }
    c->vs_n = n; c->vs_nodes = (int *)xdup(nodes, sizeof(int) * 3 * (size_t)n);
    c->vs_fn = (int *)xdup(fn, sizeof(int) * 3 * (size_t)n);
    c->vs_par = (double *)xdup(par, sizeof(double) * 9 * (size_t)n);
    return NGB_OK;
}
/* corner list of a PWL voltage source (VSRCcoeffs, VSRCfunctionOrder entries: t0 v0 t1 v1 ...), its delay
 * VSRCrdelay and, for a repeating list, the index VSRCrBreakpt of the corner the repetition restarts from (-1: none) */
int ngbCircuitSetVsourcePwl(ngb_circuit *c, int inst, int ncoef, const double *coef, double rdelay, int rbreakpt)
{
    int i, at;
    if (c->finalized || inst < 0 || inst >= c->vs_n || ncoef < 2 || (ncoef & 1)) return NGB_E_PANIC;
    if (!c->vs_pwl_ptr) {
        c->vs_pwl_ptr = (int *)xcalloc((size_t)c->vs_n + 1, sizeof(int));
        c->vs_pwl_rep = (int *)xcalloc((size_t)c->vs_n, sizeof(int));
        c->vs_pwl_rdelay = (double *)xcalloc((size_t)c->vs_n, sizeof(double));
        c->vs_pwl_len = (int *)xcalloc((size_t)c->vs_n, sizeof(int));
        for (i = 0; i < c->vs_n; i++) c->vs_pwl_rep[i] = -1;
    }
    if (c->vs_pwl_len[inst]) return NGB_E_PANIC;                  /* once per instance */
    at = c->vs_pwl_n;
    c->vs_pwl = (double *)realloc(c->vs_pwl, sizeof(double) * (size_t)(at + ncoef));
    memcpy(c->vs_pwl + at, coef, sizeof(double) * (size_t)ncoef);
    c->vs_pwl_n = at + ncoef;
    c->vs_pwl_ptr[inst] = at; c->vs_pwl_len[inst] = ncoef; c->vs_pwl_rdelay[inst] = rdelay; c->vs_pwl_rep[inst] = rbreakpt;
    c->vs_fn[c->vs_n + inst] = ncoef;                              /* VSRCfunctionOrder */
    return NGB_OK;
}
int ngbCircuitAddIsources(ngb_circuit *c, int n, const int *nodes, const int *fn, const double *par)
{
    int i;
    if (c->finalized || c->is_n) return NGB_E_PANIC;
    for (i = 0; i < n; i++)
        if (fn[i] < 0 || fn[i] > NGB_FN_AM) {
            ngb_set_error("current source %d: waveform type %d not supported (DC, PULSE, SIN, EXP, SFFM, PWL, AM)", i, fn[i]);
            return NGB_E_UNSUPP;
        }
    c->is_n = n; c->is_nodes = (int *)xdup(nodes, sizeof(int) * 2 * (size_t)n);
    c->is_fn = (int *)xdup(fn, sizeof(int) * 3 * (size_t)n);
    c->is_par = (double *)xdup(par, sizeof(double) * 10 * (size_t)n);
    return NGB_OK;
}

/* corner list of a PWL current source (ISRCcoeffs, ISRCfunctionOrder entries); ISRCload / ISRCaccept know neither a
 * delay nor a repetition (isrcload.c:291-314, isrcacct.c:181-197) */
int ngbCircuitSetIsourcePwl(ngb_circuit *c, int inst, int ncoef, const double *coef)
{
    int at;
    if (c->finalized || inst < 0 || inst >= c->is_n || ncoef < 2 || (ncoef & 1)) return NGB_E_PANIC;
    if (!c->is_pwl_ptr) {
        c->is_pwl_ptr = (int *)xcalloc((size_t)c->is_n + 1, sizeof(int));
        c->is_pwl_len = (int *)xcalloc((size_t)c->is_n, sizeof(int));
    }
    if (c->is_pwl_len[inst]) return NGB_E_PANIC;                  /* once per instance */
    at = c->is_pwl_n;
    c->is_pwl = (double *)realloc(c->is_pwl, sizeof(double) * (size_t)(at + ncoef));
    memcpy(c->is_pwl + at, coef, sizeof(double) * (size_t)ncoef);
    c->is_pwl_n = at + ncoef;
    c->is_pwl_ptr[inst] = at; c->is_pwl_len[inst] = ncoef;
    c->is_fn[c->is_n + inst] = ncoef;                              /* ISRCfunctionOrder */
    return NGB_OK;
}

/* ---- pattern ---- */
typedef struct { int row, col; } coo_t;
static int coo_cmp(const void *a, const void *b)
{
    const coo_t *x = (const coo_t *)a, *y = (const coo_t *)b;
    if (x->col != y->col) return x->col < y->col ? -1 : 1;
    if (x->row != y->row) return x->row < y->row ? -1 : 1;
    return 0;
}
typedef struct { coo_t *v; int n, cap; } coovec;
static void coo_push(coovec *a, int r, int cl)
{
    if (r <= 0 || cl <= 0) return;               /* ground row/column: the trash cell of SMPmakeElt */
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 1024; a->v = (coo_t *)realloc(a->v, sizeof(coo_t) * (size_t)a->cap); }
    a->v[a->n].row = r - 1; a->v[a->n].col = cl - 1; a->n++;
}
static int slot_lookup(const ngb_circuit *c, int req, int ceq)
{
    int col, row, lo, hi;
    if (req <= 0 || ceq <= 0) return -1;
    col = c->eq2col[ceq]; row = c->eq2col[req];
    if (col < 0 || row < 0) return -1;
    lo = c->Ap[col]; hi = c->Ap[col + 1] - 1;
    while (lo <= hi) { int mid = (lo + hi) / 2; if (c->Ai[mid] == row) return mid; if (c->Ai[mid] < row) lo = mid + 1; else hi = mid - 1; }
    return -1;
}

typedef struct { ivec tgt, row; } contribs;
static int new_row(ngb_circuit *c, contribs *cb, int target)
{
    int r;
    if (target < 0) return -1;
    r = c->nstamp_rows++;
    iv_push(&cb->tgt, target); iv_push(&cb->row, r);
    return r;
}

/* same, with a caller-chosen row: the big device tables number their rows position-major
 * (row = base + position * ninst + instance) so that neighbouring threads -- neighbouring samples of
 * one instance, or neighbouring instances when S = 1 -- write neighbouring words */
static int new_row_at(ngb_circuit *c, contribs *cb, int target, int row)
{
    (void)c;
    if (target < 0) return -1;
    iv_push(&cb->tgt, target); iv_push(&cb->row, row);
    return row;
}

int ngbCircuitFinalize(ngb_circuit *c)
{
    coovec coo = { 0, 0, 0 };
    contribs cb = { { 0, 0, 0 }, { 0, 0, 0 } };
    ivec crow = { 0, 0, 0 };
    double *cval = NULL; int ncval = 0, capcval = 0;
    int i, k, n, ncol;
    int *used;
    if (c->finalized) return NGB_OK;
    for (i = 0; i < c->vs_n; i++)
        if (c->vs_fn[i] == NGB_FN_PWL && !(c->vs_pwl_len && c->vs_pwl_len[i])) {
            ngb_set_error("voltage source %d is PWL and has no corner list (ngbCircuitSetVsourcePwl)", i); return NGB_E_PANIC; }
    for (i = 0; i < c->is_n; i++)
        if (c->is_fn[i] == NGB_FN_PWL && !(c->is_pwl_len && c->is_pwl_len[i])) {
            ngb_set_error("current source %d is PWL and has no corner list (ngbCircuitSetIsourcePwl)", i); return NGB_E_PANIC; }

    /* 1. structural entries, as the DEVsetup routines would request them */
    for (i = 0; i < c->b4_n; i++) {
        const int fl = c->b4_flags[i];
        const int rg = B4F_RGATE(fl), rb = B4F_RBODY(fl);
        const int rds = (int)c->b4_mtab[(size_t)c->b4_prow[i] * B4M_COUNT + B4M_rdsMod];
        for (k = 0; k < B4S_MAT_COUNT; k++) {
            int rr, rc;
            if (!b4_pos_allocated(k, rg, rb, rds)) continue;
            b4_stamp_roles(k, &rr, &rc);
            coo_push(&coo, c->b4_nodes[rr * c->b4_n + i], c->b4_nodes[rc * c->b4_n + i]);
        }
    }
    for (i = 0; i < c->b3_n; i++)
        for (k = B3S_RHS_COUNT; k < B3S_COUNT; k++)
            coo_push(&coo, c->b3_nodes[b3_stamp_r[k] * c->b3_n + i], c->b3_nodes[b3_stamp_c[k] * c->b3_n + i]);
    for (i = 0; i < c->cap_n; i++) {
        int p = c->cap_nodes[i], q = c->cap_nodes[c->cap_n + i];
        coo_push(&coo, p, p); coo_push(&coo, q, q); coo_push(&coo, p, q); coo_push(&coo, q, p);
    }
    for (i = 0; i < c->dio_n; i++) {                 /* diosetup.c:438-444 */
        int p = c->dio_nodes[i], q = c->dio_nodes[c->dio_n + i], pp = c->dio_nodes[2 * c->dio_n + i];
        coo_push(&coo, p, pp); coo_push(&coo, q, pp); coo_push(&coo, pp, p); coo_push(&coo, pp, q);
        coo_push(&coo, p, p); coo_push(&coo, q, q); coo_push(&coo, pp, pp);
        if (c->dio_flags[i] & DIOF_RESISTSW) {       /* diosetup.c:447-451 */
            int ps = c->dio_nodes[3 * c->dio_n + i];
            coo_push(&coo, p, ps); coo_push(&coo, q, ps); coo_push(&coo, ps, p); coo_push(&coo, ps, q); coo_push(&coo, ps, ps);
        }
        if (c->dio_flags[i] & DIOF_SELFHEAT) {       /* diosetup.c:454-467 */
            const int tn = c->dio_nodes[4 * c->dio_n + i];
            coo_push(&coo, tn, p); coo_push(&coo, tn, pp); coo_push(&coo, tn, q); coo_push(&coo, tn, tn);
            coo_push(&coo, p, tn); coo_push(&coo, pp, tn); coo_push(&coo, q, tn);
            if (c->dio_flags[i] & DIOF_RESISTSW) { int ps = c->dio_nodes[3 * c->dio_n + i]; coo_push(&coo, tn, ps); coo_push(&coo, ps, tn); }
        }
        if (c->dio_flags[i] & DIOF_REVREC) {         /* diosetup.c:470-476 */
            const int qn = c->dio_nodes[5 * c->dio_n + i];
            coo_push(&coo, qn, qn); coo_push(&coo, qn, pp); coo_push(&coo, qn, q); coo_push(&coo, pp, qn); coo_push(&coo, q, qn);
        }
    }
    for (i = 0; i < c->res_n; i++) {
        int p = c->res_nodes[i], q = c->res_nodes[c->res_n + i];
        coo_push(&coo, p, p); coo_push(&coo, q, q); coo_push(&coo, p, q); coo_push(&coo, q, p);
    }
    for (i = 0; i < c->vb_n; i++)
        for (k = 0; k < (int)(sizeof vb_struct_r / sizeof vb_struct_r[0]); k++)
            coo_push(&coo, c->vb_nodes[vb_struct_r[k] * c->vb_n + i], c->vb_nodes[vb_struct_c[k] * c->vb_n + i]);
    for (i = 0; i < c->vs_n; i++) {
        int p = c->vs_nodes[i], q = c->vs_nodes[c->vs_n + i], br = c->vs_nodes[2 * c->vs_n + i];
        coo_push(&coo, p, br); coo_push(&coo, q, br); coo_push(&coo, br, q); coo_push(&coo, br, p);
    }
    if (coo.n == 0) { ngb_set_error("empty matrix"); return NGB_E_PANIC; }

    /* 2. COO -> CSC: sort by column then row, drop empty columns ("node collapsing"), dedup */
    qsort(coo.v, (size_t)coo.n, sizeof(coo_t), coo_cmp);
    ncol = coo.v[coo.n - 1].col + 1;
    used = (int *)xcalloc((size_t)c->neq + 2, sizeof(int));
    for (i = 0; i < coo.n; i++) used[coo.v[i].col] = 1;
    c->eq2col = (int *)xcalloc((size_t)c->neq + 2, sizeof(int));
    c->col2eq = (int *)xcalloc((size_t)ncol + 1, sizeof(int));
    c->eq2col[0] = -1;
    n = 0;
    for (i = 0; i < c->neq; i++) {
        if (i < ncol && used[i]) { c->eq2col[i + 1] = n; c->col2eq[n] = i + 1; n++; }
        else c->eq2col[i + 1] = -1;
    }
    free(used);
    c->n = n;
    c->Ap = (int *)xcalloc((size_t)n + 1, sizeof(int));
    c->Ai = (int *)xcalloc((size_t)coo.n, sizeof(int));
    c->nnz = 0;
    {
        int prev_r = -1, prev_c = -1;
        for (i = 0; i < coo.n; i++) {
            int cc = c->eq2col[coo.v[i].col + 1], rr = c->eq2col[coo.v[i].row + 1];
            if (rr < 0) { ngb_set_error("row %d has entries but its column is structurally empty", coo.v[i].row + 1); free(coo.v); return NGB_E_PANIC; }
            if (cc == prev_c && rr == prev_r) continue;
            c->Ai[c->nnz] = rr; c->Ap[cc + 1] = c->nnz + 1; c->nnz++;
            prev_r = rr; prev_c = cc;
        }
        for (i = 0; i < n; i++) if (c->Ap[i + 1] < c->Ap[i]) c->Ap[i + 1] = c->Ap[i];
    }
    free(coo.v);
    c->slot_diag = (int *)xcalloc((size_t)c->nnz, sizeof(int));
    c->diag_slot = (int *)xcalloc((size_t)n, sizeof(int));
    for (i = 0; i < n; i++) {
        int p; c->diag_slot[i] = -1;
        for (p = c->Ap[i]; p < c->Ap[i + 1]; p++) if (c->Ai[p] == i) { c->slot_diag[p] = 1; c->diag_slot[i] = p; }
    }

    /* 3. stamp rows and contribution lists, in CKTload order: device types by their rank in
     *    the reference device table (bsim3 < bsim4 < cap < dio < isrc < res < vbic < vsrc, dev.c:142-209), instances
     *    in list order, positions in load order */
    c->nstamp_rows = 0;
    c->b3_spos = (int *)xcalloc((size_t)c->b3_n * B3S_COUNT + 1, sizeof(int));
    for (i = 0; i < c->b3_n; i++)
        for (k = 0; k < B3S_COUNT; k++) {
            const int rr = c->b3_nodes[b3_stamp_r[k] * c->b3_n + i], rc = c->b3_nodes[b3_stamp_c[k] * c->b3_n + i];
            const int row = k * c->b3_n + i;
            if (k < B3S_RHS_COUNT) c->b3_spos[k * c->b3_n + i] = new_row_at(c, &cb, (rr > 0 && c->eq2col[rr] >= 0) ? c->nnz + rr : -1, row);
            else c->b3_spos[k * c->b3_n + i] = new_row_at(c, &cb, slot_lookup(c, rr, rc), row);
        }
    c->nstamp_rows += B3S_COUNT * c->b3_n;
    {
        const int b4_base = c->nstamp_rows;
        c->b4_row0 = b4_base;
#define B4ROW(k_) (b4_base + (k_) * c->b4_n + i)
    c->b4_spos = (int *)xcalloc((size_t)c->b4_n * B4S_TOTAL, sizeof(int));
    c->b4_slots = (int *)xcalloc((size_t)c->b4_n * B4S_MAT_COUNT, sizeof(int));
    for (i = 0; i < c->b4_n; i++) {
        const int fl = c->b4_flags[i];
        const int rg = B4F_RGATE(fl), rb = B4F_RBODY(fl);
        const int rds = (int)c->b4_mtab[(size_t)c->b4_prow[i] * B4M_COUNT + B4M_rdsMod];
        int ord[B4S_TOTAL], no = 0, q;
        /* statement order of the serial load: right-hand side (b4ld.c:5024-5053), then matrix
         * (b4ld.c:5235-5388) */
        ord[no++] = B4R_dp; ord[no++] = B4R_gp;
        if (rg == 2) ord[no++] = B4R_ge; else if (rg == 3) ord[no++] = B4R_gm;
        if (!rb) { ord[no++] = B4R_bp; ord[no++] = B4R_sp; }
        else { ord[no++] = B4R_db; ord[no++] = B4R_bp; ord[no++] = B4R_sb; ord[no++] = B4R_sp; }
        if (rds) { ord[no++] = B4R_d; ord[no++] = B4R_s; }
        if (rg == 1) { static const int g[] = { B4S_GEge, B4S_GPge, B4S_GEgp, B4S_GPgp, B4S_GPdp, B4S_GPsp, B4S_GPbp };
            for (q = 0; q < 7; q++) ord[no++] = g[q]; }
        else if (rg == 2) { static const int g[] = { B4S_GEge, B4S_GEgp, B4S_GEdp, B4S_GEsp, B4S_GEbp, B4S_GPge, B4S_GPgp, B4S_GPdp, B4S_GPsp, B4S_GPbp };
            for (q = 0; q < 10; q++) ord[no++] = g[q]; }
        else if (rg == 3) { static const int g[] = { B4S_GEge, B4S_GEgm, B4S_GMge, B4S_GMgm, B4S_GMdp, B4S_GMgp, B4S_GMsp, B4S_GMbp,
                                                     B4S_DPgm, B4S_GPgm, B4S_SPgm, B4S_BPgm, B4S_GPgp, B4S_GPdp, B4S_GPsp, B4S_GPbp };
            for (q = 0; q < 16; q++) ord[no++] = g[q]; }
        else { static const int g[] = { B4S_GPgp, B4S_GPdp, B4S_GPsp, B4S_GPbp }; for (q = 0; q < 4; q++) ord[no++] = g[q]; }
        if (rds) { static const int g[] = { B4S_Dgp, B4S_Dsp, B4S_Dbp, B4S_Sdp, B4S_Sgp, B4S_Sbp }; for (q = 0; q < 6; q++) ord[no++] = g[q]; }
        { static const int g[] = { B4S_DPdp, B4S_DPd, B4S_DPgp, B4S_DPsp, B4S_DPbp, B4S_Ddp, B4S_Dd, B4S_SPdp, B4S_SPgp, B4S_SPsp,
                                   B4S_SPs, B4S_SPbp, B4S_Ssp, B4S_Ss, B4S_BPdp, B4S_BPgp, B4S_BPsp, B4S_BPbp };
          for (q = 0; q < 18; q++) ord[no++] = g[q]; }
        if (c->exact_order) { static const int g[] = { B4X_DPdp_g, B4X_DPgp_g, B4X_DPsp_g, B4X_DPbp_g, B4X_BPdp_g, B4X_BPgp_g, B4X_BPsp_g, B4X_BPbp_g,
                                                       B4X_SPdp_s, B4X_SPgp_s, B4X_SPsp_s, B4X_SPbp_s, B4X_BPdp_s, B4X_BPgp_s, B4X_BPsp_s, B4X_BPbp_s };
          for (q = 0; q < 16; q++) ord[no++] = g[q]; }
        if (rb) { static const int g[] = { B4S_DPdb, B4S_SPsb, B4S_DBdp, B4S_DBdb, B4S_DBbp, B4S_DBb, B4S_BPdb, B4S_BPb, B4S_BPsb, -1,
                                           B4S_SBsp, B4S_SBbp, B4S_SBb, B4S_SBsb, B4S_Bdb, B4S_Bbp, B4S_Bsb, B4S_Bb };
          for (q = 0; q < 18; q++) { if (g[q] >= 0) ord[no++] = g[q]; else if (c->exact_order) ord[no++] = B4X_BPbp_r; } }
        for (k = 0; k < B4S_TOTAL; k++) c->b4_spos[k * c->b4_n + i] = -1;
        for (k = 0; k < B4S_MAT_COUNT; k++) {
            int rr, rc, slot = -1;
            b4_stamp_roles(k, &rr, &rc);
            if (b4_pos_allocated(k, rg, rb, rds))
                slot = slot_lookup(c, c->b4_nodes[rr * c->b4_n + i], c->b4_nodes[rc * c->b4_n + i]);
            c->b4_slots[k * c->b4_n + i] = slot;
        }
        for (q = 0; q < no; q++) {
            k = ord[q];
            if (k >= B4S_MAT_COUNT && k < B4S_COUNT) {
                int eq = c->b4_nodes[b4_rhs_role(k) * c->b4_n + i];
                if (b4_pos_written(k, rg, rb, rds) && eq > 0 && c->eq2col[eq] >= 0)
                    c->b4_spos[k * c->b4_n + i] = new_row_at(c, &cb, c->nnz + eq, B4ROW(k));
            } else {
                int base = k, slot;
                if (k >= B4S_COUNT) {                 /* extra addend: same slot as its base position */
                    char nm[16]; int kk; const char *x = b4_stamp_names[k]; size_t l = strlen(x) - 2;
                    memcpy(nm, x, l); nm[l] = 0; base = -1;
                    for (kk = 0; kk < B4S_MAT_COUNT; kk++) if (!strcmp(b4_stamp_names[kk], nm)) base = kk;
                }
                slot = c->b4_slots[base * c->b4_n + i];
                if (slot >= 0 && b4_pos_written(base, rg, rb, rds))
                    c->b4_spos[k * c->b4_n + i] = new_row_at(c, &cb, slot, B4ROW(k));
            }
        }
    }
#undef B4ROW
    c->nstamp_rows += B4S_TOTAL * c->b4_n;
    }
    c->cap_spos = (int *)xcalloc((size_t)c->cap_n * 6 + 1, sizeof(int));
    for (i = 0; i < c->cap_n; i++) {
        int p = c->cap_nodes[i], q = c->cap_nodes[c->cap_n + i], nn = c->cap_n;
        c->cap_spos[0 * nn + i] = new_row(c, &cb, slot_lookup(c, p, p));
        c->cap_spos[1 * nn + i] = new_row(c, &cb, slot_lookup(c, q, q));
        c->cap_spos[2 * nn + i] = new_row(c, &cb, slot_lookup(c, p, q));
        c->cap_spos[3 * nn + i] = new_row(c, &cb, slot_lookup(c, q, p));
        c->cap_spos[4 * nn + i] = new_row(c, &cb, p > 0 ? c->nnz + p : -1);
        c->cap_spos[5 * nn + i] = new_row(c, &cb, q > 0 ? c->nnz + q : -1);
    }
    c->dio_spos = (int *)xcalloc((size_t)c->dio_n * DIOS_COUNT + 1, sizeof(int));
    for (i = 0; i < c->dio_n; i++) {
        const int nn = c->dio_n;
        int p = c->dio_nodes[i], q = c->dio_nodes[nn + i], pp = c->dio_nodes[2 * nn + i], ps = c->dio_nodes[3 * nn + i], k2;
        const int sw = (c->dio_flags[i] & DIOF_RESISTSW) != 0;
        for (k2 = 0; k2 < DIOS_COUNT; k2++) c->dio_spos[k2 * nn + i] = -1;
        {
            const int th = (c->dio_flags[i] & DIOF_SELFHEAT) != 0, rr = (c->dio_flags[i] & DIOF_REVREC) != 0;
            const int tn = c->dio_nodes[4 * nn + i], qn = c->dio_nodes[5 * nn + i];
#define DROW_R(pos, node) c->dio_spos[(pos) * nn + i] = new_row(c, &cb, (node) > 0 ? c->nnz + (node) : -1)
#define DROW_M(pos, r, cc) c->dio_spos[(pos) * nn + i] = new_row(c, &cb, slot_lookup(c, (r), (cc)))
            DROW_R(DIOS_rhsNeg, q); DROW_R(DIOS_rhsPosPrime, pp);
            if (th) { DROW_R(DIOS_thRhsPos, p); DROW_R(DIOS_thRhsPp, pp); DROW_R(DIOS_thRhsNeg, q); DROW_R(DIOS_thRhsTemp, tn); }
            if (sw) {
                DROW_R(DIOS_rhsNegSw, q); DROW_R(DIOS_rhsPosSwPrime, ps);
                if (th) { DROW_R(DIOS_thRhsPosSw, p); DROW_R(DIOS_thRhsPsp, ps); DROW_R(DIOS_thRhsNegSw, q); DROW_R(DIOS_thRhsTempSw, tn); }
            }
            DROW_M(DIOS_posPos, p, p); DROW_M(DIOS_negNeg, q, q); DROW_M(DIOS_ppPp, pp, pp); DROW_M(DIOS_posPp, p, pp);
            DROW_M(DIOS_negPp, q, pp); DROW_M(DIOS_ppPos, pp, p); DROW_M(DIOS_ppNeg, pp, q);
            if (th) {
                DROW_M(DIOS_thTempPos, tn, p); DROW_M(DIOS_thTempPp, tn, pp); DROW_M(DIOS_thTempNeg, tn, q); DROW_M(DIOS_thTempTemp, tn, tn);
                DROW_M(DIOS_thPosTemp, p, tn); DROW_M(DIOS_thPpTemp, pp, tn); DROW_M(DIOS_thNegTemp, q, tn);
            }
            if (sw) {
                DROW_M(DIOS_posPosSw, p, p); DROW_M(DIOS_negNegSw, q, q); DROW_M(DIOS_pspPsp, ps, ps); DROW_M(DIOS_posPsp, p, ps);
                DROW_M(DIOS_negPsp, q, ps); DROW_M(DIOS_pspPos, ps, p); DROW_M(DIOS_pspNeg, ps, q);
                if (th) {
                    DROW_M(DIOS_thTempPosSw, tn, p); DROW_M(DIOS_thTempPsp, tn, ps); DROW_M(DIOS_thTempNegSw, tn, q);
                    DROW_M(DIOS_thPosTempSw, p, tn); DROW_M(DIOS_thPspTemp, ps, tn); DROW_M(DIOS_thNegTempSw, q, tn);
                }
            }
            if (rr) {
                DROW_R(DIOS_rrRhsQp, qn); DROW_M(DIOS_rrQpQp, qn, qn); DROW_M(DIOS_rrQpPp, qn, pp); DROW_M(DIOS_rrQpNeg, qn, q);
                DROW_R(DIOS_rrRhsPp, pp); DROW_R(DIOS_rrRhsNeg, q); DROW_M(DIOS_rrPpQp, pp, qn); DROW_M(DIOS_rrNegQp, q, qn);
            }
#undef DROW_R
#undef DROW_M
        }
    }
    c->is_spos = (int *)xcalloc((size_t)c->is_n * 2 + 1, sizeof(int));
    for (i = 0; i < c->is_n; i++) {
        int p = c->is_nodes[i], q = c->is_nodes[c->is_n + i];
        c->is_spos[i] = new_row(c, &cb, p > 0 ? c->nnz + p : -1);
        c->is_spos[c->is_n + i] = new_row(c, &cb, q > 0 ? c->nnz + q : -1);
    }
#define CONST_ROW(target, value) do { int r_ = new_row(c, &cb, (target)); if (r_ >= 0) { iv_push(&crow, r_); \
        if (ncval == capcval) { capcval = capcval ? capcval * 2 : 256; cval = (double *)realloc(cval, sizeof(double) * (size_t)capcval); } \
        cval[ncval++] = (value); } } while (0)
    c->res_spos = (int *)xcalloc((size_t)c->res_n * 4 + 1, sizeof(int));
    for (i = 0; i < c->res_n; i++) {
        int p = c->res_nodes[i], q = c->res_nodes[c->res_n + i];
        double g = c->res_g[i];
        const int tg[4] = { slot_lookup(c, p, p), slot_lookup(c, q, q), slot_lookup(c, p, q), slot_lookup(c, q, p) };
        for (k = 0; k < 4; k++) {
            c->res_spos[k * c->res_n + i] = (tg[k] >= 0) ? c->nstamp_rows : -1;      /* the row CONST_ROW is about to create */
            CONST_ROW(tg[k], k < 2 ? g : -g);
        }
    }
    c->vb_spos = (int *)xcalloc((size_t)c->vb_n * VBIC_NSTAMPS + 1, sizeof(int));
    for (i = 0; i < c->vb_n; i++)
        for (k = 0; k < VBIC_NSTAMPS; k++) {
            const int rr = c->vb_nodes[vb_stamp_r[k] * c->vb_n + i];
            if ((c->vb_flags[i] & vb_stamp_need[k]) != vb_stamp_need[k]) { c->vb_spos[k * c->vb_n + i] = -1; continue; }   /* statement absent for this instance */
            if (vb_stamp_c[k] < 0) c->vb_spos[k * c->vb_n + i] = new_row(c, &cb, (rr > 0 && c->eq2col[rr] >= 0) ? c->nnz + rr : -1);
            else c->vb_spos[k * c->vb_n + i] = new_row(c, &cb, slot_lookup(c, rr, c->vb_nodes[vb_stamp_c[k] * c->vb_n + i]));
        }
    c->vs_spos = (int *)xcalloc((size_t)c->vs_n + 1, sizeof(int));
    c->vs_cspos = (int *)xcalloc((size_t)c->vs_n * 4 + 1, sizeof(int));
    for (i = 0; i < c->vs_n; i++) {
        int p = c->vs_nodes[i], q = c->vs_nodes[c->vs_n + i], br = c->vs_nodes[2 * c->vs_n + i];
        CONST_ROW(slot_lookup(c, p, br), 1.0); CONST_ROW(slot_lookup(c, q, br), -1.0);
        CONST_ROW(slot_lookup(c, br, p), 1.0); CONST_ROW(slot_lookup(c, br, q), -1.0);
        c->vs_spos[i] = new_row(c, &cb, br > 0 ? c->nnz + br : -1);
    }
#undef CONST_ROW
    c->nconst = crow.n; c->const_row = crow.v; c->const_val = cval;

    /* 4. contributions -> per-target lists (stable: keeps load order inside each target) */
    c->ntgt = c->nnz + c->neq + 1;
    c->tgt_ptr = (int *)xcalloc((size_t)c->ntgt + 1, sizeof(int));
    c->tgt_rows = (int *)xcalloc((size_t)cb.tgt.n, sizeof(int));
    for (i = 0; i < cb.tgt.n; i++) c->tgt_ptr[cb.tgt.v[i] + 1]++;
    for (i = 0; i < c->ntgt; i++) c->tgt_ptr[i + 1] += c->tgt_ptr[i];
    {
        int *fill = (int *)xdup(c->tgt_ptr, sizeof(int) * ((size_t)c->ntgt + 1));
        for (i = 0; i < cb.tgt.n; i++) c->tgt_rows[fill[cb.tgt.v[i]]++] = cb.row.v[i];
        free(fill);
    }
    free(cb.tgt.v); free(cb.row.v);
    c->nlong = 0;
    for (i = 0; i < c->ntgt; i++) if (c->tgt_ptr[i + 1] - c->tgt_ptr[i] > NGB_ASM_LONG) c->nlong++;
    c->long_tgt = (int *)xcalloc((size_t)c->nlong + 1, sizeof(int));
    for (i = 0, k = 0; i < c->ntgt; i++) if (c->tgt_ptr[i + 1] - c->tgt_ptr[i] > NGB_ASM_LONG) c->long_tgt[k++] = i;
    c->finalized = 1;
    return NGB_OK;
}

int ngbCircuitPatternSize(const ngb_circuit *c, int *n, int *nnz, int *nrows)
{
    if (!c->finalized) return NGB_E_PANIC;
    if (n) *n = c->n;
    if (nnz) *nnz = c->nnz;
    if (nrows) *nrows = c->nstamp_rows;
    return NGB_OK;
}
int ngbCircuitGetPattern(const ngb_circuit *c, int *Ap, int *Ai, int *diag)
{
    if (!c->finalized) return NGB_E_PANIC;
    if (Ap) memcpy(Ap, c->Ap, sizeof(int) * ((size_t)c->n + 1));
    if (Ai) memcpy(Ai, c->Ai, sizeof(int) * (size_t)c->nnz);
    if (diag) memcpy(diag, c->diag_slot, sizeof(int) * (size_t)c->n);
    return NGB_OK;
}
/* equation number (1-based, CKTnode number) of every column / row index of the pattern: columns without entries are not
 * in it, so in a circuit whose other device types are stamped elsewhere (the shim's table mode) index k is not equation k + 1 */
int ngbCircuitGetPatternEquations(const ngb_circuit *c, int *eq)
{
    if (!c->finalized) return NGB_E_PANIC;
    memcpy(eq, c->col2eq, sizeof(int) * (size_t)c->n);
    return NGB_OK;
}
int ngbCircuitGetBsim4Slots(const ngb_circuit *c, int *slots)
{
    if (!c->finalized) return NGB_E_PANIC;
    memcpy(slots, c->b4_slots, sizeof(int) * (size_t)c->b4_n * B4S_MAT_COUNT);
    return NGB_OK;
}

/* ------------------------------------------------------------------ LU task schedule */
static void free_packed(NgbLuPacked *p)
{
    free((void *)p->blob); free((void *)p->blob2); free((void *)p->aslot); free((void *)p->arow); free((void *)p->ext);
    memset(p, 0, sizeof *p);
}

/* renumber values and solve tasks level by level and pack the per-level data as 16-bit indices */
static void build_packed(ngb_circuit *c)
{
    const NgbLuSched *h = &c->sch;
    NgbLuPacked *p = &c->pk;
    const int nV = h->nV, n = h->n, ntask = h->ntask, np = c->npairs, nsp = c->nsolvepairs;
    int *vint, *tint, k, q, off = 0;
    unsigned short *b;
    int *aslot, *arow, *ext;
    free_packed(p);
    if (nV >= 65535 || ntask >= 65535 || np >= 65535 || nsp >= 65535 || n >= 65535 || h->nnz >= 65535) return;   /* generic kernel only */
    vint = (int *)xcalloc((size_t)nV, sizeof(int)); tint = (int *)xcalloc((size_t)ntask, sizeof(int));
    for (k = 0; k < nV; k++) vint[h->lev_ent[k]] = k;           /* lev_ent lists entries level by level */
    for (k = 0; k < ntask; k++) tint[h->slev_task[k]] = k;
    p->n = n; p->nnz = h->nnz; p->nV = nV; p->nlev = h->nlev; p->ntask = ntask; p->nslev = h->nslev; p->npairs = np; p->nsp = nsp;
#define SEG(field, cnt) p->field = off; off += (cnt)
    SEG(o_lev_ptr, h->nlev + 1); SEG(o_div, nV); SEG(o_pptr, nV + 1); SEG(o_pl, np); SEG(o_pu, np); SEG(o_diag, n);
    SEG(o_slev_ptr, h->nslev + 1); SEG(o_kind, ntask); SEG(o_init, ntask); SEG(o_tdiv, ntask); SEG(o_tpptr, ntask + 1);
    SEG(o_tval, nsp); SEG(o_tsrc, nsp); SEG(o_out, n);
    SEG(o_aslot, nV); SEG(o_arow, nV); SEG(o_rowptr, n + 1); SEG(o_rowslot, h->nnz);
#undef SEG
    if (off & 1) off++;
    p->blob_u16 = off;
    b = (unsigned short *)xcalloc((size_t)off, sizeof(unsigned short));
    aslot = (int *)xcalloc((size_t)nV, sizeof(int)); arow = (int *)xcalloc((size_t)nV, sizeof(int)); ext = (int *)xcalloc((size_t)nV, sizeof(int));
    for (k = 0; k <= h->nlev; k++) b[p->o_lev_ptr + k] = (unsigned short)h->lev_ptr[k];
    q = 0;
    for (k = 0; k < nV; k++) {
        const int e = h->lev_ent[k];           /* external id of internal entry k */
        int pp;
        ext[k] = e; aslot[k] = h->e_aslot[e]; arow[k] = h->e_arow[e];
        b[p->o_div + k] = (unsigned short)(h->e_div[e] >= 0 ? vint[h->e_div[e]] : 0xFFFF);
        b[p->o_pptr + k] = (unsigned short)q;
        for (pp = h->e_pptr[e]; pp < h->e_pptr[e + 1]; pp++, q++) {
            b[p->o_pl + q] = (unsigned short)vint[h->pair_l[pp]];
            b[p->o_pu + q] = (unsigned short)vint[h->pair_u[pp]];
        }
    }
    b[p->o_pptr + nV] = (unsigned short)q;
    for (k = 0; k < n; k++) b[p->o_diag + k] = (unsigned short)vint[h->diag_v[k]];
    for (k = 0; k <= h->nslev; k++) b[p->o_slev_ptr + k] = (unsigned short)h->slev_ptr[k];
    q = 0;
    for (k = 0; k < ntask; k++) {
        const int t = h->slev_task[k];
        int pp;
        b[p->o_kind + k] = (unsigned short)h->t_kind[t];
        b[p->o_init + k] = (unsigned short)(h->t_kind[t] == 0 ? h->t_init[t] : tint[h->t_init[t]]);
        b[p->o_tdiv + k] = (unsigned short)(h->t_div[t] >= 0 ? vint[h->t_div[t]] : 0xFFFF);
        b[p->o_tpptr + k] = (unsigned short)q;
        for (pp = h->t_pptr[t]; pp < h->t_pptr[t + 1]; pp++, q++) {
            b[p->o_tval + q] = (unsigned short)vint[h->t_val[pp]];
            b[p->o_tsrc + q] = (unsigned short)tint[h->t_src[pp]];
        }
    }
    b[p->o_tpptr + ntask] = (unsigned short)q;
    for (k = 0; k < n; k++) b[p->o_out + k] = (unsigned short)tint[h->out_task[k]];
    for (k = 0; k < nV; k++) {
        b[p->o_aslot + k] = (unsigned short)(aslot[k] >= 0 ? aslot[k] : 0xFFFF);
        b[p->o_arow + k] = (unsigned short)(aslot[k] >= 0 ? arow[k] : 0);
    }
    for (k = 0; k <= n; k++) b[p->o_rowptr + k] = (unsigned short)h->row_ptr[k];
    for (k = 0; k < h->nnz; k++) b[p->o_rowslot + k] = (unsigned short)h->row_slot[k];
    p->maxlp = 1;
    for (k = 0; k < h->nlev; k++) {
        const int np_lev = b[p->o_pptr + h->lev_ptr[k + 1]] - b[p->o_pptr + h->lev_ptr[k]];
        if (np_lev > p->maxlp) p->maxlp = np_lev;
    }
    for (k = 0; k < h->nslev; k++) {
        const int np_lev = b[p->o_tpptr + h->slev_ptr[k + 1]] - b[p->o_tpptr + h->slev_ptr[k]];
        if (np_lev > p->maxlp) p->maxlp = np_lev;
    }
    if (getenv("NGB_LU_STATS")) {          /* development aid: shape of the level schedule */
        int sum_max = 0, sum_prod = 0, e2;
        for (k = 0; k < h->nlev; k++) {
            int mx = 0;
            for (e2 = h->lev_ptr[k]; e2 < h->lev_ptr[k + 1]; e2++) { int cnt = b[p->o_pptr + e2 + 1] - b[p->o_pptr + e2]; if (cnt > mx) mx = cnt; }
            fprintf(stderr, "lu level %d: %d values, %d products, longest row %d\n", k, h->lev_ptr[k + 1] - h->lev_ptr[k],
                    b[p->o_pptr + h->lev_ptr[k + 1]] - b[p->o_pptr + h->lev_ptr[k]], mx);
            sum_max += mx; sum_prod += b[p->o_pptr + h->lev_ptr[k + 1]] - b[p->o_pptr + h->lev_ptr[k]];
        }
        fprintf(stderr, "lu factor: %d levels, %d values, %d products, sum of longest rows %d\n", h->nlev, nV, sum_prod, sum_max);
        sum_max = 0; sum_prod = 0;
        for (k = 0; k < h->nslev; k++) {
            int mx = 0;
            for (e2 = h->slev_ptr[k]; e2 < h->slev_ptr[k + 1]; e2++) { int cnt = b[p->o_tpptr + e2 + 1] - b[p->o_tpptr + e2]; if (cnt > mx) mx = cnt; }
            fprintf(stderr, "solve level %d: %d tasks, %d products, longest row %d\n", k, h->slev_ptr[k + 1] - h->slev_ptr[k],
                    b[p->o_tpptr + h->slev_ptr[k + 1]] - b[p->o_tpptr + h->slev_ptr[k]], mx);
            sum_max += mx; sum_prod += b[p->o_tpptr + h->slev_ptr[k + 1]] - b[p->o_tpptr + h->slev_ptr[k]];
        }
        fprintf(stderr, "lu solve: %d levels, %d tasks, %d products, sum of longest rows %d, blob %d u16\n", h->nslev, ntask, sum_prod, sum_max, off);
    }
    p->blob = b; p->aslot = aslot; p->arow = arow; p->ext = ext;
    p->row_ptr = h->row_ptr; p->row_slot = h->row_slot; p->b_eq = h->b_eq; p->out_eq = h->out_eq;
    p->ok = 1;
    /* ---- second packing: records instead of parallel index arrays (ngb_lu_sample_pk2) ----
     * A level is a list of ITEMS {first product, last product, pivot, target value}.  The subtractions of a value
     * are ordered, but the leading ones whose operands are two or more levels old do not have to wait for the
     * value's own level: they become a second item one level earlier (run by an idle lane, off the critical
     * path), and the value's own item keeps only the late products and the division.  Same operations, same
     * order per value. */
    do {
        const int neq1 = c->neq + 1;
        const int hoist = getenv("NGB_LU_NOHOIST") ? 0 : 1;
        int off2 = 0, lev0 = 0, slev0 = 0, e0, e2, *eqtask, L, nit = 0, nsit = 0, np2 = 0, nsp2 = 0, maxlp2 = 1;
        int *lv = NULL, *tlv = NULL;
        /* items: lev, p0 (into the level-ordered pair list), count, div, target, start (solve) */
        struct item { int lev, src0, cnt, dv, tgt, start, pre; } *fi = NULL, *si = NULL;
        int *fi_ptr = NULL, *si_ptr = NULL;
        unsigned short *b2;
        if (neq1 >= 65535) break;
        /* leading levels whose entries have neither products nor a pivot division stay as initialised */
        while (lev0 < h->nlev) {
            int work = 0;
            for (e2 = h->lev_ptr[lev0]; e2 < h->lev_ptr[lev0 + 1]; e2++)
                if (b[p->o_pptr + e2 + 1] != b[p->o_pptr + e2] || b[p->o_div + e2] != 0xFFFF) work = 1;
            if (work) break;
            lev0++;
        }
        while (slev0 < h->nslev) {
            int work = 0;
            for (e2 = h->slev_ptr[slev0]; e2 < h->slev_ptr[slev0 + 1]; e2++)
                if (b[p->o_tpptr + e2 + 1] != b[p->o_tpptr + e2] || b[p->o_kind + e2] != 0) work = 1;
            if (work) break;
            slev0++;
        }
        e0 = h->lev_ptr[lev0];
        eqtask = (int *)xcalloc((size_t)neq1, sizeof(int));
        for (k = 0; k < neq1; k++) eqtask[k] = 0xFFFF;
        for (k = 0; k < n; k++) {
            const int eq = h->out_eq[k];
            if (eq <= 0) continue;
            if (eq >= neq1 || eqtask[eq] != 0xFFFF) { eqtask[0] = -1; break; }     /* not a one-to-one map: first packing only */
            eqtask[eq] = tint[h->out_task[k]];
        }
        if (eqtask[0] == -1) { free(eqtask); break; }

        /* items of the factorisation */
        lv = (int *)xcalloc((size_t)nV + 1, sizeof(int)); tlv = (int *)xcalloc((size_t)ntask + 1, sizeof(int));
        for (L = 0; L < h->nlev; L++) for (k = h->lev_ptr[L]; k < h->lev_ptr[L + 1]; k++) lv[k] = L;
        for (L = 0; L < h->nslev; L++) for (k = h->slev_ptr[L]; k < h->slev_ptr[L + 1]; k++) tlv[k] = L;
        /* every product is subtracted in the first level in which (a) both operands are final and (b) the product
         * before it in the value's order has been subtracted: runs of products per (value, level) */
        fi = (struct item *)xcalloc((size_t)nV + (size_t)np + 2, sizeof *fi);
        for (k = e0; k < nV; k++) {
            const int q0 = b[p->o_pptr + k], q1 = b[p->o_pptr + k + 1];
            int q = q0, sprev = lev0;
            L = lv[k];
            while (q < q1) {
                int plv = lv[b[p->o_pl + q]] > lv[b[p->o_pu + q]] ? lv[b[p->o_pl + q]] : lv[b[p->o_pu + q]], run = q, sl;
                sl = plv + 1 > sprev ? plv + 1 : sprev;
                if (!hoist || sl > L) sl = L;
                while (run < q1) {
                    int pl2 = lv[b[p->o_pl + run]] > lv[b[p->o_pu + run]] ? lv[b[p->o_pl + run]] : lv[b[p->o_pu + run]];
                    if (hoist && pl2 + 1 > sl && sl < L) break;
                    run++;
                }
                if (sl < L) { struct item t = { sl, q, run - q, 0xFFFF, k, 0, 1 }; fi[nit++] = t; q = run; sprev = sl; }
                else break;                                  /* the rest belongs to the value's own level */
            }
            { struct item t = { L, q, q1 - q, b[p->o_div + k], k, 0, 0 }; fi[nit++] = t; }
        }
        /* items of the solve: a backward task starts from its forward task's value, which must be final first */
        si = (struct item *)xcalloc((size_t)ntask + (size_t)nsp + 2, sizeof *si);
        for (k = h->slev_ptr[slev0]; k < ntask; k++) {
            const int q0 = b[p->o_tpptr + k], q1 = b[p->o_tpptr + k + 1];
            const int kind = b[p->o_kind + k];
            int start = kind == 0 ? k : b[p->o_init + k];
            int q = q0, sprev = slev0;
            L = tlv[k];
            if (start != k && tlv[start] + 1 > sprev) sprev = tlv[start] + 1;
            while (q < q1) {
                int run = q, sl = tlv[b[p->o_tsrc + q]] + 1 > sprev ? tlv[b[p->o_tsrc + q]] + 1 : sprev;
                if (!hoist || sl > L) sl = L;
                while (run < q1) {
                    if (hoist && tlv[b[p->o_tsrc + run]] + 1 > sl && sl < L) break;
                    run++;
                }
                if (sl < L) { struct item t = { sl, q, run - q, 0xFFFF, k, start, 1 }; si[nsit++] = t; q = run; sprev = sl; start = k; }
                else break;
            }
            { struct item t = { L, q, q1 - q, kind == 1 ? b[p->o_tdiv + k] : 0xFFFF, k, start, 0 }; si[nsit++] = t; }
        }
        if (nit >= 65535 || nsit >= 65535) { free(eqtask); free(lv); free(tlv); free(fi); free(si); break; }
        /* level by level: the value's own items first, then the hoisted ones */
        fi_ptr = (int *)xcalloc((size_t)h->nlev + 2, sizeof(int)); si_ptr = (int *)xcalloc((size_t)h->nslev + 2, sizeof(int));
#define SEG2(field, cnt, width) off2 = (off2 + 3) & ~3; p->field = off2; off2 += (cnt) * (width)
        SEG2(o2_levd, h->nlev, 4); SEG2(o2_emeta, nit, 4); SEG2(o2_pair, np, 2); SEG2(o2_diag, n, 1);
        SEG2(o2_slotmap, h->nnz, 2); SEG2(o2_rowptr, n + 1, 1); SEG2(o2_rowv, h->nnz, 1);
        SEG2(o2_slevd, h->nslev, 4); SEG2(o2_tmeta, nsit, 4); SEG2(o2_ttgt, nsit, 1); SEG2(o2_tpair, nsp, 2);
        SEG2(o2_yinit, n, 4); SEG2(o2_eqtask, neq1, 1);
#undef SEG2
        off2 = (off2 + 3) & ~3;
        b2 = (unsigned short *)xcalloc((size_t)off2 + 4, sizeof(unsigned short));
        {
            int it = 0, pass, i2;
            for (L = 0; L < h->nlev; L++) {
                unsigned short *r = b2 + p->o2_levd + 4 * L;
                const int it0 = it, pb = np2;
                for (pass = 0; pass < 2; pass++)
                    for (i2 = 0; i2 < nit; i2++) {
                        unsigned short *m;
                        int q;
                        if (fi[i2].lev != L || fi[i2].pre != pass) continue;
                        m = b2 + p->o2_emeta + 4 * it++;
                        m[0] = (unsigned short)np2; m[1] = (unsigned short)(np2 + fi[i2].cnt); m[2] = (unsigned short)fi[i2].dv; m[3] = (unsigned short)fi[i2].tgt;
                        for (q = 0; q < fi[i2].cnt; q++, np2++) {
                            b2[p->o2_pair + 2 * np2] = b[p->o_pl + fi[i2].src0 + q];
                            b2[p->o2_pair + 2 * np2 + 1] = b[p->o_pu + fi[i2].src0 + q];
                        }
                    }
                r[0] = (unsigned short)it0; r[1] = (unsigned short)it; r[2] = (unsigned short)pb; r[3] = (unsigned short)np2;
                {   /* no run longer than one product: the items multiply for themselves (empty product range = no first phase) */
                    int direct = 1;
                    for (i2 = 0; i2 < nit; i2++) if (fi[i2].lev == L && fi[i2].cnt > 1) direct = 0;
                    if (direct) r[3] = r[2];
                    else if (np2 - pb > maxlp2) maxlp2 = np2 - pb;
                }
                fi_ptr[L + 1] = it;
            }
            it = 0;
            for (L = 0; L < h->nslev; L++) {
                unsigned short *r = b2 + p->o2_slevd + 4 * L;
                const int it0 = it, pb = nsp2;
                for (pass = 0; pass < 2; pass++)
                    for (i2 = 0; i2 < nsit; i2++) {
                        unsigned short *m;
                        int q;
                        if (si[i2].lev != L || si[i2].pre != pass) continue;
                        b2[p->o2_ttgt + it] = (unsigned short)si[i2].tgt;
                        m = b2 + p->o2_tmeta + 4 * it++;
                        m[0] = (unsigned short)nsp2; m[1] = (unsigned short)(nsp2 + si[i2].cnt); m[2] = (unsigned short)si[i2].start; m[3] = (unsigned short)si[i2].dv;
                        for (q = 0; q < si[i2].cnt; q++, nsp2++) {
                            b2[p->o2_tpair + 2 * nsp2] = b[p->o_tval + si[i2].src0 + q];
                            b2[p->o2_tpair + 2 * nsp2 + 1] = b[p->o_tsrc + si[i2].src0 + q];
                        }
                    }
                r[0] = (unsigned short)it0; r[1] = (unsigned short)it; r[2] = (unsigned short)pb; r[3] = (unsigned short)nsp2;
                {
                    int direct = 1;
                    for (i2 = 0; i2 < nsit; i2++) if (si[i2].lev == L && si[i2].cnt > 1) direct = 0;
                    if (direct) r[3] = r[2];
                    else if (nsp2 - pb > maxlp2) maxlp2 = nsp2 - pb;
                }
                si_ptr[L + 1] = it;
            }
        }
        for (k = 0; k < n; k++) b2[p->o2_diag + k] = b[p->o_diag + k];
        for (k = 0; k < h->nnz; k++) { b2[p->o2_slotmap + 2 * k] = 0xFFFF; b2[p->o2_slotmap + 2 * k + 1] = 0; }
        for (k = 0; k < nV; k++)
            if (aslot[k] >= 0) { b2[p->o2_slotmap + 2 * aslot[k]] = (unsigned short)k; b2[p->o2_slotmap + 2 * aslot[k] + 1] = (unsigned short)arow[k]; }
        for (k = 0; k <= n; k++) b2[p->o2_rowptr + k] = (unsigned short)h->row_ptr[k];
        {
            int bad = (np2 != np || nsp2 != nsp);           /* every product exactly once */
            for (k = 0; k < h->nnz; k++) {
                const int v = b2[p->o2_slotmap + 2 * h->row_slot[k]];
                if (v == 0xFFFF) bad = 1;               /* an entry of A outside the factors: first packing only */
                b2[p->o2_rowv + k] = (unsigned short)v;
            }
            q = 0;
            for (k = 0; k < ntask && !bad; k++)
                if (b[p->o_kind + k] == 0) {
                    unsigned short *y = b2 + p->o2_yinit + 4 * q;
                    const int row = b[p->o_init + k];
                    if (q >= n) { bad = 1; break; }
                    y[0] = (unsigned short)k; y[1] = (unsigned short)row; y[2] = (unsigned short)h->b_eq[row]; y[3] = 0;
                    q++;
                }
            if (q != n) bad = 1;
            if (bad) { free(b2); free(eqtask); free(lv); free(tlv); free(fi); free(si); free(fi_ptr); free(si_ptr); break; }
        }
        for (k = 0; k < neq1; k++) b2[p->o2_eqtask + k] = (unsigned short)eqtask[k];
        if (maxlp2 > p->maxlp) p->maxlp = maxlp2;
        p->lev0 = lev0; p->e0 = e0; p->slev0 = slev0; p->blob2 = b2; p->blob2_u16 = off2; p->ok2 = 1;
        if (getenv("NGB_LU_STATS")) {
            int npre = 0, nspre = 0, i2;
            for (i2 = 0; i2 < nit; i2++) npre += fi[i2].pre;
            for (i2 = 0; i2 < nsit; i2++) nspre += si[i2].pre;
            fprintf(stderr, "lu second packing: %d u16, first factor level %d, first solve level %d, %d + %d early runs, most products per level %d\n",
                    off2, lev0, slev0, npre, nspre, p->maxlp);
            for (L = lev0; L < h->nlev; L++) {
                int mx = 0;
                for (i2 = 0; i2 < nit; i2++) if (fi[i2].lev == L && fi[i2].cnt > mx) mx = fi[i2].cnt;
                fprintf(stderr, "  factor level %d: %d items, longest run %d\n", L, fi_ptr[L + 1] - fi_ptr[L], mx);
            }
            for (L = slev0; L < h->nslev; L++) {
                int mx = 0;
                for (i2 = 0; i2 < nsit; i2++) if (si[i2].lev == L && si[i2].cnt > mx) mx = si[i2].cnt;
                fprintf(stderr, "  solve level %d: %d items, longest run %d\n", L, si_ptr[L + 1] - si_ptr[L], mx);
            }
        }
        free(eqtask); free(lv); free(tlv); free(fi); free(si); free(fi_ptr); free(si_ptr);
    } while (0);
    free(vint); free(tint);
}

static unsigned long long lu_signature(int n, const int *Pnum, const int *Lp, const int *Li, const int *Up, const int *Ui)
{
    unsigned long long h = 1469598103934665603ULL;
    int k;
#define SIG(v) do { h ^= (unsigned long long)(unsigned)(v); h *= 1099511628211ULL; } while (0)
    for (k = 0; k < n; k++) SIG(Pnum[k]);
    for (k = 0; k <= n; k++) { SIG(Lp[k]); SIG(Up[k]); }
    for (k = 0; k < Lp[n]; k++) SIG(Li[k]);
    for (k = 0; k < Up[n]; k++) SIG(Ui[k]);
#undef SIG
    return h;
}

/* What lpivot's rule asks of every L entry for THIS pivot order to come out of a pivoting factor of the same matrix
 * (NgbLuSched.vchk).  The preferred row of a column is dynamic -- a displaced diagonal becomes the diagonal of the column whose
 * row was taken (klu_kernel.c:845-858, ngb_pivot.c) -- so the pivot sequence is replayed block by block.  Needs klu_analyze's
 * row permutation (ngbCircuitSetSymbolic); NULL without it */
static int *lu_pivot_checks(const ngb_circuit *c, int n, int nblocks, const int *R, const int *Pnum, const int *Lp, const int *Li, int unz, int nV)
{
    int *chk, *PSinv, *Pblk, *Pinv, *diagrow, *pivrow, b, k, p;
    if (!c->klu_P || c->klu_nblocks != nblocks) return NULL;
    chk = (int *)xcalloc((size_t)nV, sizeof(int)); PSinv = (int *)xcalloc((size_t)n, sizeof(int));
    Pblk = (int *)xcalloc((size_t)n, sizeof(int)); Pinv = (int *)xcalloc((size_t)n, sizeof(int));
    diagrow = (int *)xcalloc((size_t)n, sizeof(int)); pivrow = (int *)xcalloc((size_t)n, sizeof(int));
    for (k = 0; k < n; k++) PSinv[c->klu_P[k]] = k;
    for (b = 0; b < nblocks; b++) {
        const int k1 = R[b], k2 = R[b + 1], nk = k2 - k1;
        if (nk <= 1) continue;
        for (k = 0; k < nk; k++) { Pblk[k] = k; Pinv[k] = -k - 2; }
        for (k = 0; k < nk; k++) {
            const int pr = PSinv[Pnum[k1 + k]] - k1, dr = Pblk[k];
            if (pr < 0 || pr >= nk) { free(chk); chk = NULL; goto out; }        /* order and symbolic analysis do not belong together */
            diagrow[k1 + k] = dr; pivrow[k1 + k] = pr;
            if (pr != dr && Pinv[dr] < 0) { const int kbar = -Pinv[pr] - 2; Pblk[kbar] = dr; Pinv[dr] = -kbar - 2; }
            Pblk[k] = pr; Pinv[pr] = k;
        }
        for (k = k1; k < k2; k++)
            for (p = Lp[k]; p < Lp[k + 1]; p++) {
                const int i = PSinv[Pnum[Li[p]]] - k1;              /* the entry's row in the block's symbolic numbering */
                chk[unz + p] = (pivrow[k] == diagrow[k]) ? 1 : (i == diagrow[k] ? 3 : 2);
            }
    }
out:
    free(PSinv); free(Pblk); free(Pinv); free(diagrow); free(pivrow);
    return chk;
}

int ngbCircuitSetLuPattern(ngb_circuit *c, int n, int nblocks, const int *Q, const int *R, const int *Pnum,
                           const int *Lp, const int *Li, const int *Up, const int *Ui,
                           const int *Offp, const int *Offi)
{
    NgbLuSched *h = &c->sch;
    const int lnz = Lp[n], unz = Up[n], nzoff = Offp[n];
    const int nV = unz + lnz + n + nzoff;
    const int idL = unz, idD = unz + lnz, idO = unz + lnz + n;
    int *Pinv, *e_aslot, *e_arow, *e_div, *e_pptr, *pl, *pu, *pos, *level, *diag_v;
    int b, k, p, poff = 0, npairs = 0, pass, nlev = 0;
    if (!c->finalized) { ngb_set_error("circuit not finalized"); return NGB_E_PANIC; }
    if (n != c->n) { ngb_set_error("LU pattern order %d != matrix order %d", n, c->n); return NGB_E_PANIC; }
    free_sched(h);
    {   /* keep the factor as given (ngbCircuitGetLuPattern); the arguments may be these very arrays */
        int *q = (int *)xdup(Q, sizeof(int) * (size_t)n), *r = (int *)xdup(R, sizeof(int) * ((size_t)nblocks + 1));
        int *pn = (int *)xdup(Pnum, sizeof(int) * (size_t)n);
        int *lp = (int *)xdup(Lp, sizeof(int) * ((size_t)n + 1)), *li = (int *)xdup(Li, sizeof(int) * (size_t)(lnz > 0 ? lnz : 1));
        int *up = (int *)xdup(Up, sizeof(int) * ((size_t)n + 1)), *ui = (int *)xdup(Ui, sizeof(int) * (size_t)(unz > 0 ? unz : 1));
        int *op = (int *)xdup(Offp, sizeof(int) * ((size_t)n + 1)), *oi = (int *)xdup(Offi, sizeof(int) * (size_t)(nzoff > 0 ? nzoff : 1));
        free(c->klu_Q); free(c->klu_R); free(c->klu_Pnum);
        free(c->pat_Lp); free(c->pat_Li); free(c->pat_Up); free(c->pat_Ui); free(c->pat_Offp); free(c->pat_Offi);
        c->klu_Q = q; c->klu_R = r; c->klu_Pnum = pn; c->klu_nblocks = nblocks;
        c->pat_Lp = lp; c->pat_Li = li; c->pat_Up = up; c->pat_Ui = ui; c->pat_Offp = op; c->pat_Offi = oi;
        Q = q; R = r; Pnum = pn; Lp = lp; Li = li; Up = up; Ui = ui; Offp = op; Offi = oi;
    }
    c->lnz = lnz; c->unz = unz; c->nzoff = nzoff;

    Pinv = (int *)xcalloc((size_t)n, sizeof(int));
    for (k = 0; k < n; k++) Pinv[Pnum[k]] = k;
    e_aslot = (int *)xcalloc((size_t)nV, sizeof(int)); e_arow = (int *)xcalloc((size_t)nV, sizeof(int));
    e_div = (int *)xcalloc((size_t)nV, sizeof(int)); e_pptr = (int *)xcalloc((size_t)nV + 1, sizeof(int));
    diag_v = (int *)xcalloc((size_t)n, sizeof(int));
    pos = (int *)xcalloc((size_t)n, sizeof(int));
    level = (int *)xcalloc((size_t)nV, sizeof(int));
    for (k = 0; k < nV; k++) { e_aslot[k] = -1; e_div[k] = -1; }
    for (k = 0; k < n; k++) { pos[k] = -1; diag_v[k] = idD + k; }

    /* A -> entry map: the scatter loops of klu_refactor.c:300-372 */
    for (b = 0; b < nblocks; b++) {
        const int k1 = R[b], k2 = R[b + 1];
        for (k = k1; k < k2; k++) {
            const int oldcol = Q[k];
            if (k2 - k1 > 1) {
                for (p = Up[k]; p < Up[k + 1]; p++) pos[Ui[p]] = p;
                for (p = Lp[k]; p < Lp[k + 1]; p++) pos[Li[p]] = idL + p;
            }
            pos[k] = idD + k;
            for (p = c->Ap[oldcol]; p < c->Ap[oldcol + 1]; p++) {
                const int oldrow = c->Ai[p], newrow = Pinv[oldrow];
                int e;
                if (newrow < k1 && poff < nzoff) {
                    if (Offi[poff] != newrow) { ngb_set_error("off-diagonal pattern mismatch at column %d", k); goto bad; }
                    e = idO + poff; poff++;
                } else {
                    e = (newrow >= k1 && newrow < k2) ? pos[newrow] : -1;
                    if (e < 0) { ngb_set_error("A(%d,%d) has no place in the LU pattern", oldrow, oldcol); goto bad; }
                }
                e_aslot[e] = p; e_arow[e] = oldrow;
            }
            if (k2 - k1 > 1) {
                for (p = Up[k]; p < Up[k + 1]; p++) pos[Ui[p]] = -1;
                for (p = Lp[k]; p < Lp[k + 1]; p++) { pos[Li[p]] = -1; e_div[idL + p] = idD + k; }
            }
            pos[k] = -1;
        }
    }
    if (poff != nzoff) { ngb_set_error("off-diagonal count mismatch (%d != %d)", poff, nzoff); goto bad; }

    /* pairs in the order of the column loop of klu_refactor.c:377-389; two passes (count, fill) */
    pl = pu = NULL;
    for (pass = 0; pass < 2; pass++) {
        int *cursor = NULL;
        if (pass == 1) {
            int acc = 0;
            for (k = 0; k < nV; k++) { int cnt = e_pptr[k + 1]; e_pptr[k] = acc; acc += cnt; }
            /* shift: e_pptr[k+1] held counts of entry k */
            e_pptr[nV] = acc; npairs = acc;
            pl = (int *)xcalloc((size_t)npairs, sizeof(int)); pu = (int *)xcalloc((size_t)npairs, sizeof(int));
            cursor = (int *)xdup(e_pptr, sizeof(int) * ((size_t)nV + 1));
        }
        for (b = 0; b < nblocks; b++) {
            const int k1 = R[b], k2 = R[b + 1];
            if (k2 - k1 == 1) continue;
            for (k = k1; k < k2; k++) {
                int up;
                for (p = Up[k]; p < Up[k + 1]; p++) pos[Ui[p]] = p;
                for (p = Lp[k]; p < Lp[k + 1]; p++) pos[Li[p]] = idL + p;
                pos[k] = idD + k;
                for (up = Up[k]; up < Up[k + 1]; up++) {
                    const int j = Ui[up];
                    for (p = Lp[j]; p < Lp[j + 1]; p++) {
                        const int tgt = pos[Li[p]];
                        if (tgt < 0) { ngb_set_error("fill outside the LU pattern (col %d)", k); free(cursor); free(pl); free(pu); goto bad; }
                        if (pass == 0) e_pptr[tgt + 1]++;
                        else {
                            pl[cursor[tgt]] = idL + p; pu[cursor[tgt]] = up; cursor[tgt]++;
                            if (level[idL + p] + 1 > level[tgt]) level[tgt] = level[idL + p] + 1;
                            if (level[up] + 1 > level[tgt]) level[tgt] = level[up] + 1;
                        }
                    }
                }
                if (pass == 1) {
                    /* L entries wait for their pivot */
                    for (p = Lp[k]; p < Lp[k + 1]; p++)
                        if (level[idD + k] + 1 > level[idL + p]) level[idL + p] = level[idD + k] + 1;
                }
                for (p = Up[k]; p < Up[k + 1]; p++) pos[Ui[p]] = -1;
                for (p = Lp[k]; p < Lp[k + 1]; p++) pos[Li[p]] = -1;
                pos[k] = -1;
            }
        }
        free(cursor);
    }
    /* NOTE on level correctness: within column k the U entries are visited in topological
     * order, so when pair (L(i,j), U(j,k)) raises level[target] the level of U(j,k) is final. */
    /* self-check: every operand of an entry must sit on a strictly lower level */
    for (k = 0; k < nV; k++) {
        for (p = e_pptr[k]; p < e_pptr[k + 1]; p++)
            if (level[pl[p]] >= level[k] || level[pu[p]] >= level[k]) { ngb_set_error("LU schedule: operand level violation at entry %d", k); free(pl); free(pu); goto bad; }
        if (e_div[k] >= 0 && level[e_div[k]] >= level[k]) { ngb_set_error("LU schedule: pivot level violation at entry %d", k); free(pl); free(pu); goto bad; }
    }
    for (k = 0; k < nV; k++) if (level[k] + 1 > nlev) nlev = level[k] + 1;
    h->lev_ptr = (int *)xcalloc((size_t)nlev + 1, sizeof(int));
    h->lev_ent = (int *)xcalloc((size_t)nV, sizeof(int));
    {
        int *lp = (int *)h->lev_ptr, *le = (int *)h->lev_ent, *fill;
        for (k = 0; k < nV; k++) lp[level[k] + 1]++;
        for (k = 0; k < nlev; k++) lp[k + 1] += lp[k];
        fill = (int *)xdup(lp, sizeof(int) * ((size_t)nlev + 1));
        for (k = 0; k < nV; k++) le[fill[level[k]]++] = k;
        free(fill);
    }
    h->n = n; h->nnz = c->nnz; h->nV = nV; h->nlev = nlev;
    h->e_aslot = e_aslot; h->e_arow = e_arow; h->e_div = e_div; h->e_pptr = e_pptr;
    h->pair_l = pl; h->pair_u = pu; h->diag_v = diag_v;
    c->npairs = npairs;

    /* CSR view of A for the row scale factors */
    {
        int *rp = (int *)xcalloc((size_t)n + 1, sizeof(int)), *rs = (int *)xcalloc((size_t)c->nnz, sizeof(int)), *fill;
        for (p = 0; p < c->nnz; p++) rp[c->Ai[p] + 1]++;
        for (k = 0; k < n; k++) rp[k + 1] += rp[k];
        fill = (int *)xdup(rp, sizeof(int) * ((size_t)n + 1));
        for (k = 0; k < n; k++) for (p = c->Ap[k]; p < c->Ap[k + 1]; p++) rs[fill[c->Ai[p]]++] = p;
        free(fill);
        h->row_ptr = rp; h->row_slot = rs;
    }

    /* triangular solves: tasks y_i = i, x_i = n + i; pairs appended in the chronological order of
     * klu_solve.c (blocks last to first: L solve, U solve, off-diagonal update) */
    {
        const int ntask = 2 * n;
        int *t_kind = (int *)xcalloc((size_t)ntask, sizeof(int)), *t_init = (int *)xcalloc((size_t)ntask, sizeof(int));
        int *t_div = (int *)xcalloc((size_t)ntask, sizeof(int)), *t_pptr = (int *)xcalloc((size_t)ntask + 1, sizeof(int));
        int *tl = (int *)xcalloc((size_t)ntask, sizeof(int));
        int *t_val = NULL, *t_src = NULL, nsp = 0, nslev = 0;
        for (k = 0; k < n; k++) {
            t_kind[k] = 0; t_init[k] = Pnum[k]; t_div[k] = -1;
            t_kind[n + k] = 1; t_init[n + k] = k; t_div[n + k] = idD + k;
        }
        for (pass = 0; pass < 2; pass++) {
            int *cursor = NULL;
            if (pass == 1) {
                int acc = 0;
                for (k = 0; k < ntask; k++) { int cnt = t_pptr[k + 1]; t_pptr[k] = acc; acc += cnt; }
                t_pptr[ntask] = acc; nsp = acc;
                t_val = (int *)xcalloc((size_t)nsp, sizeof(int)); t_src = (int *)xcalloc((size_t)nsp, sizeof(int));
                cursor = (int *)xdup(t_pptr, sizeof(int) * ((size_t)ntask + 1));
            }
#define SOLVE_PAIR(tgt, val, src) do { if (pass == 0) t_pptr[(tgt) + 1]++; else { \
                t_val[cursor[tgt]] = (val); t_src[cursor[tgt]] = (src); cursor[tgt]++; \
                if (tl[src] + 1 > tl[tgt]) tl[tgt] = tl[src] + 1; } } while (0)
            for (b = nblocks - 1; b >= 0; b--) {
                const int k1 = R[b], k2 = R[b + 1];
                if (k2 - k1 > 1) {
                    for (k = k1; k < k2; k++)
                        for (p = Lp[k]; p < Lp[k + 1]; p++) SOLVE_PAIR(Li[p], idL + p, k);
                    for (k = k2 - 1; k >= k1; k--) {
                        if (pass == 1 && tl[k] + 1 > tl[n + k]) tl[n + k] = tl[k] + 1;     /* x_k starts from y_k */
                        for (p = Up[k]; p < Up[k + 1]; p++) SOLVE_PAIR(n + Ui[p], p, n + k);
                    }
                } else if (pass == 1) {
                    if (tl[k1] + 1 > tl[n + k1]) tl[n + k1] = tl[k1] + 1;
                }
                if (b > 0)
                    for (k = k1; k < k2; k++)
                        for (p = Offp[k]; p < Offp[k + 1]; p++) SOLVE_PAIR(Offi[p], idO + p, n + k);
            }
#undef SOLVE_PAIR
            free(cursor);
        }
        /* The level of x_k must be final before it feeds later rows.  Inside a block the U loop
         * runs k descending and x_k receives pairs only from larger k, and y rows receive L pairs
         * from smaller k before being used: the chronological sweep above is a valid order. */
        for (k = 0; k < ntask; k++) {
            for (p = t_pptr[k]; p < t_pptr[k + 1]; p++)
                if (tl[t_src[p]] >= tl[k]) { ngb_set_error("solve schedule: source level violation at task %d", k); return NGB_E_PANIC; }
            if (t_kind[k] == 1 && tl[t_init[k]] >= tl[k]) { ngb_set_error("solve schedule: x before y at task %d", k); return NGB_E_PANIC; }
        }
        for (k = 0; k < ntask; k++) if (tl[k] + 1 > nslev) nslev = tl[k] + 1;
        {
            int *sp = (int *)xcalloc((size_t)nslev + 1, sizeof(int)), *st = (int *)xcalloc((size_t)ntask, sizeof(int)), *fill;
            for (k = 0; k < ntask; k++) sp[tl[k] + 1]++;
            for (k = 0; k < nslev; k++) sp[k + 1] += sp[k];
            fill = (int *)xdup(sp, sizeof(int) * ((size_t)nslev + 1));
            for (k = 0; k < ntask; k++) st[fill[tl[k]]++] = k;
            free(fill);
            h->slev_ptr = sp; h->slev_task = st;
        }
        free(tl);
        h->ntask = ntask; h->nslev = nslev; h->t_kind = t_kind; h->t_init = t_init; h->t_div = t_div;
        h->t_pptr = t_pptr; h->t_val = t_val; h->t_src = t_src;
        c->nsolvepairs = nsp;
    }
    {
        int *b_eq = (int *)xcalloc((size_t)n, sizeof(int)), *ot = (int *)xcalloc((size_t)n, sizeof(int)), *oe = (int *)xcalloc((size_t)n, sizeof(int));
        for (k = 0; k < n; k++) { b_eq[k] = c->col2eq[k]; ot[k] = n + k; oe[k] = c->col2eq[Q[k]]; }
        h->b_eq = b_eq; h->out_task = ot; h->out_eq = oe;
    }
    free(Pinv); free(pos); free(level);
    h->vchk = lu_pivot_checks(c, n, nblocks, R, Pnum, Lp, Li, unz, nV);
    build_packed(c);
    {   /* move the finished set into its slot */
        struct ngb_luset *L = &c->lu[c->lu_target];
        if (L->valid) { free_sched(&L->sch); free_packed(&L->pk); }
        L->sch = c->sch; L->pk = c->pk; L->npairs = c->npairs; L->nsolvepairs = c->nsolvepairs;
        L->lnz = c->lnz; L->unz = c->unz; L->nzoff = c->nzoff; L->valid = 1;
        L->sig = lu_signature(n, Pnum, Lp, Li, Up, Ui);
        memset(&c->sch, 0, sizeof c->sch); memset(&c->pk, 0, sizeof c->pk);
    }
    c->have_lu = 1;
    return NGB_OK;
bad:
    free(Pinv); free(pos); free(level); free(e_aslot); free(e_arow); free(e_div); free(e_pptr); free(diag_v);
    return NGB_E_PANIC;
}

/* .nodeset (kind 0) and .ic (kind 1) nodes, what CKTic left in CKTnode.nsGiven/nodeset/icGiven/ic: CKTload
 * overrides these rows while the operating point is computed (cktload.c:118-172).  Rows are applied in
 * the order given within each kind, nodesets first, like the two node loops of the reference. */
int ngbCircuitSetNodeOverrides(ngb_circuit *c, int n, const int *eq, const int *kind, const double *value)
{
    int i, k, pass, nz = 0, m = 0;
    if (!c->finalized) { ngb_set_error("ngbCircuitSetNodeOverrides: call ngbCircuitFinalize first"); return NGB_E_PANIC; }
    free(c->ov_eq); free(c->ov_kind); free(c->ov_cur); free(c->ov_diag); free(c->ov_zptr); free(c->ov_zslot); free(c->ov_val);
    c->ov_eq = (int *)xcalloc((size_t)n + 1, sizeof(int)); c->ov_kind = (int *)xcalloc((size_t)n + 1, sizeof(int));
    c->ov_cur = (int *)xcalloc((size_t)n + 1, sizeof(int)); c->ov_diag = (int *)xcalloc((size_t)n + 1, sizeof(int));
    c->ov_zptr = (int *)xcalloc((size_t)n + 2, sizeof(int)); c->ov_val = (double *)xcalloc((size_t)n + 1, sizeof(double));
    c->ov_zslot = (int *)xcalloc((size_t)c->nnz + 1, sizeof(int));
    for (pass = 0; pass < 2; pass++)
        for (i = 0; i < n; i++) {
            int row, col;
            if ((kind[i] ? 1 : 0) != pass) continue;
            if (eq[i] <= 0 || eq[i] > c->neq || c->eq2col[eq[i]] < 0) { ngb_set_error("node override %d: equation %d is not in the matrix", i, eq[i]); return NGB_E_PANIC; }
            row = c->eq2col[eq[i]];
            c->ov_eq[m] = eq[i]; c->ov_kind[m] = pass; c->ov_val[m] = value[i]; c->ov_diag[m] = slot_lookup(c, eq[i], eq[i]);
            /* ZeroNoncurRow: every entry of the row; current-type columns stay (and flag the row) */
            for (col = 0; col < c->n; col++)
                for (k = c->Ap[col]; k < c->Ap[col + 1]; k++)
                    if (c->Ai[k] == row) {
                        if (c->node_type[c->col2eq[col]] == 4 /* SP_CURRENT */) c->ov_cur[m] = 1;
                        else {
                            if (nz >= c->nnz) { int *z = (int *)realloc(c->ov_zslot, sizeof(int) * (size_t)(nz + c->nnz + 1)); if (!z) return NGB_E_PANIC; c->ov_zslot = z; }
                            c->ov_zslot[nz++] = k;
                        }
                    }
            c->ov_zptr[++m] = nz;
        }
    c->ov_n = m;
    return NGB_OK;
}

int ngbCircuitSelectLuSet(ngb_circuit *c, int which)
{
    if (which < 0 || which >= NGB_LU_SETS) return NGB_E_PANIC;
    c->lu_target = which;
    return NGB_OK;
}
int ngbCircuitSetLuEvents(ngb_circuit *c, const int *set_of_event)
{
    int e;
    for (e = 0; e < NGB_LU_EVENTS; e++) {
        if (set_of_event[e] < 0 || set_of_event[e] >= NGB_LU_SETS || !c->lu[set_of_event[e]].valid) {
            ngb_set_error("pivoting event %d refers to pattern set %d, which is not filled", e, set_of_event[e]);
            return NGB_E_PANIC;
        }
        c->lu_event[e] = set_of_event[e];
    }
    c->lu_event_set = 1;
    return NGB_OK;
}
/* without ngbCircuitSetLuEvents: set 0 serves the operating point, set 1 (when filled) the transient */
void ngb_lu_events(const ngb_circuit *c, int ev[NGB_LU_EVENTS])
{
    int e;
    for (e = 0; e < NGB_LU_EVENTS; e++)
        ev[e] = c->lu_event_set ? c->lu_event[e] : ((e >= 2 && c->lu[1].valid) ? 1 : (c->lu[0].valid ? 0 : 1));
}
int ngbCircuitLuInfo(const ngb_circuit *c, int info[9])
{
    const struct ngb_luset *L = &c->lu[c->lu_target];
    if (!L->valid) return NGB_E_PANIC;
    info[0] = L->sch.nV; info[1] = L->sch.nlev; info[2] = L->npairs; info[3] = L->sch.ntask;
    info[4] = L->sch.nslev; info[5] = L->nsolvepairs; info[6] = L->lnz; info[7] = L->unz; info[8] = L->nzoff;
    return NGB_OK;
}

/* ------------------------------------------------------------------ batch */
static void *dev_dup(const void *host, size_t bytes)
{
    void *d = ngb_dev_malloc(bytes ? bytes : 8);
    if (d && bytes) ngb_dev_h2d(d, host, bytes);
    return d;
}
static void reg(ngb_batch *b, const char *name, void *ptr, size_t bytes)
{
    if (b->narr < NGB_MAX_ARR) { b->arr[b->narr].name = name; b->arr[b->narr].ptr = ptr; b->arr[b->narr].bytes = bytes; b->narr++; }
}
static void *dalloc(ngb_batch *b, const char *name, size_t bytes)
{
    void *p = ngb_dev_malloc(bytes ? bytes : 8);
    if (!p) { ngb_set_error("device allocation of %zu bytes for %s failed", bytes, name); b->failed = 1; return NULL; }
    reg(b, name, p, bytes);
    return p;
}
/* replicate a [nf][n] host table to a [nf][n*S] device table (sample index fastest) */
static void *dalloc_rep(ngb_batch *b, const char *name, const double *host, int nf, int n, int S)
{
    size_t T = (size_t)n * S, i; int f, s;
    double *tmp = (double *)xcalloc((size_t)nf * T, sizeof(double)), *d;
    for (f = 0; f < nf; f++) for (i = 0; i < (size_t)n; i++) for (s = 0; s < S; s++)
        tmp[(size_t)f * T + i * S + s] = host[(size_t)f * n + i];
    d = (double *)dalloc(b, name, sizeof(double) * (size_t)nf * T);
    if (d) ngb_dev_h2d(d, tmp, sizeof(double) * (size_t)nf * T);
    free(tmp);
    return d;
}

static void sched_to_dev(ngb_batch *b, const ngb_circuit *cc, int w)
{
    const struct ngb_luset *c = &cc->lu[w];
    const NgbLuSched *h = &c->sch; NgbLuSched *d = &b->dlu[w].dsch;
    *d = *h;
#define D(f, cnt) d->f = (const int *)dev_dup(h->f, sizeof(int) * (size_t)(cnt))
    D(lev_ptr, h->nlev + 1); D(lev_ent, h->nV); D(e_aslot, h->nV); D(e_arow, h->nV); D(e_div, h->nV);
    D(e_pptr, h->nV + 1); D(pair_l, c->npairs); D(pair_u, c->npairs); D(diag_v, h->n);
    D(row_ptr, h->n + 1); D(row_slot, h->nnz); D(slev_ptr, h->nslev + 1); D(slev_task, h->ntask);
    D(t_kind, h->ntask); D(t_init, h->ntask); D(t_div, h->ntask); D(t_pptr, h->ntask + 1);
    D(t_val, c->nsolvepairs); D(t_src, c->nsolvepairs); D(b_eq, h->n); D(out_task, h->n); D(out_eq, h->n);
    if (h->vchk) D(vchk, h->nV);
#undef D
}
static void packed_to_dev(ngb_batch *b, const ngb_circuit *c, int w)
{
    const NgbLuPacked *p = &c->lu[w].pk; NgbLuPacked *d = &b->dlu[w].dpk;
    *d = *p;
    if (!p->ok) return;
    d->blob = (const unsigned short *)dev_dup(p->blob, sizeof(unsigned short) * (size_t)p->blob_u16);
    d->aslot = (const int *)dev_dup(p->aslot, sizeof(int) * (size_t)p->nV);
    d->arow = (const int *)dev_dup(p->arow, sizeof(int) * (size_t)p->nV);
    d->ext = (const int *)dev_dup(p->ext, sizeof(int) * (size_t)p->nV);
    if (p->ok2) d->blob2 = (const unsigned short *)dev_dup(p->blob2, sizeof(unsigned short) * (size_t)p->blob2_u16);
    d->row_ptr = b->dlu[w].dsch.row_ptr; d->row_slot = b->dlu[w].dsch.row_slot; d->b_eq = b->dlu[w].dsch.b_eq; d->out_eq = b->dlu[w].dsch.out_eq;
}
static void sched_dev_free(NgbLuSched *d)
{
#define F(p) ngb_dev_free((void *)d->p)
    F(lev_ptr); F(lev_ent); F(e_aslot); F(e_arow); F(e_div); F(e_pptr); F(pair_l); F(pair_u);
    F(diag_v); F(row_ptr); F(row_slot); F(slev_ptr); F(slev_task); F(t_kind); F(t_init); F(t_div);
    F(t_pptr); F(t_val); F(t_src); F(b_eq); F(out_task); F(out_eq); F(vchk);
#undef F
    memset(d, 0, sizeof *d);
}

ngb_batch *ngbBatchCreate(ngb_circuit *c, int S, int device)
{
    ngb_batch *b;
    NgbCtl *k;
    int i;
    if (!c->finalized) { ngb_set_error("circuit not finalized"); return NULL; }
    if (ngb_dev_init(device) != 0) { ngb_set_error("CUDA device %d could not be initialised (no CPU fallback)", device); return NULL; }
    b = (ngb_batch *)xcalloc(1, sizeof *b);
    b->c = c; b->S = S; b->neq1 = c->neq + 1;
    k = &b->ctl; k->S = S;
    k->mode = (int *)dalloc(b, "ctl.mode", sizeof(int) * (size_t)S);
    k->active = (int *)dalloc(b, "ctl.active", sizeof(int) * (size_t)S);
    k->head = (int *)dalloc(b, "ctl.head", sizeof(int) * (size_t)S);
    k->order = (int *)dalloc(b, "ctl.order", sizeof(int) * (size_t)S);
    k->noncon = (int *)dalloc(b, "ctl.noncon", sizeof(int) * (size_t)S);
    k->xsel = (int *)dalloc(b, "ctl.xsel", sizeof(int) * (size_t)S);
    k->err = (int *)dalloc(b, "ctl.err", sizeof(int) * (size_t)S);
    k->ag0 = (double *)dalloc(b, "ctl.ag0", sizeof(double) * (size_t)S);
    k->ag1 = (double *)dalloc(b, "ctl.ag1", sizeof(double) * (size_t)S);
    k->ag2 = (double *)dalloc(b, "ctl.ag2", sizeof(double) * (size_t)S);
    if (k->ag2) ngb_dev_memset(k->ag2, 0, sizeof(double) * (size_t)S);
    k->delta = (double *)dalloc(b, "ctl.delta", sizeof(double) * (size_t)S);
    k->delta_old = (double *)dalloc(b, "ctl.delta_old", sizeof(double) * 7 * (size_t)S);
    k->time = (double *)dalloc(b, "ctl.time", sizeof(double) * (size_t)S);
    k->gmin = (double *)dalloc(b, "ctl.gmin", sizeof(double) * (size_t)S);
    k->diag_gmin = (double *)dalloc(b, "ctl.diag_gmin", sizeof(double) * (size_t)S);
    k->srcfact = (double *)dalloc(b, "ctl.srcfact", sizeof(double) * (size_t)S);
    k->lte = (double *)dalloc(b, "ctl.lte", sizeof(double) * (size_t)S);
    k->lte2 = (double *)dalloc(b, "ctl.lte2", sizeof(double) * (size_t)S);
    k->stateop = (int *)dalloc(b, "ctl.stateop", sizeof(int) * (size_t)S);
    k->lusel = (int *)dalloc(b, "ctl.lusel", sizeof(int) * (size_t)S);
    if (k->lusel) ngb_dev_memset(k->lusel, 0, sizeof(int) * (size_t)S);
    k->nhist = c->opt.maxorder + 2;
    if (k->nhist > NGB_NHIST) k->nhist = NGB_NHIST;
    k->gear = c->opt.method == NGB_GEAR;
    k->reltol = c->opt.reltol; k->abstol = c->opt.abstol; k->chgtol = c->opt.chgtol; k->trtol = c->opt.trtol;
    {
        int *one = (int *)xcalloc((size_t)S, sizeof(int)); double *dv = (double *)xcalloc((size_t)S, sizeof(double));
        for (i = 0; i < S; i++) one[i] = 1;
        ngb_dev_h2d(k->active, one, sizeof(int) * (size_t)S);
        ngb_dev_h2d(k->order, one, sizeof(int) * (size_t)S);
        for (i = 0; i < S; i++) dv[i] = c->opt.gmin;
        ngb_dev_h2d(k->gmin, dv, sizeof(double) * (size_t)S);
        for (i = 0; i < S; i++) dv[i] = 1.0;
        ngb_dev_h2d(k->srcfact, dv, sizeof(double) * (size_t)S);
        for (i = 0; i < S; i++) dv[i] = 1e300;
        ngb_dev_h2d(k->lte, dv, sizeof(double) * (size_t)S);
        ngb_dev_h2d(k->lte2, dv, sizeof(double) * (size_t)S);
        free(one); free(dv);
    }
    b->x = (double *)dalloc(b, "x", sizeof(double) * 2 * (size_t)b->neq1 * S);
    b->Ax = (double *)dalloc(b, "Ax", sizeof(double) * (size_t)c->nnz * S);
    b->stamp = (double *)dalloc(b, "stamp", sizeof(double) * (size_t)c->nstamp_rows * S);
    b->errflag = (int *)dalloc(b, "errflag", sizeof(int) * 4);
    b->d_node_type = (int *)dev_dup(c->node_type, sizeof(int) * (size_t)b->neq1);
    b->d_tgt_ptr = (int *)dev_dup(c->tgt_ptr, sizeof(int) * ((size_t)c->ntgt + 1));
    b->d_tgt_rows = (int *)dev_dup(c->tgt_rows, sizeof(int) * (size_t)c->tgt_ptr[c->ntgt]);
    b->d_slot_diag = (int *)dev_dup(c->slot_diag, sizeof(int) * (size_t)c->nnz);
    b->d_long_tgt = (int *)dev_dup(c->long_tgt, sizeof(int) * (size_t)c->nlong);
    if (c->nlong) {
        int k2, mx = 0;
        b->long_len = (int *)xcalloc((size_t)c->nlong, sizeof(int));
        for (k2 = 0; k2 < c->nlong; k2++) {
            b->long_len[k2] = c->tgt_ptr[c->long_tgt[k2] + 1] - c->tgt_ptr[c->long_tgt[k2]];
            if ((b->long_len[k2] + NGB_ASM_CHUNK - 1) / NGB_ASM_CHUNK > mx) mx = (b->long_len[k2] + NGB_ASM_CHUNK - 1) / NGB_ASM_CHUNK;
        }
        b->long_cap = mx * S;
        b->long_part = (double *)ngb_dev_malloc(sizeof(double) * 2 * (size_t)b->long_cap);
    }
    /* constant stamp rows (resistors, source incidence): written once */
    if (c->nconst) {
        double *row = (double *)xcalloc((size_t)S, sizeof(double));
        for (i = 0; i < c->nconst; i++) {
            int s; for (s = 0; s < S; s++) row[s] = c->const_val[i];
            ngb_dev_h2d(b->stamp + (size_t)c->const_row[i] * S, row, sizeof(double) * (size_t)S);
        }
        free(row);
    }
    if (c->b4_n) {
        const size_t T = (size_t)c->b4_n * S;
        b->b4_inst = (double *)dalloc_rep(b, "b4.inst", c->b4_inst, B4I_COUNT, c->b4_n, S);
        b->b4_state = (double *)dalloc(b, "b4.state", sizeof(double) * NGB_NHIST * B4ST_COUNT * T);
        b->b4_op = (double *)dalloc(b, "b4.op", sizeof(double) * B4O_COUNT * T);
        b->b4_mtab = (double *)dev_dup(c->b4_mtab, sizeof(double) * (size_t)c->b4_nrows * B4M_COUNT);
        b->b4_key = b4_batch_key(c, c->b4_mtab, c->b4_nrows, c->b4_prow);
        { const char *e = getenv("NGB_B4_GENERIC"); b->b4_force_generic = (e && atoi(e)) ? 1 : 0; }
        b->b4_ptab = (double *)dev_dup(c->b4_ptab, sizeof(double) * (size_t)c->b4_nrows * B4P_COUNT);
        reg(b, "b4.mtab", b->b4_mtab, sizeof(double) * (size_t)c->b4_nrows * B4M_COUNT);
        reg(b, "b4.ptab", b->b4_ptab, sizeof(double) * (size_t)c->b4_nrows * B4P_COUNT);
        b->b4_prow = (int *)dev_dup(c->b4_prow, sizeof(int) * (size_t)c->b4_n);
        b->b4_flags = (int *)dev_dup(c->b4_flags, sizeof(int) * (size_t)c->b4_n);
        b->b4_nodes = (int *)dev_dup(c->b4_nodes, sizeof(int) * (size_t)c->b4_n * B4N_COUNT);
        b->b4_spos = (int *)dev_dup(c->b4_spos, sizeof(int) * (size_t)c->b4_n * B4S_TOTAL);
    }
    if (c->b3_n) {
        const size_t T = (size_t)c->b3_n * S;
        b->b3_inst = (double *)dalloc_rep(b, "b3.inst", c->b3_inst, B3I_COUNT, c->b3_n, S);
        b->b3_state = (double *)dalloc(b, "b3.state", sizeof(double) * NGB_NHIST * B3ST_COUNT * T);
        b->b3_von = (double *)dalloc(b, "b3.von", sizeof(double) * T);
        b->b3_mtab = (double *)dev_dup(c->b3_mtab, sizeof(double) * (size_t)c->b3_nrows * B3M_COUNT);
        b->b3_ptab = (double *)dev_dup(c->b3_ptab, sizeof(double) * (size_t)c->b3_nrows * B3P_COUNT);
        b->b3_prow = (int *)dev_dup(c->b3_prow, sizeof(int) * (size_t)c->b3_n);
        b->b3_flags = (int *)dev_dup(c->b3_flags, sizeof(int) * (size_t)c->b3_n);
        b->b3_nodes = (int *)dev_dup(c->b3_nodes, sizeof(int) * (size_t)c->b3_n * B3N_COUNT);
        b->b3_spos = (int *)dev_dup(c->b3_spos, sizeof(int) * (size_t)c->b3_n * B3S_COUNT);
    }
    if (c->cap_n) {
        const size_t T = (size_t)c->cap_n * S;
        b->cap_par = (double *)dalloc_rep(b, "cap.par", c->cap_par, 3, c->cap_n, S);
        b->cap_state = (double *)dalloc(b, "cap.state", sizeof(double) * NGB_NHIST * 2 * T);
        b->cap_nodes = (int *)dev_dup(c->cap_nodes, sizeof(int) * 2 * (size_t)c->cap_n);
        b->cap_spos = (int *)dev_dup(c->cap_spos, sizeof(int) * 6 * (size_t)c->cap_n);
    }
    if (c->dio_n) {
        const size_t T = (size_t)c->dio_n * S;
        b->dio_par = (double *)dalloc_rep(b, "dio.par", c->dio_par, DIOP_COUNT, c->dio_n, S);
        b->dio_state = (double *)dalloc(b, "dio.state", sizeof(double) * NGB_NHIST * DIOST_COUNT * T);
        b->dio_nodes = (int *)dev_dup(c->dio_nodes, sizeof(int) * DION_COUNT * (size_t)c->dio_n);
        b->dio_flags = (int *)dev_dup(c->dio_flags, sizeof(int) * (size_t)c->dio_n);
        b->dio_spos = (int *)dev_dup(c->dio_spos, sizeof(int) * DIOS_COUNT * (size_t)c->dio_n);
    }
    if (c->ov_n) {
        b->ov_val = (double *)dalloc_rep(b, "node.override", c->ov_val, 1, c->ov_n, S);
        b->ov_eq = (int *)dev_dup(c->ov_eq, sizeof(int) * (size_t)c->ov_n); b->ov_kind = (int *)dev_dup(c->ov_kind, sizeof(int) * (size_t)c->ov_n);
        b->ov_cur = (int *)dev_dup(c->ov_cur, sizeof(int) * (size_t)c->ov_n); b->ov_diag = (int *)dev_dup(c->ov_diag, sizeof(int) * (size_t)c->ov_n);
        b->ov_zptr = (int *)dev_dup(c->ov_zptr, sizeof(int) * ((size_t)c->ov_n + 1));
        b->ov_zslot = (int *)dev_dup(c->ov_zslot, sizeof(int) * ((size_t)c->ov_zptr[c->ov_n] + 1));
    }
    if (c->vb_n) {
        const size_t T = (size_t)c->vb_n * S;
        b->vb_par = (double *)dalloc_rep(b, "vbic.par", c->vb_par, VBIC_NP, c->vb_n, S);
        b->vb_aux = (double *)dalloc_rep(b, "vbic.aux", c->vb_aux, VBA_COUNT, c->vb_n, S);
        b->vb_state = (double *)dalloc(b, "vbic.state", sizeof(double) * NGB_NHIST * VBS_COUNT * T);
        b->vb_nodes = (int *)dev_dup(c->vb_nodes, sizeof(int) * VBN_COUNT * (size_t)c->vb_n);
        b->vb_flags = (int *)dev_dup(c->vb_flags, sizeof(int) * (size_t)c->vb_n);
        b->vb_spos = (int *)dev_dup(c->vb_spos, sizeof(int) * VBIC_NSTAMPS * (size_t)c->vb_n);
    }
    if (c->vs_n) {
        b->vs_par = (double *)dalloc_rep(b, "vsrc.par", c->vs_par, 9, c->vs_n, S);
        b->vs_fn = (int *)dev_dup(c->vs_fn, sizeof(int) * 3 * (size_t)c->vs_n);
        if (c->vs_pwl_ptr) {
            b->vs_pwl_ptr = (int *)dev_dup(c->vs_pwl_ptr, sizeof(int) * ((size_t)c->vs_n + 1));
            b->vs_pwl_rep = (int *)dev_dup(c->vs_pwl_rep, sizeof(int) * (size_t)c->vs_n);
            b->vs_pwl_rdelay = (double *)dev_dup(c->vs_pwl_rdelay, sizeof(double) * (size_t)c->vs_n);
            b->vs_pwl = (double *)dev_dup(c->vs_pwl, sizeof(double) * (size_t)c->vs_pwl_n);
        }
        b->vs_spos = (int *)dev_dup(c->vs_spos, sizeof(int) * (size_t)c->vs_n);
    }
    if (c->is_n) {
        b->is_par = (double *)dalloc_rep(b, "isrc.par", c->is_par, 10, c->is_n, S);
        b->is_fn = (int *)dev_dup(c->is_fn, sizeof(int) * 3 * (size_t)c->is_n);
        b->is_spos = (int *)dev_dup(c->is_spos, sizeof(int) * 2 * (size_t)c->is_n);
        if (c->is_pwl_ptr) {
            b->is_pwl_ptr = (int *)dev_dup(c->is_pwl_ptr, sizeof(int) * ((size_t)c->is_n + 1));
            b->is_pwl = (double *)dev_dup(c->is_pwl, sizeof(double) * (size_t)c->is_pwl_n);
        }
    }
    if (c->have_lu) {
        int w, nVmax = 0;
        for (w = 0; w < NGB_LU_SETS; w++)
            if (c->lu[w].valid) {
                sched_to_dev(b, c, w); packed_to_dev(b, c, w); b->dlu[w].valid = 1;
                if (c->lu[w].sch.nV > nVmax) nVmax = c->lu[w].sch.nV;
                if (c->lu[w].sch.ntask > b->lu_ntask_cap) b->lu_ntask_cap = c->lu[w].sch.ntask;
            }
        b->lu_which = c->lu[0].valid ? 0 : 1;
        b->V = (double *)dalloc(b, "lu.V", sizeof(double) * (size_t)nVmax * S);
        b->Rs = (double *)dalloc(b, "lu.Rs", sizeof(double) * (size_t)c->n * S);
        b->Zw = (double *)dalloc(b, "lu.Z", sizeof(double) * (size_t)b->lu_ntask_cap * S);
        b->nodeconv = (int *)dalloc(b, "lu.nodeconv", sizeof(int) * (size_t)S);
        b->singular = (int *)dalloc(b, "lu.singular", sizeof(int) * (size_t)S);
        b->have_lu = 1;
    }
    if (b->failed) { ngbBatchDestroy(b); return NULL; }
    ngb_dev_sync();
    return b;
}

static void batch_free_lu(ngb_batch *b)
{
    int w;
    for (w = 0; w < NGB_LU_SETS; w++)
        if (b->dlu[w].valid) {
            if (b->dlu[w].dpk.ok2) ngb_dev_free((void *)b->dlu[w].dpk.blob2);
            if (b->dlu[w].dpk.ok) { ngb_dev_free((void *)b->dlu[w].dpk.blob); ngb_dev_free((void *)b->dlu[w].dpk.aslot);
                                    ngb_dev_free((void *)b->dlu[w].dpk.arow); ngb_dev_free((void *)b->dlu[w].dpk.ext); }
            sched_dev_free(&b->dlu[w].dsch);
            b->dlu[w].valid = 0;
        }
}

/* The reference's answer to a zero pivot in a refactor (niiter.c:162-195), and what it does at every pivoting event of a
 * run: factor the sample's OWN matrix with pivoting.  ngb_repivot_compute runs the pivoting factor (ngb_pivot.c) on the
 * circuit's symbolic analysis; ngb_repivot_commit matches the result against the pattern sets already there or makes
 * it a new one, uploaded for this batch.  commit returns 0 with *set_out, E_SINGULAR when the matrix really is singular,
 * E_UNSUPP when there is no symbolic analysis to factor on or no free slot. */
void ngb_repivot_compute(const ngb_circuit *c, const double *Ax, NgbRepivot *r)
{
    const int n = c->n;
    memset(r, 0, sizeof *r);
    r->sing = -1;
    if (!c->klu_P) { r->rc = NGB_E_UNSUPP; return; }
    r->Pnum = (int *)malloc(sizeof(int) * (size_t)n); r->Lp = (int *)malloc(sizeof(int) * ((size_t)n + 1));
    r->Up = (int *)malloc(sizeof(int) * ((size_t)n + 1)); r->Offp = (int *)malloc(sizeof(int) * ((size_t)n + 1));
    if (!r->Pnum || !r->Lp || !r->Up || !r->Offp) { r->rc = NGB_E_PANIC; return; }
    r->rc = ngb_pivot_factor(n, c->Ap, c->Ai, Ax, c->klu_nblocks, c->klu_P, c->klu_Q, c->klu_R, c->pivtol > 0 ? c->pivtol : 0.001,
                             r->Pnum, r->Lp, &r->Li, r->Up, &r->Ui, r->Offp, &r->Offi, &r->sing);
}
void ngb_repivot_free(NgbRepivot *r)
{
    free(r->Pnum); free(r->Lp); free(r->Up); free(r->Offp); free(r->Li); free(r->Ui); free(r->Offi);
    memset(r, 0, sizeof *r);
}
int ngb_repivot_commit(ngb_batch *b, NgbRepivot *r, int *set_out)
{
    ngb_circuit *c = b->c;
    const int n = c->n, keep_target = c->lu_target;
    int w, rc;
    unsigned long long sig;
    if (r->rc) return r->rc;
    sig = lu_signature(n, r->Pnum, r->Lp, r->Li, r->Up, r->Ui);
    for (w = 0; w < NGB_LU_SETS; w++) if (c->lu[w].valid && c->lu[w].sig == sig) break;
    if (w == NGB_LU_SETS) {
        int *Q, *R;
        for (w = NGB_LU_EVENTS; w < NGB_LU_SETS; w++) if (!c->lu[w].valid) break;
        if (w == NGB_LU_SETS) { ngb_set_error("a sample needs one pivot order more than the %d pattern sets of a circuit", NGB_LU_SETS); return NGB_E_UNSUPP; }
        Q = (int *)xdup(c->klu_Q, sizeof(int) * (size_t)n); R = (int *)xdup(c->klu_R, sizeof(int) * ((size_t)c->klu_nblocks + 1));
        c->lu_target = w;
        rc = ngbCircuitSetLuPattern(c, n, c->klu_nblocks, Q, R, r->Pnum, r->Lp, r->Li, r->Up, r->Ui, r->Offp, r->Offi);
        c->lu_target = keep_target;
        free(Q); free(R);
        if (rc) return rc;
    }
    if (!b->dlu[w].valid) { sched_to_dev(b, c, w); packed_to_dev(b, c, w); b->dlu[w].valid = 1; if (b->failed) return NGB_E_PANIC; }
    *set_out = w;
    return NGB_OK;
}

/* 0: every sample refactors on the batch's pattern sets (one per pivoting event of the recorded run); 1: every sample's
 * own matrix is factored with pivoting at the reference's pivoting events (niiter.c:107-111, 335, 343) -- needs the
 * symbolic analysis; -1 (default): 1 when there is one */
int ngbCircuitSetPivotMode(ngb_circuit *c, int mode)
{
    if (mode == 1 && !c->klu_P) { ngb_set_error("per-sample pivoting needs a symbolic analysis (ngbCircuitSetSymbolic / ngbCircuitAnalyze)"); return NGB_E_PANIC; }
    c->pivot_mode = mode;
    return NGB_OK;
}

/* the circuit's LU pattern sets changed after the batch was created (a later SMPreorder in the
 * reference-side shim): upload them again */
int ngbBatchRefreshLu(ngb_batch *b)
{
    const ngb_circuit *c = b->c;
    const int S = b->S;
    int w, nVmax = 0, ntmax = 0;
    if (!c->have_lu) { ngb_set_error("no LU pattern on the circuit"); return NGB_E_PANIC; }
    ngb_dev_sync();
    batch_free_lu(b);
    for (w = 0; w < NGB_LU_SETS; w++)
        if (c->lu[w].valid) {
            sched_to_dev(b, c, w); packed_to_dev(b, c, w); b->dlu[w].valid = 1;
            if (c->lu[w].sch.nV > nVmax) nVmax = c->lu[w].sch.nV;
            if (c->lu[w].sch.ntask > ntmax) ntmax = c->lu[w].sch.ntask;
        }
    b->lu_which = c->lu[0].valid ? 0 : 1;
    if (!b->have_lu || (size_t)nVmax * S * sizeof(double) > (size_t)ngbBatchArrayBytes(b, "lu.V") || ntmax > b->lu_ntask_cap) {
        int i;
        for (i = 0; i < b->narr; i++)                       /* drop the old, smaller work arrays from the registry */
            if (!strcmp(b->arr[i].name, "lu.V") || !strcmp(b->arr[i].name, "lu.Rs") || !strcmp(b->arr[i].name, "lu.nodeconv") ||
                !strcmp(b->arr[i].name, "lu.singular") || !strcmp(b->arr[i].name, "lu.Z")) { ngb_dev_free(b->arr[i].ptr); b->arr[i] = b->arr[--b->narr]; i--; }
        if (ntmax > b->lu_ntask_cap) b->lu_ntask_cap = ntmax;
        b->V = (double *)dalloc(b, "lu.V", sizeof(double) * (size_t)nVmax * S);
        b->Rs = (double *)dalloc(b, "lu.Rs", sizeof(double) * (size_t)c->n * S);
        b->Zw = (double *)dalloc(b, "lu.Z", sizeof(double) * (size_t)b->lu_ntask_cap * S);
        b->nodeconv = (int *)dalloc(b, "lu.nodeconv", sizeof(int) * (size_t)S);
        b->singular = (int *)dalloc(b, "lu.singular", sizeof(int) * (size_t)S);
    }
    b->have_lu = 1;
    return b->failed ? NGB_E_PANIC : NGB_OK;
}

void ngbBatchDestroy(ngb_batch *b)
{
    int i;
    if (!b) return;
    free(b->ms_eq); free(b->ms_kind); free(b->ms_count); free(b->ms_val); free(b->ms_td);
    for (i = 0; i < b->narr; i++)
        if (strcmp(b->arr[i].name, "b4.mtab") && strcmp(b->arr[i].name, "b4.ptab")) ngb_dev_free(b->arr[i].ptr);
    free(b->long_len); ngb_dev_free(b->long_part);
    ngb_dev_free(b->d_node_type); ngb_dev_free(b->d_tgt_ptr); ngb_dev_free(b->d_tgt_rows); ngb_dev_free(b->d_slot_diag); ngb_dev_free(b->d_long_tgt);
    if (b->b4_rows_block) { ngb_dev_l2_persist(NULL, 0); ngb_dev_free(b->b4_rows_block); } else { ngb_dev_free(b->b4_mtab); ngb_dev_free(b->b4_ptab); }
    ngb_dev_free(b->b4_prow); ngb_dev_free(b->b4_flags);
    ngb_dev_free(b->b4_nodes); ngb_dev_free(b->b4_spos); ngb_dev_free(b->b4_prow_t);
    ngb_dev_free(b->cap_nodes); ngb_dev_free(b->cap_spos);
    ngb_dev_free(b->dio_nodes); ngb_dev_free(b->dio_flags); ngb_dev_free(b->dio_spos);
    ngb_dev_free(b->vb_nodes); ngb_dev_free(b->vb_flags); ngb_dev_free(b->vb_spos);
    ngb_dev_free(b->ov_eq); ngb_dev_free(b->ov_kind); ngb_dev_free(b->ov_cur); ngb_dev_free(b->ov_diag); ngb_dev_free(b->ov_zptr); ngb_dev_free(b->ov_zslot);
    ngb_dev_free(b->b3_mtab); ngb_dev_free(b->b3_ptab); ngb_dev_free(b->b3_prow); ngb_dev_free(b->b3_flags); ngb_dev_free(b->b3_nodes); ngb_dev_free(b->b3_spos);
    ngb_dev_free(b->vs_pwl_ptr); ngb_dev_free(b->vs_pwl_rep); ngb_dev_free(b->vs_pwl_rdelay); ngb_dev_free(b->vs_pwl);
    ngb_dev_free(b->is_pwl_ptr); ngb_dev_free(b->is_pwl);
    ngb_dev_free(b->vs_fn); ngb_dev_free(b->vs_spos); ngb_dev_free(b->is_fn); ngb_dev_free(b->is_spos);
    if (b->have_lu) {
        batch_free_lu(b);
    }
    ngb_tran_free(b);
    free(b);
}

static int find_arr(ngb_batch *b, const char *name)
{
    int i;
    for (i = 0; i < b->narr; i++) if (!strcmp(b->arr[i].name, name)) return i;
    ngb_set_error("no device array named '%s'", name);
    return -1;
}
long ngbBatchArrayBytes(ngb_batch *b, const char *name) { int i = find_arr(b, name); return i < 0 ? -1 : (long)b->arr[i].bytes; }
void *ngbBatchDevPtr(ngb_batch *b, const char *name) { int i = find_arr(b, name); return i < 0 ? NULL : b->arr[i].ptr; }
int ngbBatchUpload(ngb_batch *b, const char *name, const void *host, long bytes, long offset)
{
    int i = find_arr(b, name);
    if (i < 0) return NGB_E_PANIC;
    if (offset < 0 || (size_t)(offset + bytes) > b->arr[i].bytes) { ngb_set_error("upload to %s out of range", name); return NGB_E_PANIC; }
    return ngb_dev_h2d((char *)b->arr[i].ptr + offset, host, (size_t)bytes);
}
int ngbBatchDownload(ngb_batch *b, const char *name, void *host, long bytes, long offset)
{
    int i = find_arr(b, name);
    if (i < 0) return NGB_E_PANIC;
    if (offset < 0 || (size_t)(offset + bytes) > b->arr[i].bytes) { ngb_set_error("download from %s out of range", name); return NGB_E_PANIC; }
    return ngb_dev_d2h(host, (char *)b->arr[i].ptr + offset, (size_t)bytes);
}
int ngbBatchSetOpFull(ngb_batch *b, int on) { b->op_full = on; return NGB_OK; }

/* per-sample resistor conductances (parameter sweeps): g [nres][S], what REStemp leaves in RESconduct */
int ngbBatchSetResistors(ngb_batch *b, const double *g)
{
    const ngb_circuit *c = b->c;
    const int S = b->S;
    double *neg = (double *)xcalloc((size_t)S, sizeof(double));
    int i, k, s, rc = NGB_OK;
    for (i = 0; i < c->res_n && !rc; i++) {
        for (s = 0; s < S; s++) neg[s] = -g[(size_t)i * S + s];
        for (k = 0; k < 4 && !rc; k++) {
            const int r = c->res_spos[k * c->res_n + i];
            if (r >= 0) rc = ngb_dev_h2d(b->stamp + (size_t)r * S, k < 2 ? g + (size_t)i * S : neg, sizeof(double) * (size_t)S);
        }
    }
    free(neg);
    return rc;
}

/* per-thread parameter rows (Monte-Carlo with model-parameter mismatch): prow [ninst*S] */
/* which columns of the per-sample rows differ between the samples of an instance: the load reads only those at the thread's
 * own row and every other one at the row of the instance's sample 0 (the same value, at a warp-uniform address).  Row pairs are
 * compared once per distinct (row of sample 0, row of sample s) pair of consecutive instances -- the instances of a card share
 * their rows -- so the cost is one pass over the table, not over the threads */
static void b4_find_vary(ngb_batch *b, const int *prow_t, const double *mtab, const double *ptab)
{
    const int n = b->c->b4_n, S = b->S;
    int i, s, f, seen[16], nseen = 0;
    memset(b->b4_mvary, 0, sizeof b->b4_mvary);
    memset(b->b4_pvary, 0, sizeof b->b4_pvary);
    for (i = 0; i < n; i++) {
        const int *pr = prow_t + (size_t)i * S;
        {   /* same rows as an instance already looked at (the transistors of one card) */
            int k, dup = 0;
            for (k = 0; k < nseen && !dup; k++) dup = !memcmp(pr, prow_t + (size_t)seen[k] * S, sizeof(int) * (size_t)S);
            if (dup) continue;
            if (nseen < 16) seen[nseen++] = i;
        }
        for (s = 1; s < S; s++) {
            const double *m0 = mtab + (size_t)pr[0] * B4M_COUNT, *m1 = mtab + (size_t)pr[s] * B4M_COUNT;
            const double *p0 = ptab + (size_t)pr[0] * B4P_COUNT, *p1 = ptab + (size_t)pr[s] * B4P_COUNT;
            if (pr[s] == pr[0]) continue;
            /* bit patterns, not values: -0.0 and NaN payloads are different parameters too */
            for (f = 0; f < B4M_COUNT; f++) if (memcmp(&m0[f], &m1[f], sizeof(double))) b->b4_mvary[f] = (int)(sizeof(double) * B4M_COUNT);
            for (f = 0; f < B4P_COUNT; f++) if (memcmp(&p0[f], &p1[f], sizeof(double))) b->b4_pvary[f] = (int)(sizeof(double) * B4P_COUNT);
        }
    }
}

int ngbBatchSetBsim4Rows(ngb_batch *b, const int *prow_t, int nrows, const double *mtab, const double *ptab)
{
    const size_t T = (size_t)b->c->b4_n * b->S;
    /* both tables in one block: every row is read by all the instances of its sample, in different CTAs at different times,
     * while each load streams ~330 MB of states and stamps through the 126 MB L2 -- the block is marked L2-persisting */
    const size_t mb = (sizeof(double) * (size_t)nrows * B4M_COUNT + 255) & ~(size_t)255, pb = sizeof(double) * (size_t)nrows * B4P_COUNT;
    ngb_dev_free(b->b4_prow_t);
    if (b->b4_rows_block) ngb_dev_free(b->b4_rows_block); else { ngb_dev_free(b->b4_mtab); ngb_dev_free(b->b4_ptab); }
    b->b4_prow_t = (int *)dev_dup(prow_t, sizeof(int) * T);
    b->b4_rows_block = ngb_dev_malloc(mb + pb);
    b->b4_mtab = (double *)b->b4_rows_block; b->b4_ptab = b->b4_rows_block ? (double *)((char *)b->b4_rows_block + mb) : NULL;
    if (!b->b4_prow_t || !b->b4_rows_block) return NGB_E_PANIC;
    ngb_dev_h2d(b->b4_mtab, mtab, sizeof(double) * (size_t)nrows * B4M_COUNT);
    ngb_dev_h2d(b->b4_ptab, ptab, pb);
    {   /* what stays in L2 between the launches of a step: the parameter rows (default), or -- NGB_L2_PERSIST=2, experiment --
         * the stamp rows the loads write and the assembly reads back */
        const char *e = getenv("NGB_L2_PERSIST");
        if (e && atoi(e) == 2) ngb_dev_l2_persist(b->stamp, sizeof(double) * (size_t)b->c->nstamp_rows * b->S);
        else ngb_dev_l2_persist(b->b4_rows_block, mb + pb);
    }
    b->b4_key = b4_batch_key(b->c, mtab, nrows, NULL);
    {   /* overlay reading of the rows (bsim4_eval.cuh, B4OVL) unless NGB_B4_OVERLAY=0 */
        const char *e = getenv("NGB_B4_OVERLAY");
        b->b4_overlay = (e && !atoi(e)) ? 0 : 1;
        if (b->b4_overlay) {
            b4_find_vary(b, prow_t, mtab, ptab);
            if (b->b4_key != NGB_B4_GENERIC) b->b4_key |= B4K_PACK(rowsO, 1);
        }
    }
    return NGB_OK;
}

/* the variant key of a batch (bsim4_variants.h): the selectors of every model row in use and of every instance must agree */
static unsigned b4_batch_key(const ngb_circuit *c, const double *mtab, int nrows, const int *prow_inst)
{
    unsigned key = NGB_B4_GENERIC;
    int i, first = 1;
    if (!c->b4_n) return key;
    for (i = 0; i < (prow_inst ? c->b4_n : nrows); i++) {
        const double *M = mtab + (size_t)(prow_inst ? prow_inst[i] : i) * B4M_COUNT;
        const int fl = c->b4_flags[prow_inst ? i : 0];
        const int v[13] = { (int)M[B4M_mobMod], (int)M[B4M_capMod], (int)M[B4M_cvchargeMod], (int)M[B4M_dioMod], (int)M[B4M_rdsMod],
                            (int)M[B4M_igcMod], (int)M[B4M_igbMod], (int)M[B4M_gidlMod], (int)M[B4M_tempMod], (int)M[B4M_mtrlMod],
                            (int)M[B4M_mtrlCompatMod], B4F_RBODY(fl), B4F_RGATE(fl) };
        unsigned k;
        if (!(B4K_FITS(mobMod, v[0]) && B4K_FITS(capMod, v[1]) && B4K_FITS(cvchargeMod, v[2]) && B4K_FITS(dioMod, v[3]) &&
              B4K_FITS(rdsMod, v[4]) && B4K_FITS(igcMod, v[5]) && B4K_FITS(igbMod, v[6]) && B4K_FITS(gidlMod, v[7]) &&
              B4K_FITS(tempMod, v[8]) && B4K_FITS(mtrlMod, v[9]) && B4K_FITS(mtrlCompatMod, v[10]))) return NGB_B4_GENERIC;
        k = NGB_B4_KEY(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11], v[12], c->opt.method == NGB_GEAR);
        if (first) { key = k; first = 0; }
        else if (k != key) return NGB_B4_GENERIC;
    }
    if (!prow_inst)          /* per-thread rows: the flags are still per instance */
        for (i = 1; i < c->b4_n; i++)
            if (B4F_RBODY(c->b4_flags[i]) != B4F_RBODY(c->b4_flags[0]) || B4F_RGATE(c->b4_flags[i]) != B4F_RGATE(c->b4_flags[0])) return NGB_B4_GENERIC;
    return key;
}
/* which BSIM4 load kernel the batch runs: key[0] = the batch's variant key (0xffffffff: instances differ), key[1] = 1 when
 * the library carries a specialised instantiation for it and it is in use */
int ngbBatchBsim4Variant(ngb_batch *b, unsigned key[2])
{
    key[0] = b->b4_key;
    key[1] = (b->b4_key != NGB_B4_GENERIC && !b->b4_force_generic && b4_variant_built(b->b4_key)) ? 1u : 0u;
    return NGB_OK;
}
/* per-sample rows read as an overlay: how many model / bin columns differ between the samples of an instance (those are read
 * at the thread's own row, all others at the row of sample 0); returns 0 and leaves the counts at -1 when the batch has no
 * per-sample rows or the overlay is switched off (NGB_B4_OVERLAY=0) */
int ngbBatchBsim4Overlay(ngb_batch *b, int *model_columns, int *bin_columns)
{
    int f, nm = 0, np = 0;
    *model_columns = *bin_columns = -1;
    if (!b->b4_prow_t || !b->b4_overlay) return 0;
    for (f = 0; f < B4M_COUNT; f++) nm += b->b4_mvary[f] != 0;
    for (f = 0; f < B4P_COUNT; f++) np += b->b4_pvary[f] != 0;
    *model_columns = nm; *bin_columns = np;
    return 1;
}
void ngbBatchSetBsim4Generic(ngb_batch *b, int on) { b->b4_force_generic = on ? 1 : 0; }

/* ------------------------------------------------------------------ hot path launches */
void ngb_fill_b4ctx(ngb_batch *b, B4Ctx *x)
{
    const ngb_circuit *c = b->c;
    memset(x, 0, sizeof *x);
    x->ninst = c->b4_n; x->S = b->S; x->T = c->b4_n * b->S;
    x->mtab = b->b4_mtab; x->ptab = b->b4_ptab;
    x->prow = b->b4_prow_t ? b->b4_prow_t : b->b4_prow; x->prow_per_thread = b->b4_prow_t ? 1 : 0;
    x->variant = (!b->b4_force_generic && b4_variant_built(b->b4_key)) ? b->b4_key : NGB_B4_GENERIC;
    if (b->b4_prow_t && b->b4_overlay) {
        x->overlay = 1;
        memcpy(x->mvary, b->b4_mvary, sizeof x->mvary); memcpy(x->pvary, b->b4_pvary, sizeof x->pvary);
    }
    x->inst = b->b4_inst; x->flags = b->b4_flags; x->nodes = b->b4_nodes; x->spos = b->b4_spos;
    x->stamp = b->stamp; x->state = b->b4_state; x->op = b->b4_op; x->op_full = b->op_full;
    x->x = b->x; x->neq1 = b->neq1; x->ctl = b->ctl; x->temp = c->opt.temp; x->vt0 = c->opt.vt0;
    x->split = c->exact_order; x->lte_deferred = b->lte_deferred; x->nodeconv = b->nodeconv;
    x->srow0 = c->b4_row0;
}
void ngb_fill_capctx(ngb_batch *b, NgbCapCtx *x)
{
    const ngb_circuit *c = b->c;
    memset(x, 0, sizeof *x);
    x->ninst = c->cap_n; x->S = b->S; x->T = c->cap_n * b->S; x->nodes = b->cap_nodes; x->par = b->cap_par;
    x->spos = b->cap_spos; x->state = b->cap_state; x->stamp = b->stamp; x->x = b->x; x->neq1 = b->neq1; x->ctl = b->ctl;
}
void ngb_fill_b3ctx(ngb_batch *b, B3Ctx *x)
{
    const ngb_circuit *c = b->c;
    memset(x, 0, sizeof *x);
    x->ninst = c->b3_n; x->S = b->S; x->T = c->b3_n * b->S; x->mtab = b->b3_mtab; x->ptab = b->b3_ptab; x->prow = b->b3_prow;
    x->inst = b->b3_inst; x->flags = b->b3_flags; x->nodes = b->b3_nodes; x->spos = b->b3_spos; x->stamp = b->stamp;
    x->state = b->b3_state; x->von = b->b3_von; x->x = b->x; x->neq1 = b->neq1; x->ctl = b->ctl;
    x->temp = c->opt.temp; x->vt0 = c->opt.vt0;
}
void ngb_fill_vbctx(ngb_batch *b, NgbVbicCtx *x)
{
    const ngb_circuit *c = b->c;
    memset(x, 0, sizeof *x);
    x->ninst = c->vb_n; x->S = b->S; x->T = c->vb_n * b->S; x->nstamps = VBIC_NSTAMPS; x->nodes = b->vb_nodes; x->flags = b->vb_flags;
    x->par = b->vb_par; x->aux = b->vb_aux; x->spos = b->vb_spos; x->state = b->vb_state; x->stamp = b->stamp;
    x->x = b->x; x->neq1 = b->neq1; x->ctl = b->ctl;
}
void ngb_fill_dioctx(ngb_batch *b, NgbDioCtx *x)
{
    const ngb_circuit *c = b->c;
    memset(x, 0, sizeof *x);
    x->ninst = c->dio_n; x->S = b->S; x->T = c->dio_n * b->S; x->nodes = b->dio_nodes; x->flags = b->dio_flags; x->par = b->dio_par;
    x->spos = b->dio_spos; x->state = b->dio_state; x->stamp = b->stamp; x->x = b->x; x->neq1 = b->neq1; x->ctl = b->ctl;
    x->reltol = c->opt.reltol; x->abstol = c->opt.abstol; x->vntol = c->opt.vntol; x->chgtol = c->opt.chgtol; x->trtol = c->opt.trtol;
}
void ngb_fill_srcctx(ngb_batch *b, NgbSrcCtx *x, int is_current)
{
    const ngb_circuit *c = b->c;
    memset(x, 0, sizeof *x);
    x->S = b->S; x->is_current = is_current; x->stamp = b->stamp; x->tstep = c->opt.tstep; x->tstop = c->opt.tstop; x->ctl = b->ctl;
    if (is_current) { x->ninst = c->is_n; x->fn = b->is_fn; x->par = b->is_par; x->spos = b->is_spos;
                      x->pwl_ptr = b->is_pwl_ptr; x->pwl = b->is_pwl; }
    else { x->ninst = c->vs_n; x->fn = b->vs_fn; x->par = b->vs_par; x->spos = b->vs_spos;
           x->pwl_ptr = b->vs_pwl_ptr; x->pwl_rep = b->vs_pwl_rep; x->pwl_rdelay = b->vs_pwl_rdelay; x->pwl = b->vs_pwl; }
    x->T = x->ninst * b->S;
}
void ngb_fill_asmctx(ngb_batch *b, NgbAsmCtx *x)
{
    const ngb_circuit *c = b->c;
    memset(x, 0, sizeof *x);
    x->S = b->S; x->nnz = c->nnz; x->neq1 = b->neq1; x->tgt_ptr = b->d_tgt_ptr; x->tgt_rows = b->d_tgt_rows;
    x->slot_diag = b->d_slot_diag; x->long_tgt = b->d_long_tgt; x->nlong = c->nlong; x->long_len_host = b->long_len; x->long_part = b->long_part; x->long_cap = b->long_cap; x->stamp = b->stamp; x->Ax = b->Ax; x->x = b->x; x->add_diag_gmin = 1; x->ctl = b->ctl;
    x->nov = c->ov_n; x->ov_eq = b->ov_eq; x->ov_kind = b->ov_kind; x->ov_cur = b->ov_cur; x->ov_diag = b->ov_diag;
    x->ov_zptr = b->ov_zptr; x->ov_zslot = b->ov_zslot; x->ov_val = b->ov_val;
}
void ngb_fill_luctx(ngb_batch *b, NgbLuCtx *x, int do_factor, int do_solve, int which)
{
    const ngb_circuit *c = b->c;
    memset(x, 0, sizeof *x);
    x->sch = b->dlu[which].dsch; x->pk = b->dlu[which].dpk; x->which = which; x->S = b->S; x->neq1 = b->neq1; x->Ax = b->Ax; x->V = b->V; x->Rs = b->Rs; x->x = b->x;
    x->do_factor = do_factor; x->do_solve = do_solve; x->node_type = b->d_node_type;
    x->reltol = c->opt.reltol; x->abstol = c->opt.abstol; x->vntol = c->opt.vntol;
    x->nodeconv = b->nodeconv; x->singular_col = b->singular; x->ctl = b->ctl;
    x->gV = b->V; x->gRs = b->Rs; x->gZ = b->Zw;
    x->verify = b->lu_verify; x->pivtol = c->pivtol > 0 ? c->pivtol : 0.001;
}

static int check_errflag(ngb_batch *b)
{
    int e[4] = { 0, 0, 0, 0 };
    ngb_dev_d2h(e, b->errflag, sizeof e);
    if (e[0]) { ngb_set_error("device load reported error %d", e[0]); return e[0]; }
    return NGB_OK;
}

int ngb_enqueue_load(ngb_batch *b)
{
    const ngb_circuit *c = b->c;
    int r;
    /* one branch per device type (they write disjoint stamp rows and states); the reference's order of the
     * types only matters for the summation order, which the assembly fixes */
    ngb_dev_branch_begin();
    r = 0;
    /* the small, latency-bound kernels first: their few CTAs are resident when the BSIM4 load fills the rest */
    if (!r && c->vb_n) { NgbVbicCtx x; ngb_dev_branch(1); ngb_fill_vbctx(b, &x); r = ngb_launch_vbic_load(&x, b->errflag); }
    if (!r && c->b3_n) { B3Ctx x; ngb_dev_branch(0); ngb_fill_b3ctx(b, &x); r = ngb_launch_bsim3_load(&x, b->errflag); }
    if (!r && c->dio_n) { NgbDioCtx x; ngb_dev_branch(2); ngb_fill_dioctx(b, &x); r = ngb_launch_dio_load(&x, b->errflag); }
    if (!r && c->cap_n) { NgbCapCtx x; ngb_dev_branch(3); ngb_fill_capctx(b, &x); r = ngb_launch_cap_load(&x, b->errflag); }
    if (!r && c->is_n) { NgbSrcCtx x; ngb_dev_branch(3); ngb_fill_srcctx(b, &x, 1); r = ngb_launch_src_load(&x); }
    if (!r && c->vs_n) { NgbSrcCtx x; ngb_dev_branch(3); ngb_fill_srcctx(b, &x, 0); r = ngb_launch_src_load(&x); }
    ngb_dev_stage_mark(1);
    if (!r && c->b4_n) { B4Ctx x; ngb_dev_branch(-1); ngb_fill_b4ctx(b, &x); r = ngb_launch_bsim4_load(&x, b->errflag); }
    {
        const int rj = ngb_dev_branch_end();
        if (r || rj) return r ? r : rj;
    }
    ngb_dev_stage_mark(2);
    { NgbAsmCtx x; ngb_fill_asmctx(b, &x); if ((r = ngb_launch_assemble(&x))) return r; }
    ngb_dev_stage_mark(3);
    return NGB_OK;
}

/* CKTload alone.  The truncation-error bounds (DEVtrunc -> CKTterr) are NOT evaluated here unless asked for with
 * ngbBatchSetLoadLte: at this level the caller's own CKTtrunc runs DEVtrunc on the host state vectors (the shim), and inside
 * the load they cost five CKTterr bodies per instance plus an atomic minimum per charge on ctl.lte[sample] -- with one
 * sample and a million instances (config 4) a million threads on one address, 3x the whole load (measured) */
int ngbLoad(ngb_batch *b)
{
    int r;
    double *l1 = b->ctl.lte, *l2 = b->ctl.lte2;
    ngb_dev_memset(b->errflag, 0, sizeof(int) * 4);
    ngb_launch_clear_i32(b->ctl.noncon, 0, b->S);          /* CKTnoncon = 0 (niiter.c:71) */
    if (!b->load_lte) b->ctl.lte = b->ctl.lte2 = NULL;
    r = ngb_enqueue_load(b);
    b->ctl.lte = l1; b->ctl.lte2 = l2;
    if (r) return r;
    if ((r = ngb_dev_sync())) return r;
    return check_errflag(b);
}
void ngbBatchSetLoadLte(ngb_batch *b, int on) { b->load_lte = on ? 1 : 0; }
static int lu_call(ngb_batch *b, int f, int s)
{
    NgbLuCtx x; int r;
    if (!b->have_lu) { ngb_set_error("no LU pattern: call ngbCircuitSetLuPattern or ngbCircuitAnalyze before ngbBatchCreate"); return NGB_E_PANIC; }
    ngb_fill_luctx(b, &x, f, s, b->lu_which);
    x.ctl.lusel = NULL;                                   /* direct calls: every active sample, on the selected set */
    if (s) ngb_launch_clear_i32(b->nodeconv, 0, b->S);
    if ((r = ngb_launch_lu(&x))) return r;
    return ngb_dev_sync();
}
int ngbLuFac(ngb_batch *b) { return lu_call(b, 1, 0); }
int ngbSolve(ngb_batch *b) { return lu_call(b, 0, 1); }
int ngbLuFacSolve(ngb_batch *b) { return lu_call(b, 1, 1); }
int ngbNewtonStep(ngb_batch *b) { int r = ngbLoad(b); return r ? r : ngbLuFacSolve(b); }
