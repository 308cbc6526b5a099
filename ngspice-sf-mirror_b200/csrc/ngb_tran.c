/* ngb_tran.c -- host driver of the device-resident transient analysis.
 *
 * The host only enqueues "ticks" (one Newton step for every running sample: device loads,
 * assembly + LU + solve + node convergence, controller) and polls a done counter; every
 * decision of NIiter / DCtran is taken on the device per sample (ngb_tran.cuh).
 * Stands in for DCtran (src/spicelib/analysis/dctran.c:66) called through CKTdoJob, for a
 * whole batch of circuits at once.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include "ngb_dev.h"
#include "ngb_host.h"
#include "../../include/ngb200.h"

struct ngb_tran {
    NgbTranCtx x;              /* device pointers */
    int max_points, nsave;
    int *d_save_eq;
    int *d_mask;               /* [S] samples of a repivot_suspended pass */
    int *d_ms_eq, *d_ms_kind, *d_ms_count; double *d_ms_val, *d_ms_td;
    long ticks; int repivots;
    int keep_set[NGB_LU_SETS]; /* pattern sets a re-pivoted sample was moved to: launched every step from then on */
    int stage;                 /* 0: some sample is still in the operating point; 1: all in the transient;
                                * 2: all past the last pivoting event (only the final pattern set is in use) */
    /* CUDA graph of one Newton step, one per launch sequence (= per stage);
     * graph_state 0 = not tried, 1 = ready, -1 = unavailable */
    void *graph[3]; int graph_nodes[3], graph_state[3];
};

static void *dz(size_t bytes) { return ngb_dev_malloc(bytes ? bytes : 8); }

void ngb_tran_free(struct ngb_batch *b)
{
    struct ngb_tran *t = b->tran;
    if (!t) return;
    ngb_dev_free(t->x.phase); ngb_dev_free(t->x.iterno); ngb_dev_free(t->x.firsttime); ngb_dev_free(t->x.nbreak);
    ngb_dev_free(t->x.npts); ngb_dev_free(t->x.brkflag); ngb_dev_free(t->x.accepted); ngb_dev_free(t->x.rejected);
    ngb_dev_free(t->x.numiter); ngb_dev_free(t->x.timepts); ngb_dev_free(t->x.save_delta); ngb_dev_free(t->x.old_delta);
    ngb_dev_free(t->x.breaks); ngb_dev_free(t->x.out_time); ngb_dev_free(t->x.out_val); ngb_dev_free(t->x.ndone); ngb_dev_free(t->x.evstage); ngb_dev_free(t->x.verify);
    ngb_dev_free(t->x.gm_stage); ngb_dev_free(t->x.gm_factor); ngb_dev_free(t->x.gm_oldgmin); ngb_dev_free(t->x.gm_xold);
    ngb_dev_free(t->x.gm_startgmin); ngb_dev_free(t->x.gs_conv); ngb_dev_free(t->x.gs_raise); ngb_dev_free(t->x.gs_i);
    { int a; for (a = 0; a < t->x.gm_narr; a++) ngb_dev_free(t->x.gm_arr[a].old); }
    ngb_dev_free(t->d_save_eq); ngb_dev_free(t->x.isrc_break); ngb_dev_free(t->x.vsrc_break);
    ngb_dev_free(t->x.susp); ngb_dev_free(t->d_mask); ngb_dev_free(t->x.ipass);
    ngb_dev_free(t->d_ms_eq); ngb_dev_free(t->d_ms_kind); ngb_dev_free(t->d_ms_count); ngb_dev_free(t->d_ms_val); ngb_dev_free(t->d_ms_td);
    ngb_dev_free(t->x.ms_i); ngb_dev_free(t->x.ms_d);
    ngb_dev_graph_destroy(t->graph[0]); ngb_dev_graph_destroy(t->graph[1]); ngb_dev_graph_destroy(t->graph[2]);
    free(t);
    b->tran = NULL;
}

static int tran_setup(ngb_batch *b, int max_points, const int *save_eq, int nsave)
{
    const ngb_circuit *c = b->c;
    const int S = b->S;
    struct ngb_tran *t;
    NgbTranCtx *x;
    int i, s;
    ngb_tran_free(b);
    t = (struct ngb_tran *)calloc(1, sizeof *t);
    b->tran = t;
    b->lte_deferred = getenv("NGB_LTE_INLOAD") ? 0 : 1;       /* BSIM4trunc after the solve, converged samples only */
    x = &t->x;
    x->ctl = b->ctl; x->S = S; x->neq1 = b->neq1; x->x = b->x;
    x->nodeconv = b->nodeconv; x->nodeconv_w = b->nodeconv;
    x->phase = (int *)dz(sizeof(int) * S); x->iterno = (int *)dz(sizeof(int) * S);
    x->firsttime = (int *)dz(sizeof(int) * S); x->nbreak = (int *)dz(sizeof(int) * S);
    x->npts = (int *)dz(sizeof(int) * S); x->brkflag = (int *)dz(sizeof(int) * S);
    x->accepted = (int *)dz(sizeof(int) * S); x->rejected = (int *)dz(sizeof(int) * S);
    x->numiter = (int *)dz(sizeof(int) * S); x->timepts = (int *)dz(sizeof(int) * S);
    x->save_delta = (double *)dz(sizeof(double) * S); x->old_delta = (double *)dz(sizeof(double) * S);
    x->maxbrk = c->vs_n + c->is_n + 4;                      /* one pending corner per source at most, both ends, one spare */
    if (x->maxbrk < NGB_MAXBRK) x->maxbrk = NGB_MAXBRK;
    x->breaks = (double *)dz(sizeof(double) * (size_t)x->maxbrk * S);
    x->out_time = (double *)dz(sizeof(double) * (size_t)S * max_points);
    x->out_val = (double *)dz(sizeof(double) * (size_t)S * max_points * (nsave ? nsave : 1));
    x->ndone = (int *)dz(sizeof(int) * 4); x->evstage = (int *)dz(sizeof(int) * S);
    x->susp = (int *)dz(sizeof(int) * S); t->d_mask = (int *)dz(sizeof(int) * S);
    x->verify = (int *)dz(sizeof(int) * S); b->lu_verify = x->verify;
    x->ipass = (int *)dz(sizeof(int) * S);
    if (b->ms_n > 0) {
        const int nm = b->ms_n;
        double *nanv = (double *)malloc(sizeof(double) * 3 * (size_t)nm * S);
        int m2; size_t q;
        t->d_ms_eq = (int *)dz(sizeof(int) * nm); t->d_ms_kind = (int *)dz(sizeof(int) * nm); t->d_ms_count = (int *)dz(sizeof(int) * nm);
        t->d_ms_val = (double *)dz(sizeof(double) * nm); t->d_ms_td = (double *)dz(sizeof(double) * nm);
        x->ms_i = (int *)dz(sizeof(int) * 4 * (size_t)nm * S); x->ms_d = (double *)dz(sizeof(double) * 3 * (size_t)nm * S);
        if (!nanv || !x->ms_i || !x->ms_d) { free(nanv); ngb_set_error("measurement buffers: out of memory"); return NGB_E_PANIC; }
        ngb_dev_h2d(t->d_ms_eq, b->ms_eq, sizeof(int) * nm); ngb_dev_h2d(t->d_ms_kind, b->ms_kind, sizeof(int) * nm);
        ngb_dev_h2d(t->d_ms_count, b->ms_count, sizeof(int) * nm);
        ngb_dev_h2d(t->d_ms_val, b->ms_val, sizeof(double) * nm); ngb_dev_h2d(t->d_ms_td, b->ms_td, sizeof(double) * nm);
        for (m2 = 0; m2 < nm; m2++)
            for (q = 0; q < 3 * (size_t)S; q++) nanv[(size_t)m2 * 3 * S + q] = (q >= 2 * (size_t)S) ? NAN : 0.0;   /* m_measured = NAN until found */
        ngb_dev_h2d(x->ms_d, nanv, sizeof(double) * 3 * (size_t)nm * S);
        free(nanv);
        x->nmeas = nm; x->ms_eq = t->d_ms_eq; x->ms_kind = t->d_ms_kind; x->ms_count = t->d_ms_count; x->ms_val = t->d_ms_val; x->ms_td = t->d_ms_td;
    }
    { int i2; x->had_nodeset = 0; for (i2 = 0; i2 < c->ov_n; i2++) if (c->ov_kind[i2] == 0) x->had_nodeset = 1; }
    ngb_fill_srcctx(b, &x->isrc, 1); ngb_fill_srcctx(b, &x->vsrc, 0);
    x->isrc_break = (double *)dz(sizeof(double) * (size_t)(x->isrc.ninst > 0 ? x->isrc.ninst : 1) * S);
    x->vsrc_break = (double *)dz(sizeof(double) * (size_t)(x->vsrc.ninst > 0 ? x->vsrc.ninst : 1) * S);
    ngb_launch_fill_f64(x->isrc_break, -1.0, (x->isrc.ninst > 0 ? x->isrc.ninst : 1) * S);
    ngb_launch_fill_f64(x->vsrc_break, -1.0, (x->vsrc.ninst > 0 ? x->vsrc.ninst : 1) * S);
    t->d_save_eq = (int *)dz(sizeof(int) * (nsave ? nsave : 1));
    if (!x->out_val || !x->out_time || !x->breaks) { ngb_set_error("transient buffers: out of device memory"); return NGB_E_PANIC; }
    if (nsave) ngb_dev_h2d(t->d_save_eq, save_eq, sizeof(int) * (size_t)nsave);
    x->save_eq = t->d_save_eq; x->max_points = max_points; x->nsave = nsave;
    t->max_points = max_points; t->nsave = nsave;
    x->tstep = c->opt.tstep; x->tstop = c->opt.tstop; x->tmax = c->opt.tmax; x->tstart = c->opt.tstart;
    x->delmin = c->opt.delmin; x->minbreak = c->opt.minbreak; x->xmu = c->opt.xmu;
    if (!c->opt.uic) {
        /* dynamic gmin stepping needs a copy of the solution and of every device's CKTstate0 per sample */
        struct { double *st; int K, n; } tab[5] = {
            { b->b4_state, B4ST_COUNT, c->b4_n }, { b->b3_state, B3ST_COUNT, c->b3_n }, { b->vb_state, VBS_COUNT, c->vb_n },
            { b->dio_state, DIOST_COUNT, c->dio_n }, { b->cap_state, 2, c->cap_n } };
        x->gm_stage = (int *)dz(sizeof(int) * S);
        x->gm_factor = (double *)dz(sizeof(double) * S); x->gm_oldgmin = (double *)dz(sizeof(double) * S);
        x->gm_xold = (double *)dz(sizeof(double) * (size_t)b->neq1 * S);
        x->gm_startgmin = (double *)dz(sizeof(double) * S); x->gs_conv = (double *)dz(sizeof(double) * S);
        x->gs_raise = (double *)dz(sizeof(double) * S); x->gs_i = (int *)dz(sizeof(int) * S);
        x->num_gmin_steps = c->opt.num_gmin_steps; x->num_src_steps = c->opt.num_src_steps;
        x->itl2 = c->opt.itl2; x->gmin_factor = c->opt.gmin_factor; x->gshunt = c->opt.gshunt;
        for (i = 0; i < 5; i++)
            if (tab[i].st && tab[i].n > 0) {
                x->gm_arr[x->gm_narr].state = tab[i].st; x->gm_arr[x->gm_narr].K = tab[i].K; x->gm_arr[x->gm_narr].ninst = tab[i].n;
                x->gm_arr[x->gm_narr].old = (double *)dz(sizeof(double) * (size_t)tab[i].K * tab[i].n * S);
                if (!x->gm_arr[x->gm_narr].old) { ngb_set_error("gmin-stepping buffers: out of device memory"); return NGB_E_PANIC; }
                x->gm_narr++;
            }
        x->gm_enable = 1;
    }
    ngb_lu_events(c, x->lu_event);
    x->nluset = (x->lu_event[0] != x->lu_event[1] || x->lu_event[1] != x->lu_event[2] || x->lu_event[2] != x->lu_event[3]) ? 2 : 1;
    x->pivot_events = (c->pivot_mode == 1 || (c->pivot_mode < 0 && c->klu_P && !getenv("NGB_BATCH_PIVOT"))) ? 1 : 0;
    if (x->pivot_events) x->nluset = 2;              /* every sample selects its own pattern set */
    /* pivoting events: the device checks the recorded order of the event on the sample's own matrix (NgbLuSched.vchk) and only
     * the samples for which KLU's rule would choose differently go to the host (NGB_HOST_PIVOT=1: all of them, as before) */
    x->dev_verify = 0;
    if (x->pivot_events && !getenv("NGB_HOST_PIVOT")) {
        int e, ok = 1;
        for (e = 0; e < NGB_LU_EVENTS; e++) if (!c->lu[x->lu_event[e]].valid || !c->lu[x->lu_event[e]].sch.vchk) ok = 0;
        x->dev_verify = ok;
    }
    memset(t->keep_set, 0, sizeof t->keep_set);
    x->maxorder = c->opt.maxorder; x->uic = c->opt.uic; x->max_iter_tran = c->opt.itl4; x->max_iter_dc = c->opt.itl1;
    if (x->minbreak == 0) x->minbreak = x->tmax * 5e-5;            /* dctran.c:163-164 */

    /* initial per-sample state: DCtran entry (dctran.c:117-236) */
    {
        int *iv = (int *)calloc((size_t)S, sizeof(int));
        double *dv = (double *)calloc((size_t)S * x->maxbrk, sizeof(double));
        const int mode0 = (c->opt.uic ? NGB_MODEUIC : 0) | NGB_MODETRANOP | NGB_MODEINITJCT;
        for (s = 0; s < S; s++) iv[s] = mode0;
        ngb_dev_h2d(b->ctl.mode, iv, sizeof(int) * S);
        for (s = 0; s < S; s++) iv[s] = c->opt.uic ? NGB_PH_OPUIC : NGB_PH_DCOP;
        ngb_dev_h2d(x->phase, iv, sizeof(int) * S);
        for (s = 0; s < S; s++) iv[s] = 1;
        ngb_dev_h2d(x->firsttime, iv, sizeof(int) * S);
        ngb_dev_h2d(x->brkflag, iv, sizeof(int) * S);              /* CKTbreak = 1 (dctran.c:195): ISRCaccept's PWL reads it */
        ngb_dev_h2d(b->ctl.active, iv, sizeof(int) * S);
        ngb_dev_h2d(b->ctl.order, iv, sizeof(int) * S);
        for (s = 0; s < S; s++) iv[s] = 2;
        ngb_dev_h2d(x->nbreak, iv, sizeof(int) * S);
        /* NIiter under MODETRANOP|MODEUIC swaps rhs/rhsOld before its single CKTload */
        for (s = 0; s < S; s++) iv[s] = c->opt.uic ? 1 : 0;
        ngb_dev_h2d(b->ctl.xsel, iv, sizeof(int) * S);
        for (s = 0; s < S; s++) iv[s] = (x->pivot_events && !x->dev_verify) ? -1 : x->lu_event[c->opt.uic ? 2 : 0];   /* the first factor of a run pivots */
        ngb_dev_h2d(b->ctl.lusel, iv, sizeof(int) * S);
        for (s = 0; s < S; s++) iv[s] = x->dev_verify ? 1 : 0;
        ngb_dev_h2d(x->verify, iv, sizeof(int) * S);
        memset(iv, 0, sizeof(int) * S);
        ngb_dev_h2d(b->ctl.head, iv, sizeof(int) * S);
        ngb_dev_h2d(b->ctl.noncon, iv, sizeof(int) * S);
        ngb_dev_h2d(b->ctl.err, iv, sizeof(int) * S);
        ngb_dev_h2d(b->ctl.stateop, iv, sizeof(int) * S);
        ngb_dev_h2d(b->nodeconv, iv, sizeof(int) * S);
        for (i = 0; i < x->maxbrk; i++) for (s = 0; s < S; s++) dv[(size_t)i * S + s] = (i == 0) ? 0.0 : c->opt.tstop;
        ngb_dev_h2d(x->breaks, dv, sizeof(double) * (size_t)x->maxbrk * S);
        memset(dv, 0, sizeof(double) * S);
        ngb_dev_h2d(b->ctl.time, dv, sizeof(double) * S);
        ngb_dev_h2d(b->ctl.delta, dv, sizeof(double) * S);
        ngb_dev_h2d(b->ctl.ag0, dv, sizeof(double) * S);
        ngb_dev_h2d(b->ctl.ag1, dv, sizeof(double) * S);
        ngb_dev_h2d(b->ctl.diag_gmin, dv, sizeof(double) * S);
        free(iv); free(dv);
    }
    if (!c->opt.uic && c->opt.no_op_iter && (c->opt.num_gmin_steps == 1 || c->opt.num_src_steps == 1)) {
        /* CKTnoOpIter (cktop.c:42-55): no plain NIiter -- every sample starts inside the first fallback, in the
         * state the controller would have put it in (solution and states are still zero) */
        int *iv = (int *)calloc((size_t)S, sizeof(int)); double *dv = (double *)calloc((size_t)S, sizeof(double));
        const int dyn = c->opt.num_gmin_steps == 1;
        for (s = 0; s < S; s++) iv[s] = dyn ? 1 : 10;
        ngb_dev_h2d(x->gm_stage, iv, sizeof(int) * S);
        if (dyn) {
            for (s = 0; s < S; s++) dv[s] = c->opt.gmin_factor;
            ngb_dev_h2d(x->gm_factor, dv, sizeof(double) * S);
            for (s = 0; s < S; s++) dv[s] = 1e-2;
            ngb_dev_h2d(x->gm_oldgmin, dv, sizeof(double) * S);
            for (s = 0; s < S; s++) dv[s] = 1e-2 / c->opt.gmin_factor;
            ngb_dev_h2d(b->ctl.diag_gmin, dv, sizeof(double) * S);
        } else {
            memset(dv, 0, sizeof(double) * S);
            ngb_dev_h2d(b->ctl.srcfact, dv, sizeof(double) * S);         /* gs_conv is already zero */
        }
        free(iv); free(dv);
    }
    ngb_launch_fill_f64(b->ctl.lte, 1e300, S);
    ngb_launch_fill_f64(b->ctl.lte2, 1e300, S);
    ngb_dev_memset(b->x, 0, sizeof(double) * 2 * (size_t)b->neq1 * S);
    if (b->b4_state) ngb_dev_memset(b->b4_state, 0, sizeof(double) * NGB_NHIST * B4ST_COUNT * (size_t)c->b4_n * S);
    if (b->b3_state) { ngb_dev_memset(b->b3_state, 0, sizeof(double) * NGB_NHIST * B3ST_COUNT * (size_t)c->b3_n * S);
                       ngb_dev_memset(b->b3_von, 0, sizeof(double) * (size_t)c->b3_n * S); }
    if (b->vb_state) ngb_dev_memset(b->vb_state, 0, sizeof(double) * NGB_NHIST * VBS_COUNT * (size_t)c->vb_n * S);
    if (b->dio_state) ngb_dev_memset(b->dio_state, 0, sizeof(double) * NGB_NHIST * DIOST_COUNT * (size_t)c->dio_n * S);
    if (b->cap_state) ngb_dev_memset(b->cap_state, 0, sizeof(double) * NGB_NHIST * 2 * (size_t)c->cap_n * S);
    if (b->b4_op) ngb_dev_memset(b->b4_op, 0, sizeof(double) * B4O_COUNT * (size_t)c->b4_n * S);
    return ngb_dev_sync();
}

static int enqueue_tick_direct(ngb_batch *b, int with_lu)
{
    int r;
    ngb_dev_stage_begin();
    if ((r = ngb_enqueue_load(b))) return r;
    if (with_lu) {
        const int *ev = b->tran->x.lu_event;
        const int first = b->tran->stage == 0 ? 0 : (b->tran->stage == 1 ? 2 : 3);      /* events still ahead of some sample */
        int w, e;
        for (w = 0; w < NGB_LU_SETS; w++) {
            NgbLuCtx lx;
            int used = 0;
            if (!b->dlu[w].valid) continue;
            if (!b->tran->x.pivot_events || b->tran->x.dev_verify) for (e = first; e < NGB_LU_EVENTS; e++) if (ev[e] == w) used = 1;
            if (b->tran->keep_set[w]) used = 1;
            if (!used) continue;                                            /* no sample can be on this set any more */
            ngb_fill_luctx(b, &lx, 1, 1, w);
            if (b->tran->x.nluset == 1) lx.ctl.lusel = NULL;               /* one set: no per-sample selection */
            lx.V = NULL;                       /* fused factor+solve: the factors never leave shared memory */
            if ((r = ngb_launch_lu(&lx))) return r;
        }
        ngb_dev_stage_mark(4);
    }
    if (with_lu && b->lte_deferred && b->c->b4_n) {     /* BSIM4trunc for the samples whose iteration can have converged */
        B4Ctx x;
        ngb_fill_b4ctx(b, &x);
        if ((r = ngb_launch_bsim4_lte(&x))) return r;
        ngb_dev_stage_mark(5);
    }
    r = ngb_launch_tran_control(&b->tran->x);
    ngb_dev_stage_mark(6);
    return r;
}


/* Samples whose refactor met a zero pivot wait, inactive, with susp == 1 (ngb_tran.cuh).  The reference answers a
 * zero pivot by factoring the same matrix again with pivoting inside the same iteration (niiter.c:162-195, KLU mode:
 * SMPreorder without reloading, KLUloadDiagGmin = 0); here the host does that per sample with the library's own pivoting
 * factor (ngb_pivot.c), moves the sample to the resulting pattern set and repeats the rest of its Newton step -- assembly
 * (the failed attempt solved in place over the right-hand side), refactor + solve, BSIM4trunc, controller -- for these
 * samples alone.  A sample whose matrix the pivoting factor finds singular gets susp == 2: its NIiter returns
 * E_SINGULAR, which CKTop answers with its fallbacks and DCtran with a shorter step. */
static int repivot_suspended(ngb_batch *b)
{
    struct ngb_tran *t = b->tran;
    const ngb_circuit *c = b->c;
    const int S = b->S, nnz = c->nnz;
    int *susp = (int *)calloc((size_t)S, sizeof(int)), *mask = (int *)calloc((size_t)S, sizeof(int)), *list = (int *)calloc((size_t)S, sizeof(int));
    int s, k, m = 0, r = NGB_OK, w, used[NGB_LU_SETS], g, zero4[1] = { 0 };
    double *Ax = NULL;
    NgbRepivot *res = NULL;
    if (!susp || !mask || !list) { r = NGB_E_PANIC; goto out; }
    memset(used, 0, sizeof used);
    ngb_dev_sync();
    ngb_dev_d2h(susp, t->x.susp, sizeof(int) * (size_t)S);
    for (s = 0; s < S; s++) if (susp[s] == 1) { list[m++] = s; mask[s] = 1; }
    if (!m) goto out;
    /* the matrices: one transfer when most of the batch waits (a pivoting event), sample by sample otherwise */
    Ax = (double *)malloc(sizeof(double) * (size_t)nnz * (size_t)(m > S / 8 ? S : m));
    res = (NgbRepivot *)calloc((size_t)m, sizeof(NgbRepivot));
    if (!Ax || !res) { r = NGB_E_PANIC; goto out; }
    if (m > S / 8) ngb_dev_d2h(Ax, b->Ax, sizeof(double) * (size_t)nnz * S);
    else for (k = 0; k < m; k++) ngb_dev_d2h(Ax + (size_t)k * nnz, b->Ax + (size_t)list[k] * nnz, sizeof(double) * (size_t)nnz);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 16)
#endif
    for (k = 0; k < m; k++)
        ngb_repivot_compute(c, Ax + (size_t)(m > S / 8 ? list[k] : k) * nnz, &res[k]);
    {
        int *lus = (int *)malloc(sizeof(int) * (size_t)S), *st = (int *)calloc((size_t)S, sizeof(int));
        int *act = (int *)malloc(sizeof(int) * (size_t)S), *er = (int *)malloc(sizeof(int) * (size_t)S), *nc = (int *)malloc(sizeof(int) * (size_t)S);
        double *l1 = (double *)malloc(sizeof(double) * (size_t)S), *l2 = (double *)malloc(sizeof(double) * (size_t)S);
        if (!lus || !st || !act || !er || !nc || !l1 || !l2) r = NGB_E_PANIC;
        else {
            ngb_dev_d2h(lus, b->ctl.lusel, sizeof(int) * (size_t)S); ngb_dev_d2h(act, b->ctl.active, sizeof(int) * (size_t)S);
            ngb_dev_d2h(er, b->ctl.err, sizeof(int) * (size_t)S); ngb_dev_d2h(nc, b->nodeconv, sizeof(int) * (size_t)S);
            ngb_dev_d2h(l1, b->ctl.lte, sizeof(double) * (size_t)S); ngb_dev_d2h(l2, b->ctl.lte2, sizeof(double) * (size_t)S);
            memcpy(st, susp, sizeof(int) * (size_t)S);
            for (k = 0; k < m && r == NGB_OK; k++) {
                s = list[k];
                r = ngb_repivot_commit(b, &res[k], &w);
                if (r == NGB_OK) { lus[s] = w; used[w] = 1; t->keep_set[w] = 1; t->repivots++; st[s] = 0; er[s] = 0; }
                else if (r == NGB_E_UNSUPP && t->x.pivot_events && c->lu[t->x.lu_event[3]].valid && res[k].rc == NGB_OK) {
                    /* no free pattern set: this sample keeps the batch's recorded order (round-1 behaviour) */
                    w = t->x.lu_event[3]; lus[s] = w; used[w] = 1; t->keep_set[w] = 1; st[s] = 0; er[s] = 0; r = NGB_OK;
                    if (!b->dlu[w].valid) r = NGB_E_PANIC;
                }
                else if (r == NGB_E_SINGULAR || r == NGB_E_UNSUPP) { st[s] = 2; er[s] = NGB_E_SINGULAR; if (lus[s] < 0) lus[s] = t->x.lu_event[3]; r = NGB_OK; }   /* the controller takes it from here */
                act[s] = 1; nc[s] = 0; l1[s] = 1e300; l2[s] = 1e300;
            }
            if (r == NGB_OK) {
                ngb_dev_h2d(b->ctl.lusel, lus, sizeof(int) * (size_t)S); ngb_dev_h2d(b->ctl.active, act, sizeof(int) * (size_t)S);
                ngb_dev_h2d(b->ctl.err, er, sizeof(int) * (size_t)S); ngb_dev_h2d(b->nodeconv, nc, sizeof(int) * (size_t)S);
                ngb_dev_h2d(b->ctl.lte, l1, sizeof(double) * (size_t)S); ngb_dev_h2d(b->ctl.lte2, l2, sizeof(double) * (size_t)S);
                ngb_dev_h2d(t->x.susp, st, sizeof(int) * (size_t)S);
            }
        }
        free(lus); free(st); free(act); free(er); free(nc); free(l1); free(l2);
    }
    if (r == NGB_OK) {
        ngb_dev_h2d(t->d_mask, mask, sizeof(int) * (size_t)S);
        ngb_dev_h2d(t->x.ndone + 3, zero4, sizeof(int));
        /* per-sample pattern selection is needed from now on, and the captured launch sequences are stale */
        t->x.nluset = 2;
        for (g = 0; g < 3; g++) { ngb_dev_graph_destroy(t->graph[g]); t->graph[g] = NULL; t->graph_state[g] = 0; }
        {
            NgbAsmCtx ax;
            ngb_fill_asmctx(b, &ax);
            ax.ctl.active = t->d_mask;
            r = ngb_launch_assemble(&ax);
        }
        for (w = 0; w < NGB_LU_SETS && r == NGB_OK; w++) {
            NgbLuCtx lx;
            if (!used[w]) continue;
            ngb_fill_luctx(b, &lx, 1, 1, w);
            lx.ctl.active = t->d_mask;
            lx.V = NULL; lx.verify = NULL;           /* these samples' orders come from the pivoting factor itself */
            r = ngb_launch_lu(&lx);
        }
        if (r == NGB_OK && b->lte_deferred && b->c->b4_n) {
            B4Ctx x;
            ngb_fill_b4ctx(b, &x);
            x.ctl.active = t->d_mask;
            r = ngb_launch_bsim4_lte(&x);
        }
        if (r == NGB_OK) {
            NgbTranCtx cx = t->x;
            cx.only = t->d_mask;
            r = ngb_launch_tran_control(&cx);
        }
        if (r == NGB_OK) r = ngb_dev_sync();
    }
out:
    if (res) for (k = 0; k < m; k++) ngb_repivot_free(&res[k]);
    free(res); free(Ax); free(susp); free(mask); free(list);
    return r;
}

/* one Newton step for the whole batch.  The launch sequence is the same every step, so it is captured
 * once into a CUDA graph and replayed (six to eight launches become one); steps sampled by the
 * kernel-timing facility, and the host build, go kernel by kernel. */
static int enqueue_tick(ngb_batch *b, int with_lu)
{
    struct ngb_tran *t = b->tran;
    const int g = t->stage;
    int r;
    if (!with_lu || ngb_dev_profile_due()) return enqueue_tick_direct(b, with_lu);
    if (t->graph_state[g] == 0) {
        if (getenv("NGB_NO_GRAPH") || ngb_dev_graph_begin()) {
            t->graph_state[g] = -1;
        } else {
            r = enqueue_tick_direct(b, with_lu);
            if (ngb_dev_graph_end(&t->graph[g], &t->graph_nodes[g]) || r) { t->graph_state[g] = -1; if (r) return r; }
            else t->graph_state[g] = 1;
        }
    }
    if (t->graph_state[g] == 1) return ngb_dev_graph_launch(t->graph[g], t->graph_nodes[g]);
    return enqueue_tick_direct(b, with_lu);
}

int ngbTranRun(ngb_batch *b, int max_points, const int *save_eq, int nsave)
{
    const int S = b->S;
    int r, done[4] = { 0, 0, 0, 0 }, e[4] = { 0, 0, 0, 0 };
    long tick = 0, max_ticks;
    int check_every = 64;
    if (!b->have_lu) { ngb_set_error("no LU pattern set for this circuit"); return NGB_E_PANIC; }
    if (b->c->opt.tstop <= 0 || b->c->opt.tstep <= 0) { ngb_set_error("transient parameters not set"); return NGB_E_PANIC; }
    if ((r = tran_setup(b, max_points, save_eq, nsave))) return r;
    ngb_dev_memset(b->errflag, 0, sizeof(int) * 4);
    {   /* safety cap on the Newton steps of the run: 64 per stored point, or (nothing stored, measurements only) per print step x 16 */
        long pts = max_points;
        const double steps = b->c->opt.tstop / b->c->opt.tstep;
        if (b->ms_n > 0 && steps * 16 > (double)pts) pts = (long)(steps * 16);
        max_ticks = 64L * pts + 1024;
    }
    if (b->c->opt.uic) {
        /* CKTop under UIC: one CKTload, no factorisation (niiter.c:41-47) */
        if ((r = enqueue_tick(b, 0))) return r;
        tick++;
    }
    while (tick < max_ticks) {
        int i;
        /* the pivoting events of a run sit in its first iterations: look after every step there */
        const int burst = (b->tran->x.pivot_events && tick < 6) ? 1 : check_every;
        for (i = 0; i < burst; i++, tick++)
            if ((r = enqueue_tick(b, 1))) return r;
        ngb_dev_d2h(done, b->tran->x.ndone, sizeof(int) * 4);
        if (done[3] > 0) {                                   /* zero pivots: re-pivot those samples (niiter.c:162-195) */
            if ((r = repivot_suspended(b))) return r;
            ngb_dev_d2h(done, b->tran->x.ndone, sizeof(int) * 4);
        }
        b->tran->stage = (done[2] >= S) ? 2 : ((done[1] >= S) ? 1 : 0);
        ngb_dev_d2h(e, b->errflag, sizeof e);
        if (e[0]) { ngb_set_error("device load reported error %d", e[0]); return e[0]; }
        if (done[0] >= S) break;
    }
    b->tran->ticks = tick;
    b->lte_deferred = 0;                 /* direct ngbLoad calls keep the bound inside the load */
    if (done[0] < S) { ngb_set_error("transient did not finish within %ld Newton steps", max_ticks); return NGB_E_ITERLIM; }
    return NGB_OK;
}

int ngbTranStats(ngb_batch *b, int *accepted, int *rejected, int *numiter, int *npoints)
{
    struct ngb_tran *t = b->tran;
    const size_t n = sizeof(int) * (size_t)b->S;
    if (!t) return NGB_E_PANIC;
    if (accepted) ngb_dev_d2h(accepted, t->x.accepted, n);
    if (rejected) ngb_dev_d2h(rejected, t->x.rejected, n);
    if (numiter) ngb_dev_d2h(numiter, t->x.numiter, n);
    if (npoints) ngb_dev_d2h(npoints, t->x.npts, n);
    return NGB_OK;
}
/* what DCtran returned for each sample: 0, or the reference's error number (E_SINGULAR 102, E_ITERLIM 103 -- the operating
 * point could not be found --, E_TIMESTEP 106, ...).  ngbTranRun itself fails only when the batch as a whole cannot go on */
int ngbTranErrors(ngb_batch *b, int *err)
{
    if (!b->tran || !err) return NGB_E_PANIC;
    ngb_dev_d2h(err, b->ctl.err, sizeof(int) * (size_t)b->S);
    return NGB_OK;
}
long ngbTranWaveBytes(ngb_batch *b)
{
    struct ngb_tran *t = b->tran;
    return t ? (long)(sizeof(double) * (size_t)b->S * t->max_points * (t->nsave ? t->nsave : 1)) : 0;
}
int ngbTranWaves(ngb_batch *b, double *times, double *values)
{
    struct ngb_tran *t = b->tran;
    if (!t) return NGB_E_PANIC;
    if (times) ngb_dev_d2h(times, t->x.out_time, sizeof(double) * (size_t)b->S * t->max_points);
    if (values && t->nsave) ngb_dev_d2h(values, t->x.out_val, sizeof(double) * (size_t)b->S * t->max_points * t->nsave);
    return NGB_OK;
}
/* Batch-aware binary rawfile of the last ngbTranRun: one plot per sample, one after the other in ONE file (the way the
 * reference lays out the plots of several analyses in a rawfile; `load` reads them as tran1, tran2, ...).  Each plot is what
 * `ngspice -b -r` writes for a transient run -- the header of fileInit (src/frontend/outitf.c:881-923: Title / Date /
 * Command / Plotname / Flags / No. Variables / No. Points padded to 8 columns / Variables), the variable lines of
 * fileInit_pass2 (:997-1029: "\t<index>\t<name>\t<type>", index 0 is `time`), `Binary:` and then the points row by row as
 * raw doubles (fileStartPoint / fileAddRealValue / fileEndPoint, :1048-1092).  names / types: the nsave saved equations in
 * the order given to ngbTranRun ("v(out)" / "voltage", "i(vdd)" / "current"); date NULL = now in the reference's datestring
 * format (src/misc/misc_time.c).  The waveforms come from the device in two copies, whatever the number of samples. */
int ngbTranWriteRaw(ngb_batch *b, const char *path, const char *title, const char *date, const char *const *names, const char *const *types,
                    int first_sample, int nsamples)
{
    struct ngb_tran *t = b ? b->tran : NULL;
    const int S = b ? b->S : 0;
    int s, k, r = NGB_OK, *npts = NULL;
    double *tm = NULL, *val = NULL, *row = NULL;
    char datebuf[64];
    FILE *fp;
    if (!t || t->max_points <= 0) { ngb_set_error("no stored waveforms: run ngbTranRun with max_points > 0 first"); return NGB_E_PANIC; }
    if (!path || !names || !types || first_sample < 0 || nsamples < 1 || first_sample + nsamples > S) { ngb_set_error("ngbTranWriteRaw: bad arguments"); return NGB_E_PANIC; }
    if (!date) {
        time_t now = time(NULL);
        struct tm tmv;
        localtime_r(&now, &tmv);
        strftime(datebuf, sizeof datebuf, "%a %b %e %H:%M:%S  %Y", &tmv);          /* asctime without its newline */
        date = datebuf;
    }
    npts = (int *)malloc(sizeof(int) * (size_t)S);
    tm = (double *)malloc(sizeof(double) * (size_t)S * t->max_points);
    val = (double *)malloc(sizeof(double) * (size_t)S * t->max_points * (size_t)(t->nsave ? t->nsave : 1));
    row = (double *)malloc(sizeof(double) * (size_t)(t->nsave + 1));
    if (!npts || !tm || !val || !row) { r = NGB_E_PANIC; ngb_set_error("out of memory"); goto out; }
    ngb_dev_d2h(npts, t->x.npts, sizeof(int) * (size_t)S);
    if ((r = ngbTranWaves(b, tm, val))) goto out;
    if (!(fp = fopen(path, "wb"))) { ngb_set_error("cannot open %s", path); r = NGB_E_PANIC; goto out; }
    for (s = first_sample; s < first_sample + nsamples; s++) {
        const int n = npts[s] < t->max_points ? npts[s] : t->max_points;     /* points past max_points were not stored */
        int p;
        fprintf(fp, "Title: %s\nDate: %s\nCommand: ngb200, sample %d of %d\nPlotname: Transient Analysis\nFlags: real\n",
                title ? title : "", date, s, S);
        fprintf(fp, "No. Variables: %d\nNo. Points: %-8d\nVariables:\n\t0\ttime\ttime\n", t->nsave + 1, n);
        for (k = 0; k < t->nsave; k++) fprintf(fp, "\t%d\t%s\t%s\n", k + 1, names[k], types[k]);
        fprintf(fp, "Binary:\n");
        for (p = 0; p < n; p++) {
            row[0] = tm[(size_t)s * t->max_points + p];
            for (k = 0; k < t->nsave; k++) row[k + 1] = val[((size_t)s * t->max_points + p) * t->nsave + k];
            if (fwrite(row, sizeof(double), (size_t)t->nsave + 1, fp) != (size_t)t->nsave + 1) { ngb_set_error("write to %s failed", path); r = NGB_E_PANIC; break; }
        }
        if (r) break;
    }
    if (fclose(fp) && !r) { ngb_set_error("write to %s failed", path); r = NGB_E_PANIC; }
out:
    free(npts); free(tm); free(val); free(row);
    return r;
}
long ngbTranTicks(ngb_batch *b) { return b->tran ? b->tran->ticks : 0; }
int ngbTranRepivots(ngb_batch *b) { return b->tran ? b->tran->repivots : 0; }
void *ngbTranDevWaves(ngb_batch *b, int which) { return b->tran ? (which ? (void *)b->tran->x.out_val : (void *)b->tran->x.out_time) : NULL; }

/* `.meas tran` clauses for the next ngbTranRun, evaluated on the device while the points are produced (com_measure_when,
 * src/frontend/com_measure2.c:378-663): clause k watches equation eq[k] for its count[k]-th RISE (kind 0) / FALL (1) /
 * CROSS (2) through val[k], ignoring points before td[k].  n = 0 removes them */
int ngbTranSetMeasures(ngb_batch *b, int n, const int *eq, const int *kind, const int *count, const double *val, const double *td)
{
    int k;
    free(b->ms_eq); free(b->ms_kind); free(b->ms_count); free(b->ms_val); free(b->ms_td);
    b->ms_eq = b->ms_kind = b->ms_count = NULL; b->ms_val = b->ms_td = NULL; b->ms_n = 0;
    if (n <= 0) return NGB_OK;
    for (k = 0; k < n; k++)
        if (eq[k] < 1 || eq[k] >= b->neq1 || kind[k] < 0 || kind[k] > 2 || count[k] < 1) { ngb_set_error("measurement clause %d is malformed", k); return NGB_E_PANIC; }
    b->ms_eq = (int *)malloc(sizeof(int) * (size_t)n); b->ms_kind = (int *)malloc(sizeof(int) * (size_t)n); b->ms_count = (int *)malloc(sizeof(int) * (size_t)n);
    b->ms_val = (double *)malloc(sizeof(double) * (size_t)n); b->ms_td = (double *)malloc(sizeof(double) * (size_t)n);
    if (!b->ms_eq || !b->ms_kind || !b->ms_count || !b->ms_val || !b->ms_td) return NGB_E_PANIC;
    memcpy(b->ms_eq, eq, sizeof(int) * (size_t)n); memcpy(b->ms_kind, kind, sizeof(int) * (size_t)n); memcpy(b->ms_count, count, sizeof(int) * (size_t)n);
    memcpy(b->ms_val, val, sizeof(double) * (size_t)n); memcpy(b->ms_td, td, sizeof(double) * (size_t)n);
    b->ms_n = n;
    return NGB_OK;
}
/* the measured times, out [n][S]; NaN where the transition did not occur (m_measured = NAN, com_measure2.c:659) */
int ngbTranMeasures(ngb_batch *b, double *out)
{
    struct ngb_tran *t = b->tran;
    int m;
    if (!t || !t->x.nmeas) { ngb_set_error("no measurement clauses were set for the last ngbTranRun"); return NGB_E_PANIC; }
    for (m = 0; m < t->x.nmeas; m++)
        ngb_dev_d2h(out + (size_t)m * b->S, t->x.ms_d + ((size_t)m * 3 + 2) * b->S, sizeof(double) * (size_t)b->S);
    return NGB_OK;
}
