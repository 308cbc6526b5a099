/* ngb_tran.cuh -- per-sample transient controller: the control flow of NIiter and DCtran
 * executed on the device, one thread per sample, once per Newton step ("tick").
 *
 * What it mirrors (paths under /root/reference/src):
 *   maths/ni/niiter.c:28-367        iteration count, CKTnoncon / NIconvTest decision, the INITF
 *                                   state machine, the rhs/rhsOld pointer swap
 *   spicelib/analysis/dctran.c:66-971 (non-XSPICE, non-SHARED_MODULE build)
 *                                   initial step, breakpoint table, step rejection on
 *                                   non-convergence (delta/8), CKTtrunc accept/reject with the
 *                                   order-2 trial, delmin guard, state-ring rotation, output
 *   maths/ni/nicomcof.c:14-47       TRAPEZOIDAL coefficients
 *   spicelib/analysis/cktsetbk.c, cktclrbk.c, maths/misc/equality.c  breakpoint helpers
 *
 * The heavy parts of a tick (device loads, LU, solve, LTE estimates, node convergence) are
 * done by the other kernels; this thread only takes the decisions and rewrites the sample's
 * control block for the next tick.  Every sample has its own time axis.
 */
#ifndef NGB_TRAN_CUH
#define NGB_TRAN_CUH
#include "ngb_types.h"

#define NGB_PH_IDLE   0
#define NGB_PH_OPUIC  1      /* the single CKTload of NIiter under MODETRANOP|MODEUIC          */
#define NGB_PH_DCOP   2      /* Newton iteration of the DC operating point (CKTop)             */
#define NGB_PH_TRAN   3      /* Newton iteration of a transient time point                     */
#define NGB_PH_DONE   4
#define NGB_PH_FAIL   5
#define NGB_MAXBRK    16      /* smallest breakpoint table; the transient driver sizes it by the sources that set breakpoints */

typedef struct NgbTranCtx {
    NgbCtl ctl;
    int S, neq1;
    double *x;                 /* [2][neq1][S] */
    const int *nodeconv;       /* [S] node part of NIconvTest from the LU kernel */
    int *nodeconv_w;           /* same array, for the reset */
    /* per-sample controller state, [S] */
    int *phase, *iterno, *firsttime, *nbreak, *npts, *brkflag;
    int *accepted, *rejected, *numiter, *timepts;
    double *save_delta, *old_delta, *breaks;   /* breaks [maxbrk][S] */
    int maxbrk;                /* rows of the breakpoint table: every PULSE / PWL source keeps one pending breakpoint (CKTsetBreak grows
                                * its table without limit, cktsetbk.c), + the two ends + one in flight */
    /* outputs */
    int max_points, nsave;
    const int *save_eq;        /* [nsave] */
    double *out_time;          /* [S][max_points] */
    double *out_val;           /* [S][max_points][nsave] */
    int *evstage;              /* [S] see ngb_ev_advance */
    int *ndone;                /* [0] samples in DONE or FAIL, [1] past the operating point, [2] past the last pivoting event,
                                * [3] samples waiting for the host's pivoting factor (susp == 1) */
    int *susp;                 /* [S] 0; 1: the refactor met a zero pivot, the sample waits (inactive) for the host to factor its matrix
                                * again with pivoting, niiter.c:162-195; 2: that factor found the matrix singular -- NIiter returns E_SINGULAR */
    const int *only;           /* not NULL: this launch handles only the samples with only[s] != 0 (the ones the host has just re-pivoted) */
    int *verify;               /* [S] a pivoting event is due in the sample's next iteration and the LU kernel is to check the event's
                                * recorded pivot order against KLU's rule on the sample's own matrix (NgbLuCtx.verify) */
    int dev_verify;            /* 1: pivoting events are answered that way; only samples that fail the check go to the host */
    /* measurement clauses evaluated while the points are produced (com_measure_when, com_measure2.c:378-663): the n-th
     * RISE / FALL / CROSS of one saved quantity through a constant, linearly interpolated between the two output points
     * around it.  A `.meas tran x TRIG .. TARG ..` is two clauses (result = second - first).  No waveform has to be kept */
    int nmeas;
    const int *ms_eq;          /* [nmeas] equation whose CKTrhsOld value is watched */
    const int *ms_kind;        /* [nmeas] 0 RISE, 1 FALL, 2 CROSS */
    const int *ms_count;       /* [nmeas] which one (1 = first) */
    const double *ms_val;      /* [nmeas] the level (VAL=) */
    const double *ms_td;       /* [nmeas] points before this time are not looked at (TD=) */
    int *ms_i;                 /* [nmeas][4][S] first, section, rise+fall count packed: see ngb_measure_point */
    double *ms_d;              /* [nmeas][3][S] previous value, previous time, result (NaN until found) */
    int *ipass;                /* [S] NIiter's `ipass`: set by every MODEINITFIX iteration; with nodesets in the circuit the first
                                * MODEINITFLOAT iteration of an operating point then counts as not converged (niiter.c:307-331) */
    int had_nodeset;           /* CKThadNodeset (cktic.c:46): some node carries a .nodeset */
    int pivot_events;          /* 1: at every pivoting event of NIiter the sample's own matrix is factored with pivoting by the host
                                * (ctl.lusel = -1: no refactor launch takes the sample, the controller parks it like a zero pivot);
                                * 0: the sample moves to the batch's recorded pattern set of that event */
    /* breakpoint-generating sources (VSRCaccept / ISRCaccept): tables of the load kernels plus the
     * per-sample VSRCbreak_time / ISRCbreak_time, [ninst][S], -1 at setup (vsrcset.c:34) */
    NgbSrcCtx isrc, vsrc;
    double *isrc_break, *vsrc_break;
    /* circuit scalars */
    double tstep, tstop, tmax, tstart, delmin, minbreak, xmu;
    int maxorder, uic, max_iter_tran, max_iter_dc;
    /* dynamic gmin stepping (cktop.c:162-274), the operating-point fallback when plain Newton fails:
     * gm_stage 0 = plain NIiter, 1 = inside the stepping loop, 2 = final NIiter with CKTdiagGmin = gshunt */
    int *gm_stage;             /* [S] */
    double *gm_factor, *gm_oldgmin;   /* [S] */
    /* the rest of CKTop's chain (cktop.c:62-96): gm_stage 3 / 4 = new_gmin's loop (it steps CKTgmin itself, :350-463) and
     * its final NIiter; 10 / 11 / 12 = gillespie_src (:481-660): first solve with the sources at zero, its diagonal-gmin
     * ladder when that fails, the source-raising loop */
    double *gm_startgmin, *gs_conv, *gs_raise;   /* [S] CKTgmin on entry of new_gmin; ConvFact; raise */
    int *gs_i;                        /* [S] step of the ladder */
    int num_gmin_steps, num_src_steps, itl2;     /* CKTnumGminSteps, CKTnumSrcSteps (0: skip, 1: the routes built here), CKTdcTrcvMaxIter */
    double gmin_factor, gshunt;       /* CKTgminFactor, CKTgshunt */
    double *gm_xold;           /* [neq1][S] OldRhsOld */
    struct { double *state, *old; int K, ninst; } gm_arr[5];   /* device state tables and their OldCKTstate0 copies [K][ninst*S] */
    int gm_narr, gm_enable;
    int nluset;                /* > 1: samples move between pattern sets at the pivoting events */
    int lu_event[NGB_LU_EVENTS];   /* pattern set of each pivoting event (ngb_types.h) */
} NgbTranCtx;

NGB_HD int ngb_almost_equal_ulps(double A, double B, int maxUlps)
{
    long long a, b, d;
    if (A == B) return 1;
#ifdef __CUDA_ARCH__
    a = __double_as_longlong(A); b = __double_as_longlong(B);
#else
    union { double d; long long i; } ua, ub; ua.d = A; ub.d = B; a = ua.i; b = ub.i;
#endif
    if (a < 0) a = (long long)0x8000000000000000ULL - a;
    if (b < 0) b = (long long)0x8000000000000000ULL - b;
    d = a - b; if (d < 0) d = -d;
    return d <= maxUlps;
}

#define TBRK(i) c->breaks[(size_t)(i) * S + s]

/* CKTclrBreak */
NGB_HD void ngb_clr_break(const NgbTranCtx *c, int s)
{
    const int S = c->S;
    int nb = c->nbreak[s];
    if (nb > 2) {
        for (int j = 1; j < nb; j++) TBRK(j - 1) = TBRK(j);
        c->nbreak[s] = nb - 1;
    } else {
        TBRK(0) = TBRK(1);
        TBRK(1) = c->tstop;
    }
}

/* CKTsetBreak */
NGB_HD int ngb_set_break(const NgbTranCtx *c, int s, double time, double now)
{
    const int S = c->S;
    int nb = c->nbreak[s];
    if (ngb_almost_equal_ulps(time, now, 3)) return NGB_OK;
    if (now > time) return NGB_E_PANIC;
    for (int i = 0; i < nb; i++) {
        if (TBRK(i) > time) {
            if ((TBRK(i) - time) <= c->minbreak) { TBRK(i) = time; return NGB_OK; }
            if (i > 0 && time - TBRK(i - 1) <= c->minbreak) return NGB_OK;
            if (nb >= c->maxbrk) return NGB_E_PANIC;
            for (int j = nb; j > i; j--) TBRK(j) = TBRK(j - 1);
            TBRK(i) = time;
            c->nbreak[s] = nb + 1;
            return NGB_OK;
        }
    }
    if (time - TBRK(nb - 1) <= c->minbreak) return NGB_OK;
    if (nb >= c->maxbrk) return NGB_E_PANIC;
    TBRK(nb) = time;
    c->nbreak[s] = nb + 1;
    return NGB_OK;
}

/* NIcomCof (nicomcof.c:20-121): TRAPEZOIDAL in closed form; GEAR solves the (order+1)-square system the reference
 * sets up -- powers of (sum of the last steps)/delta -- by the same elimination, in the same order */
NGB_HD void ngb_comcof(const NgbTranCtx *c, int s, int order, double delta)
{
    if (c->ctl.gear) {
        double mat[3][3] = { { 0 } }, ag[3] = { 0, 0, 0 }, arg = 0, arg1;
        int i, j, k;
        const int S = c->S;
        ag[1] = -1 / delta;
        for (i = 0; i <= order; i++) mat[0][i] = 1;
        for (i = 1; i <= order; i++) mat[i][0] = 0;
        for (i = 1; i <= order; i++) {
            arg += c->ctl.delta_old[(size_t)(i - 1) * S + s];
            arg1 = 1;
            for (j = 1; j <= order; j++) { arg1 *= arg / delta; mat[j][i] = arg1; }
        }
        for (i = 1; i <= order; i++)
            for (j = i + 1; j <= order; j++) {
                mat[j][i] /= mat[i][i];
                for (k = i + 1; k <= order; k++) mat[j][k] -= mat[j][i] * mat[i][k];
            }
        for (i = 1; i <= order; i++)
            for (j = i + 1; j <= order; j++) ag[j] = ag[j] - mat[j][i] * ag[i];
        ag[order] /= mat[order][order];
        for (i = order - 1; i >= 0; i--) {
            for (j = i + 1; j <= order; j++) ag[i] = ag[i] - mat[i][j] * ag[j];
            ag[i] /= mat[i][i];
        }
        c->ctl.ag0[s] = ag[0]; c->ctl.ag1[s] = ag[1]; c->ctl.ag2[s] = ag[2];
        return;
    }
    if (order == 1) {
        c->ctl.ag0[s] = 1 / delta;
        c->ctl.ag1[s] = -1 / delta;
    } else {
        c->ctl.ag0[s] = 1.0 / delta / (1.0 - c->xmu);
        c->ctl.ag1[s] = c->xmu / (1.0 - c->xmu);
    }
}

/* start the Newton iteration of the next attempt at a time point: top of the for(;;) of
 * dctran.c:666-707 */
NGB_HD void ngb_begin_point(const NgbTranCtx *c, int s)
{
    const int S = c->S;
    const double delta = c->ctl.delta[s];
    c->old_delta[s] = delta;
    c->ctl.time[s] += delta;
    c->ctl.delta_old[(size_t)0 * S + s] = delta;
    ngb_comcof(c, s, c->ctl.order[s], delta);
    c->iterno[s] = 0;
    c->phase[s] = NGB_PH_TRAN;
}

/* pivoting-event bookkeeping: evstage[s] = 1 once the sample has left the operating point, 2 once it is
 * past the last re-pivoting of its run; ndone[1], ndone[2] count the samples at or beyond each stage so
 * that the host can stop launching pattern sets nobody is on */
NGB_HD void ngb_ev_advance(const NgbTranCtx *c, int s, int to)
{
    while (c->evstage[s] < to) {
        const int st = ++c->evstage[s];
#ifdef __CUDA_ARCH__
        atomicAdd(c->ndone + st, 1);
#else
        c->ndone[st] += 1;
#endif
    }
}

NGB_HD void ngb_finish(const NgbTranCtx *c, int s, int phase, int err)
{
    ngb_ev_advance(c, s, 2);
    c->phase[s] = phase;
    c->ctl.active[s] = 0;
    if (err) c->ctl.err[s] = err;
#ifdef __CUDA_ARCH__
    atomicAdd(c->ndone, 1);
#else
    c->ndone[0] += 1;
#endif
}

/* VSRCaccept / ISRCaccept (vsrc/vsrcacct.c:22-152, isrc/isrcacct.c): PULSE sources ask for a
 * breakpoint at their next corner; DC, SIN, EXP, SFFM, AM never do */
NGB_HD int ngb_src_accept(const NgbTranCtx *c, const NgbSrcCtx *sc, double *brk, int poff, int s, double now)
{
    const int S = c->S;
    for (int inst = 0; inst < sc->ninst; inst++) {
        const int ftype = NGB_LDG(&sc->fn[inst]);
        const int forder = NGB_LDG(&sc->fn[sc->ninst + inst]);
        const size_t t = (size_t)inst * S + s;
#define SCO(k) NGB_LDG(&sc->par[(size_t)(poff + (k)) * sc->T + t])
        if (ftype == NGB_FN_PULSE) {
            const double TD = forder > 2 ? SCO(2) : 0.0;
            const double TR = (forder > 3 && SCO(3) > 0.0) ? SCO(3) : c->tstep;
            const double TF = (forder > 4 && SCO(4) > 0.0) ? SCO(4) : c->tstep;
            const double PW = (forder > 5 && SCO(5) >= 0.0) ? SCO(5) : 0.0;
            const double PER = (forder > 6 && SCO(6) > 0.0) ? SCO(6) : TR + TF + PW;
            const double PHASE = forder > 7 ? SCO(7) : 0.0;
            double time = now - TD;
            if (PHASE > 0.0 && time > PHASE * PER) continue;
            if (now >= brk[t]) {
                double wait, atime;
                if (time >= PER) time -= PER * floor(time / PER);
                atime = time + c->minbreak;
                if (atime < 0.0) wait = -time;
                else if (atime < TR) wait = TR - time;
                else if (atime < TR + PW) wait = TR + PW - time;
                else if (atime < TR + PW + TF) wait = TR + PW + TF - time;
                else wait = PER - time;
                brk[t] = now + wait;
                { const int e = ngb_set_break(c, s, brk[t], now); if (e) return e; }
                brk[t] -= c->minbreak;
            }
        } else if (ftype == NGB_FN_PWL && sc->is_current) {     /* isrcacct.c:181-197: only on a breakpoint (CKTbreak) */
            const double *co = sc->pwl + NGB_LDG(&sc->pwl_ptr[inst]);
            if (!c->brkflag[s]) continue;
            if (now < NGB_LDG(&co[0])) {
                const int e = ngb_set_break(c, s, NGB_LDG(&co[0]), now);       /* the reference drops this error code */
                (void)e;
                continue;
            }
            for (int i = 0; i < forder / 2 - 1; i++)
                if (ngb_almost_equal_ulps(NGB_LDG(&co[2 * i]), now, 3)) {
                    const int e = ngb_set_break(c, s, NGB_LDG(&co[2 * i + 2]), now);
                    if (e) return e;
                    break;
                }
        } else if (ftype == NGB_FN_PWL) {       /* vsrcacct.c:174-226 */
            if (now >= brk[t]) {
                const double *co = sc->pwl + NGB_LDG(&sc->pwl_ptr[inst]);
                const int rep = NGB_LDG(&sc->pwl_rep[inst]);
                double time = now - NGB_LDG(&sc->pwl_rdelay[inst]);
                const double end = NGB_LDG(&co[forder - 2]);
                if (time > end) {
                    if (rep >= 0) {
                        const double period = end - NGB_LDG(&co[rep]);
                        time -= NGB_LDG(&co[rep]);
                        time -= period * floor(time / period);
                        time += NGB_LDG(&co[rep]);
                    } else { brk[t] = c->tstop; continue; }
                }
                const double atime = time + c->minbreak;
                for (int i = 0; i < forder; i += 2)
                    if (NGB_LDG(&co[i]) > atime) {
                        brk[t] = now + NGB_LDG(&co[i]) - time;
                        { const int e = ngb_set_break(c, s, brk[t], now); if (e) return e; }
                        brk[t] -= c->minbreak;
                        break;
                    }
            }
        } else if (ftype != 0 && ftype != NGB_FN_SINE && ftype != NGB_FN_EXP && ftype != NGB_FN_SFFM && ftype != NGB_FN_AM) {
            return NGB_E_UNSUPP;
        }
#undef SCO
    }
    return NGB_OK;
}

/* one output point (time, value of the watched equation) for measurement clause m of sample s: the loop body of
 * com_measure_when for a real vector against a constant (com_measure2.c:455-660, the branch without a second vector);
 * `first` counts the points looked at: the second one only initialises the side (and may count a transition without
 * measuring it -- mirrored), later ones count transitions and take the measurement when the requested count is met */
NGB_HD void ngb_measure_point(const NgbTranCtx *c, int m, int s, double scale, double value)
{
    const int S = c->S;
    int *mi = c->ms_i + (size_t)m * 4 * S;
    double *md = c->ms_d + (size_t)m * 3 * S;
    int first = mi[s], section = mi[(size_t)S + s], rise = mi[(size_t)2 * S + s], fall = mi[(size_t)3 * S + s];
    const double val = c->ms_val[m];
    const double prevValue = md[s], prevScale = md[(size_t)S + s];
    if (first < 0) return;                                   /* measured */
    if (scale < c->ms_td[m]) return;
    if (first == 1) {
        rise = fall = 0;
        if (value < val) { section = 0; if (prevValue >= val) fall = 1; }
        else { section = 1; if (prevValue < val) rise = 1; }
    }
    if (first > 1) {
        if (section == 0 && value >= val) { section = 1; rise++; }
        else if (section == 1 && value <= val) { section = 0; fall++; }
        const int kind = c->ms_kind[m], want = c->ms_count[m];
        const int have = kind == 0 ? rise : (kind == 1 ? fall : rise + fall);
        if (have == want) {
            md[(size_t)2 * S + s] = prevScale + (val - prevValue) * (scale - prevScale) / (value - prevValue);
            mi[s] = -1;
            return;
        }
    }
    mi[s] = first + 1; mi[(size_t)S + s] = section; mi[(size_t)2 * S + s] = rise; mi[(size_t)3 * S + s] = fall;
    md[s] = value; md[(size_t)S + s] = scale;
}

/* the nextTime: label of dctran.c:355-665 -- accept the point, output, breakpoints, rotate */
NGB_HD void ngb_next_time(const NgbTranCtx *c, int s)
{
    const int S = c->S;
    const double time = c->ctl.time[s];
    const int mode = c->ctl.mode[s];
    double delta;
    /* CKTaccept: device types in DEVices order, isrc before vsrc (dev.c:142-209) */
    {
        int e = ngb_src_accept(c, &c->isrc, c->isrc_break, 2, s, time);
        if (!e) e = ngb_src_accept(c, &c->vsrc, c->vsrc_break, 1, s, time);
        if (e) { ngb_finish(c, s, NGB_PH_FAIL, e); return; }
    }
    if (time > TBRK(0)) ngb_clr_break(c, s);
    c->accepted[s] += 1;
    c->brkflag[s] = 0;
    /* CKTdump of CKTrhsOld */
    if (((mode & NGB_MODEUIC) && time > 0 && time >= c->tstart) || (!(mode & NGB_MODEUIC) && time >= c->tstart)) {
        const int n = c->npts[s];
        if (n < c->max_points) {
            const double *xo = c->x + (size_t)c->ctl.xsel[s] * c->neq1 * S;
            c->out_time[(size_t)s * c->max_points + n] = time;
            for (int k = 0; k < c->nsave; k++)
                c->out_val[((size_t)s * c->max_points + n) * c->nsave + k] = xo[(size_t)c->save_eq[k] * S + s];
        }
        c->npts[s] = n + 1;
        if (c->nmeas) {
            const double *xo = c->x + (size_t)c->ctl.xsel[s] * c->neq1 * S;
            for (int m = 0; m < c->nmeas; m++) ngb_measure_point(c, m, s, time, xo[(size_t)c->ms_eq[m] * S + s]);
        }
    }
    if (ngb_almost_equal_ulps(time, c->tstop, 100)) { ngb_finish(c, s, NGB_PH_DONE, 0); return; }

    /* resume: */
    delta = c->ctl.delta[s];
    delta = NGB_MIN(delta, c->tmax);
    if (ngb_almost_equal_ulps(time, TBRK(0), 100) || TBRK(0) - time <= c->delmin) {
        double lim = NGB_MIN(c->save_delta[s], TBRK(1) - TBRK(0));
        c->ctl.order[s] = 1;
        lim = .1 * lim;
        delta = NGB_MIN(delta, lim);
        if (c->firsttime[s]) {
            if (mode & NGB_MODEUIC) ngb_set_break(c, s, c->tstep, time);
            delta /= 10;
        }
        { const double lo = c->delmin * 2.0; delta = NGB_MAX(delta, lo); }
    } else if (time + delta >= TBRK(0)) {
        c->save_delta[s] = delta;
        delta = TBRK(0) - time;
        c->brkflag[s] = 1;
    } else if (time + 1.9 * delta > TBRK(0)) {
        c->save_delta[s] = delta;
        delta = (TBRK(0) - time) / 2.;
    }
    c->ctl.delta[s] = delta;
    for (int i = 5; i >= 0; i--)
        c->ctl.delta_old[(size_t)(i + 1) * S + s] = c->ctl.delta_old[(size_t)i * S + s];
    c->ctl.delta_old[(size_t)0 * S + s] = delta;
    /* rotate the state ring: states[i+1] = states[i], states[0] = old states[maxOrder+1] */
    {
        const int nh = c->ctl.nhist;
        const int sop = c->ctl.stateop[s];
        c->ctl.head[s] = (c->ctl.head[s] + nh - 1) % nh;
        /* pending copies seen from the rotated frame: state1 = state0 is what the rotation itself
         * does; state2 = state1, state3 = state1 leaves state3 = state2 and state0 = state2 to do (the old state3 is the
         * new state0: its stale content is what DIOload's limiting of a separate sidewall diode reads under MODEINITPRED,
         * dioload.c:177-192 copies DIOvoltage but not DIOvoltageSW) */
        c->ctl.stateop[s] = (sop & NGB_OP_COPY1_23) ? NGB_OP_COPY23 : 0;
    }
    ngb_begin_point(c, s);
}

/* the pattern set a pivoting event sends the sample to: without per-sample pivoting the set the recorded run's factor produced;
 * with it either that set plus the request to verify its order on the sample's matrix in the LU kernel, or -1 (the sample waits
 * for the host's pivoting factor) */
#define NGB_EVENT_LUSEL(c, s, e) (!(c)->pivot_events ? (c)->lu_event[e] : ((c)->dev_verify ? ((c)->verify[s] = 1, (c)->lu_event[e]) : -1))

/* CKTstate0 and CKTrhsOld of sample s: op 0 zero them, 1 save to the Old copies, 2 restore from them (cktop.c:182-186, 210-214, 244-248) */
NGB_HD void ngb_gm_states(const NgbTranCtx *c, int s, int op)
{
    const int S = c->S, nh = c->ctl.nhist, head = c->ctl.head[s];
    double *xo = c->x + (size_t)c->ctl.xsel[s] * c->neq1 * S;
    for (int i = 1; i < c->neq1; i++) {
        const size_t k = (size_t)i * S + s;
        if (op == 0) xo[k] = 0.0; else if (op == 1) c->gm_xold[k] = xo[k]; else xo[k] = c->gm_xold[k];
    }
    for (int a = 0; a < c->gm_narr; a++) {
        const size_t T = (size_t)c->gm_arr[a].ninst * S;
        const int K = c->gm_arr[a].K;
        double *st = c->gm_arr[a].state + (size_t)(head % nh) * K * T;
        double *old = c->gm_arr[a].old;
        for (int k = 0; k < K; k++)
            for (int inst = 0; inst < c->gm_arr[a].ninst; inst++) {
                const size_t j = (size_t)k * T + (size_t)inst * S + s;
                if (op == 0) st[j] = 0.0; else if (op == 1) old[j] = st[j]; else st[j] = old[j];
            }
    }
}

/* start the next NIiter call of the operating point for sample s */
NGB_HD void ngb_gm_next_niiter(const NgbTranCtx *c, int s, int mode)
{
    c->ctl.mode[s] = mode;
    c->iterno[s] = 0;
    c->ipass[s] = 0;
    if ((mode & NGB_MODEINITJCT) && c->nluset > 1) c->ctl.lusel[s] = NGB_EVENT_LUSEL(c, s, 0);
}

/* One controller step for sample s, after the load (+ LU + solve) of this tick. */
NGB_HD void ngb_tran_control(const NgbTranCtx *c, int s)
{
    const int S = c->S;
    int phase = c->phase[s];
    if (phase == NGB_PH_IDLE || phase == NGB_PH_DONE || phase == NGB_PH_FAIL) return;
    if (c->only ? !c->only[s] : (c->susp && c->susp[s] == 1)) return;
    if (c->verify) c->verify[s] = 0;         /* the check, if one was due, ran in this step's LU launch */
    /* state copies requested for the load that just ran are done */
    const int sop_done = c->ctl.stateop[s];
    c->ctl.stateop[s] = 0;
    (void)sop_done;

    int forced = -1;                         /* NIiter's return value when SMPluFac / SMPreorder failed */
    if (c->pivot_events && !c->only && !c->ctl.err[s] && phase != NGB_PH_OPUIC && c->ctl.lusel[s] == -1) {
        /* SMPreorder is due in this iteration (NISHOULDREORDER): the load and the assembly of this step are done, no
         * refactor launch took the sample; it waits for the host's pivoting factor of its matrix */
        c->susp[s] = 1; c->ctl.active[s] = 0;
#ifdef __CUDA_ARCH__
        atomicAdd(c->ndone + 3, 1);
#else
        c->ndone[3] += 1;
#endif
        c->ctl.stateop[s] = sop_done;        /* nothing of this step is consumed yet */
        return;
    }
    if (c->ctl.err[s]) {
        if (c->ctl.err[s] != NGB_E_SINGULAR || !c->susp) { ngb_finish(c, s, NGB_PH_FAIL, c->ctl.err[s]); return; }
        if (c->susp[s] == 0 && !c->only) {
            /* zero pivot in the refactor: the reference factors the same matrix again with pivoting and goes on with the
             * iteration.  The sample stands still until the host has done that (ngb_tran.c: repivot_suspended) */
            c->ctl.err[s] = 0; c->susp[s] = 1; c->ctl.active[s] = 0;
#ifdef __CUDA_ARCH__
            atomicAdd(c->ndone + 3, 1);
#else
            c->ndone[3] += 1;
#endif
            return;
        }
        /* the pivoting factor failed as well: "seems to be singular - pass the bad news up" (niiter.c:176-190) */
        c->ctl.err[s] = 0; c->susp[s] = 0; forced = NGB_E_SINGULAR;
    }

    if (phase == NGB_PH_OPUIC) {
        /* CKTop returned OK after one CKTload; DCtran sets up the transient (dctran.c:271-330) */
        c->timepts[s] += 1;
        c->ctl.order[s] = 1;
        for (int i = 0; i < 7; i++) c->ctl.delta_old[(size_t)i * S + s] = c->tmax;
        c->ctl.delta[s] = NGB_MIN(c->tstop / 100, c->tstep) / 10;
        c->save_delta[s] = c->tstop / 50;
        c->ctl.mode[s] = (c->ctl.mode[s] & NGB_MODEUIC) | NGB_MODETRAN | NGB_MODEINITTRAN;
        c->ctl.ag0[s] = 0; c->ctl.ag1[s] = 0;
        c->ctl.stateop[s] = NGB_OP_COPY01;
        c->ctl.noncon[s] = 0; c->nodeconv_w[s] = 0; c->ctl.lte[s] = 1e300; c->ctl.lte2[s] = 1e300;
        if (c->nluset > 1) c->ctl.lusel[s] = NGB_EVENT_LUSEL(c, s, 2);        /* the first factor of the run is still ahead */
        ngb_ev_advance(c, s, 1);
        ngb_next_time(c, s);
        return;
    }

    /* ---- NIiter after SMPsolve (niiter.c:254-362) ---- */
    int iterno = c->iterno[s] + 1;
    int mode = c->ctl.mode[s];
    int noncon = c->ctl.noncon[s];
    /* NIiter raises maxIter to 100 (niiter.c:37); the gmin steps run with CKTdcTrcvMaxIter (itl2, default 50) */
    /* the steps of the ladders (dynamic_gmin 1, new_gmin 3, gillespie_src 10-12) run with CKTdcTrcvMaxIter (itl2), the plain NIiter
     * (0) and the closing NIiter of the gmin ladders (2, 4) with CKTdcMaxIter (itl1): cktop.c:34, 201, 259, 388, 447, 506-585 */
    const int gst = (phase == NGB_PH_DCOP && c->gm_stage) ? c->gm_stage[s] : 0;
    const int maxiter = (phase == NGB_PH_DCOP) ? ((gst == 1 || gst == 3 || gst >= 10) ? NGB_MAX(c->itl2, 100) : NGB_MAX(c->max_iter_dc, 100))
                                               : NGB_MAX(c->max_iter_tran, 100);
    int niret = -1;                          /* -1: keep iterating, 0: converged, >0: error */
    if (forced >= 0) {
        iterno -= 1;                          /* the failed iteration is not counted (STATnumIter += iterno before iterno++) */
        niret = forced;
    } else if ((c->iterno[s] = iterno) > maxiter) {
        niret = NGB_E_ITERLIM;
    } else {
        if ((noncon == 0) && (iterno != 1)) noncon = c->nodeconv[s] ? 1 : 0;   /* NIconvTest */
        else noncon = 1;
        if (mode & NGB_MODEINITFLOAT) {
            if ((mode & NGB_MODEDC) && c->had_nodeset) {
                if (c->ipass[s]) noncon = c->ipass[s];
                c->ipass[s] = 0;
            }
            if (noncon == 0) niret = NGB_OK;
        } else if (mode & NGB_MODEINITJCT) {
            mode = (mode & ~NGB_INITF) | NGB_MODEINITFIX;
            if (c->nluset > 1) c->ctl.lusel[s] = NGB_EVENT_LUSEL(c, s, 1);           /* NISHOULDREORDER, niiter.c:335 */
        } else if (mode & NGB_MODEINITFIX) {
            if (noncon == 0) mode = (mode & ~NGB_INITF) | NGB_MODEINITFLOAT;
            c->ipass[s] = 1;
        } else if (mode & (NGB_MODEINITTRAN | NGB_MODEINITPRED | NGB_MODEINITSMSIG)) {
            if ((mode & NGB_MODEINITTRAN) && iterno <= 1) {
                if (c->nluset > 1) c->ctl.lusel[s] = NGB_EVENT_LUSEL(c, s, 3);       /* NISHOULDREORDER, niiter.c:343-344 */
                ngb_ev_advance(c, s, 2);
            }
            mode = (mode & ~NGB_INITF) | NGB_MODEINITFLOAT;
        } else {
            niret = NGB_E_PANIC;
        }
    }
#ifdef NGB_TRAN_DEBUG_PRINT
    fprintf(stderr, "DBG s=%d t=%.17g it=%d mode=%x noncon_in=%d noncon=%d niret=%d order=%d delta=%.17g\n", s, c->ctl.time[s], iterno, c->ctl.mode[s], c->ctl.noncon[s], noncon, niret, c->ctl.order[s], c->ctl.delta[s]);
#endif
    /* reset the per-tick flags for the next load */
    const double lte = c->ctl.lte[s], lte2 = c->ctl.lte2[s];
    c->ctl.noncon[s] = 0; c->nodeconv_w[s] = 0; c->ctl.lte[s] = 1e300; c->ctl.lte2[s] = 1e300;

    if (niret < 0) {
        c->ctl.mode[s] = mode;
        c->ctl.xsel[s] ^= 1;                 /* SWAP(CKTrhs, CKTrhsOld) */
        return;
    }
    c->numiter[s] += iterno;

    if (phase == NGB_PH_DCOP && c->gm_enable) {
        /* CKTop (cktop.c:27-112): plain NIiter; then, with CKTnumGminSteps == 1, dynamic_gmin (:162-274) and new_gmin
         * (:350-463); then, with CKTnumSrcSteps == 1, gillespie_src (:481-660).  spice3_gmin / spice3_src (counts > 1) are
         * refused by ngbCircuitSetOpFallbacks, OPtran is not built: a sample none of them brings home fails with the
         * last NIiter's code */
        const int firstmode = (mode & NGB_MODEUIC) | NGB_MODETRANOP | NGB_MODEINITJCT;
        const int contmode = (mode & NGB_MODEUIC) | NGB_MODETRANOP | NGB_MODEINITFLOAT;
        const int stage = c->gm_stage[s];
        const int itl2 = c->itl2;                               /* CKTdcTrcvMaxIter */
        const double gmin_factor = c->gmin_factor;              /* CKTgminFactor */
        int start = 0;                                          /* the fallback that begins now: 1 dynamic_gmin, 3 new_gmin, 10 gillespie_src */
        int raise_init = 0;
        if (niret != NGB_OK) {
            if (stage == 0) start = c->num_gmin_steps == 1 ? 1 : (c->num_src_steps == 1 ? 10 : 0);
            else if (stage == 2) start = 3;
            else if (stage == 4) start = c->num_src_steps == 1 ? 10 : 0;
        }
        if (start == 1) {
            c->gm_stage[s] = 1;
            ngb_gm_states(c, s, 0);
            c->gm_factor[s] = gmin_factor;
            c->gm_oldgmin[s] = 1e-2;
            c->ctl.diag_gmin[s] = 1e-2 / gmin_factor;
            ngb_gm_next_niiter(c, s, firstmode);
            return;
        }
        if (start == 3) {
            c->gm_stage[s] = 3;
            ngb_gm_states(c, s, 0);
            c->gm_startgmin[s] = c->ctl.gmin[s];
            c->gm_factor[s] = gmin_factor;
            c->gm_oldgmin[s] = 1e-2;
            c->ctl.gmin[s] = 1e-2 / gmin_factor;
            ngb_gm_next_niiter(c, s, firstmode);
            return;
        }
        if (start == 10) {
            c->gm_stage[s] = 10;
            c->ctl.srcfact[s] = 0.0;
            c->gs_conv[s] = 0.0;
            ngb_gm_states(c, s, 0);
            ngb_gm_next_niiter(c, s, firstmode);
            return;
        }
        if (stage == 1 || stage == 3) {
            /* the two gmin ladders differ in what they step (CKTdiagGmin through LoadGmin / CKTgmin inside the device
             * models), in the floor of the shrinking factor and in the value they leave behind */
            double *g = stage == 1 ? &c->ctl.diag_gmin[s] : &c->ctl.gmin[s];
            const double g0 = stage == 1 ? c->ctl.gmin[s] : c->gm_startgmin[s];
            const double gtarget = NGB_MAX(g0, c->gshunt);                       /* MAX(CKTgmin, CKTgshunt) */
            double factor = c->gm_factor[s];
            int leave = 0;
            if (niret == NGB_OK) {
                mode = contmode;
                if (*g <= gtarget) {
                    leave = 1;
                } else {
                    ngb_gm_states(c, s, 1);
                    if (iterno <= itl2 / 4) { factor *= sqrt(factor); if (factor > gmin_factor) factor = gmin_factor; }
                    if (iterno > 3 * itl2 / 4) factor = NGB_MAX(sqrt(factor), stage == 1 ? 1.00005 : 3);
                    c->gm_oldgmin[s] = *g;
                    if (*g < factor * gtarget) { factor = *g / gtarget; *g = gtarget; }
                    else *g /= factor;
                }
            } else if (factor < 1.00005) {
                leave = 1;                                       /* "Last gmin step failed" */
            } else {
                factor = sqrt(sqrt(factor));
                *g = c->gm_oldgmin[s] / factor;
                ngb_gm_states(c, s, 2);
            }
            c->gm_factor[s] = factor;
            if (leave) { *g = stage == 1 ? c->gshunt : gtarget; c->gm_stage[s] = stage + 1; }   /* CKTdiagGmin = CKTgshunt / CKTgmin = MAX(start value, CKTgshunt); final NIiter */
            ngb_gm_next_niiter(c, s, mode);
            return;
        }
        if (stage == 10) {
            if (niret != NGB_OK) {                              /* the ladder: CKTdiagGmin from 1e10 * gmin down, eleven steps */
                double dg = (c->gshunt <= 0) ? c->ctl.gmin[s] : c->gshunt;
                for (int i = 0; i < 10; i++) dg *= 10;
                c->ctl.diag_gmin[s] = dg;
                c->gs_i[s] = 0;
                c->gm_stage[s] = 11;
                ngb_gm_next_niiter(c, s, mode);
                return;
            }
            raise_init = 1;
        } else if (stage == 11) {
            if (niret != NGB_OK) {                              /* "gmin step failed": no solution at zero sources */
                c->ctl.diag_gmin[s] = c->gshunt; c->ctl.srcfact[s] = 1.0;
                ngb_finish(c, s, NGB_PH_FAIL, NGB_E_ITERLIM);
                return;
            }
            c->ctl.diag_gmin[s] /= 10;
            mode = contmode;
            c->gs_i[s] += 1;
            if (c->gs_i[s] <= 10) { ngb_gm_next_niiter(c, s, mode); return; }
            c->ctl.diag_gmin[s] = c->gshunt;
            raise_init = 1;
        }
        if (raise_init) {
            ngb_gm_states(c, s, 1);
            c->gs_raise[s] = 0.001;
            c->ctl.srcfact[s] = c->gs_conv[s] + 0.001;
            c->gm_stage[s] = 12;
            ngb_gm_next_niiter(c, s, mode);
            return;
        }
        if (stage == 12) {
            double conv = c->gs_conv[s], raise = c->gs_raise[s], sf = c->ctl.srcfact[s];
            int stop = 0;
            mode = contmode;
            if (niret == NGB_OK) {
                conv = sf;
                ngb_gm_states(c, s, 1);
                sf = conv + raise;
                if (iterno <= itl2 / 4) raise *= 1.5;
                if (iterno > 3 * itl2 / 4) raise *= 0.5;
            } else if (sf - conv < 1e-8) {
                stop = 1;
            } else {
                raise /= 10;
                if (raise > 0.01) raise = 0.01;
                sf = conv;
                ngb_gm_states(c, s, 2);
            }
            if (sf > 1) sf = 1;
            c->gs_conv[s] = conv; c->gs_raise[s] = raise; c->ctl.srcfact[s] = sf;
            if (!stop && raise >= 1e-7 && conv < 1) { ngb_gm_next_niiter(c, s, mode); return; }
            c->ctl.diag_gmin[s] = c->ctl.gmin[s];                /* "CKTdiagGmin = CKTgmin = gminstart" (:645): it stays for the run */
            c->ctl.srcfact[s] = 1.0;
            c->gm_stage[s] = 13;
            if (conv != 1) { ngb_finish(c, s, NGB_PH_FAIL, NGB_E_ITERLIM); return; }
            niret = NGB_OK;                                      /* "Source stepping completed": the operating point stands */
        }
    }
    if (phase == NGB_PH_DCOP) {
        if (niret != NGB_OK) { ngb_finish(c, s, NGB_PH_FAIL, niret); return; }
        c->timepts[s] += 1;
        c->ctl.order[s] = 1;
        for (int i = 0; i < 7; i++) c->ctl.delta_old[(size_t)i * S + s] = c->tmax;
        c->ctl.delta[s] = NGB_MIN(c->tstop / 100, c->tstep) / 10;
        c->save_delta[s] = c->tstop / 50;
        c->ctl.mode[s] = (mode & NGB_MODEUIC) | NGB_MODETRAN | NGB_MODEINITTRAN;
        c->ctl.ag0[s] = 0; c->ctl.ag1[s] = 0;
        c->ctl.stateop[s] = NGB_OP_COPY01;
        /* NIiter re-pivots in the first iteration under MODEINITTRAN (niiter.c:107-111) */
        if (c->nluset > 1) c->ctl.lusel[s] = NGB_EVENT_LUSEL(c, s, 2);
        ngb_ev_advance(c, s, 1);
        ngb_next_time(c, s);
        return;
    }

    /* ---- DCtran after NIiter (dctran.c:708-913) ---- */
    {
        const int firsttime = c->firsttime[s];
        double delta = c->ctl.delta[s];
        c->timepts[s] += 1;
        c->ctl.mode[s] = (mode & NGB_MODEUIC) | NGB_MODETRAN | NGB_MODEINITPRED;
        if (firsttime) c->ctl.stateop[s] |= NGB_OP_COPY1_23;
        if (niret != NGB_OK) {
            c->ctl.time[s] -= delta;
            c->rejected[s] += 1;
            delta = delta / 8;
            if (firsttime) c->ctl.mode[s] = (mode & NGB_MODEUIC) | NGB_MODETRAN | NGB_MODEINITTRAN;
            c->ctl.order[s] = 1;
        } else {
            if (firsttime) {
                c->firsttime[s] = 0;
                ngb_next_time(c, s);
                return;
            }
            /* CKTtrunc: *timeStep = MIN(2 * *timeStep, timetemp) */
            double newdelta = NGB_MIN(2 * delta, lte);
            if (newdelta > .9 * delta) {
                if ((c->ctl.order[s] == 1) && (c->maxorder > 1)) {
                    newdelta = NGB_MIN(2 * delta, lte2);
                    c->ctl.order[s] = 2;
                    if (newdelta <= 1.05 * delta) c->ctl.order[s] = 1;
                }
                c->ctl.delta[s] = newdelta;
                ngb_next_time(c, s);
                return;
            }
            c->ctl.time[s] -= delta;
            c->rejected[s] += 1;
            delta = newdelta;
        }
        if (delta <= c->delmin) {
            if (c->old_delta[s] > c->delmin) delta = c->delmin;
            else { ngb_finish(c, s, NGB_PH_FAIL, NGB_E_TIMESTEP); return; }
        }
        c->ctl.delta[s] = delta;
        ngb_begin_point(c, s);
    }
}
#undef TBRK
#endif
