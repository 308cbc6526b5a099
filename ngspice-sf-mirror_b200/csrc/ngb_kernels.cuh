/* ngb_kernels.cuh -- kernel bodies other than the BSIM4 evaluation: linear elements and
 * sources, matrix/RHS assembly, numeric LU refactor + triangular solves + node convergence.
 *
 * Reference behaviour restated here (file:line under /root/reference/src):
 *   capacitor load      spicelib/devices/cap/capload.c:15-110
 *   V/I source load     spicelib/devices/vsrc/vsrcload.c:27-470, isrc/isrcload.c:27-380 (DC, PULSE, SINE)
 *   zero RHS / SMPclear spicelib/analysis/cktload.c:62-65, maths/KLU/klusmp.c:485-498
 *   diagonal gmin       maths/KLU/klusmp.c:1764-1777 (LoadGmin_CSC)
 *   row scaling         maths/KLU/klu_scale.c (scale = 2: Rs[i] = max_j |A_ij|, 0 -> 1)
 *   numeric refactor    maths/KLU/klu_refactor.c:285-426 (scaled branch)
 *   triangular solves   maths/KLU/klu_solve.c (nrhs = 1), klu.c:201-250 (lsolve), :304-345 (usolve)
 *   SMPsolve wrapper    maths/KLU/klusmp.c:950-1011 (node-collapsing map, zeroing of the RHS)
 *   node convergence    maths/ni/niconv.c:41-77
 *
 * The LU is expressed as per-entry tasks: every L/U/diagonal/off-block value is
 *     v = A[slot]/Rs[row] - sum_k V[pair_l[k]] * V[pair_u[k]]     (then / pivot for L entries)
 * with the pairs in exactly the order the left-looking KLU loop applies them, so each value
 * goes through the same sequence of roundings as on the CPU (no FMA contraction); tasks are
 * grouped in dependency levels that a group of threads executes with a barrier between levels.
 * The solves are 2n row tasks (forward y_i, backward x_i) built the same way.
 */
#ifndef NGB_KERNELS_CUH
#define NGB_KERNELS_CUH

#include "ngb_types.h"
#include "devsup.cuh"

#define NGB_SP_VOLTAGE 3


/* ------------------------------------------------------------------ capacitors */
NGB_HD int ngb_cap_thread(const NgbCapCtx *c, size_t t)
{
    const int S = c->S;
    const int inst = (int)(t / (size_t)S);
    const int s = (int)(t - (size_t)inst * S);
    if (!NGB_LDG(&c->ctl.active[s])) return NGB_OK;
    const int mode = NGB_LDG(&c->ctl.mode[s]);
    const int head = NGB_LDG(&c->ctl.head[s]);
    const double cap = NGB_LDG(&c->par[(size_t)0 * c->T + t]);
    const double m = NGB_LDG(&c->par[(size_t)1 * c->T + t]);
    double geq = 0.0, ceq = 0.0;
    int stamped = 0;
#define CST(h, k) c->state[((size_t)(((head) + (h)) % c->ctl.nhist) * 2 + (k)) * c->T + t]
    {
        const int sop = NGB_LDG(&c->ctl.stateop[s]);
        if (sop) {
            for (int k = 0; k < 2; k++) {
                if (sop & NGB_OP_COPY01) CST(1, k) = CST(0, k);
                if (sop & NGB_OP_COPY1_23) { const double v = CST(1, k); CST(2, k) = v; if (c->ctl.nhist > 3) CST(3, k) = v; }
                if (sop & NGB_OP_COPY23) { const double v = CST(2, k); CST(0, k) = v; if (c->ctl.nhist > 3) CST(3, k) = v; }
            }
        }
    }
    if (mode & (NGB_MODETRAN | NGB_MODEAC | NGB_MODETRANOP)) {
        const int cond1 = (((mode & NGB_MODEDC) && (mode & NGB_MODEINITJCT)) ||
                           ((mode & NGB_MODEUIC) && (mode & NGB_MODEINITTRAN)));
        double vcap;
        if (cond1) {
            vcap = NGB_LDG(&c->par[(size_t)2 * c->T + t]);
        } else {
            const double *xo = c->x + (size_t)NGB_LDG(&c->ctl.xsel[s]) * c->neq1 * S;
            vcap = NGB_LDG(&xo[(size_t)NGB_LDG(&c->nodes[inst]) * S + s])
                 - NGB_LDG(&xo[(size_t)NGB_LDG(&c->nodes[c->ninst + inst]) * S + s]);
        }
        if (mode & (NGB_MODETRAN | NGB_MODEAC)) {
            const int order = NGB_LDG(&c->ctl.order[s]);
            const double ag0 = NGB_LDG(&c->ctl.ag0[s]), ag1 = NGB_LDG(&c->ctl.ag1[s]);
            double q0, q1, cc;
            if (order != 1 && order != 2) return NGB_E_ORDER;
            if (mode & NGB_MODEINITPRED) {
                q0 = CST(1, 0);
                CST(0, 0) = q0;
            } else {
                q0 = cap * vcap;
                CST(0, 0) = q0;
                if (mode & NGB_MODEINITTRAN) CST(1, 0) = q0;
            }
            q1 = CST(1, 0);
            {
                const int gear = c->ctl.gear;
                cc = ngb_integrate(gear, order, ag0, ag1, gear ? NGB_LDG(&c->ctl.ag2[s]) : 0.0, q0, q1,
                                   (gear && order == 2) ? CST(2, 0) : 0.0, (order == 2) ? CST(1, 1) : 0.0);
            }
            CST(0, 1) = cc;
            ceq = cc - ag0 * q0;
            geq = ag0 * cap;
            /* CAPtrunc -> CKTterr */
            if (c->ctl.lte)
                ngb_lte_state(&c->ctl, s, c->state, 2, (size_t)c->T, t, head, 0, order);
            if (mode & NGB_MODEINITTRAN) CST(1, 1) = cc;
            stamped = 1;
        } else {
            CST(0, 0) = cap * vcap;
        }
    }
#undef CST
    {
        const double g = stamped ? m * geq : 0.0, i = stamped ? m * ceq : 0.0;
        const double v[6] = { g, g, -g, -g, -i, i };
        for (int k = 0; k < 6; k++) {
            int r = NGB_LDG(&c->spos[k * c->ninst + inst]);
            if (r >= 0) c->stamp[(size_t)r * S + s] = v[k];
        }
    }
    return NGB_OK;
}

/* ------------------------------------------------------------------ independent sources */
NGB_HD double ngb_src_value(const NgbSrcCtx *c, size_t t, int inst, int s, int mode, int poff)
{
    const int ftype = NGB_LDG(&c->fn[inst]);
    const int forder = NGB_LDG(&c->fn[c->ninst + inst]);
    const int dcGiven = NGB_LDG(&c->fn[2 * c->ninst + inst]);
    const double dc = NGB_LDG(&c->par[(size_t)0 * c->T + t]);
#define SCO(k) NGB_LDG(&c->par[(size_t)(poff + (k)) * c->T + t])
    double value, time;
    if ((mode & (NGB_MODEDCOP | NGB_MODEDCTRANCURVE)) && dcGiven) {
        value = dc * NGB_LDG(&c->ctl.srcfact[s]);
    } else {
        time = (mode & NGB_MODEDC) ? 0.0 : NGB_LDG(&c->ctl.time[s]);
        switch (ftype) {
        default:
            value = dc;
            break;
        case NGB_FN_PULSE: {
            const double V1 = SCO(0), V2 = SCO(1);
            const double TD = forder > 2 ? SCO(2) : 0.0;
            const double TR = (forder > 3 && SCO(3) > 0.0) ? SCO(3) : c->tstep;
            const double TF = (forder > 4 && SCO(4) > 0.0) ? SCO(4) : c->tstep;
            const double PW = (forder > 5 && SCO(5) >= 0.0) ? SCO(5) : 0.0;
            const double PER = (forder > 6 && SCO(6) > 0.0) ? SCO(6) : TR + TF + PW;
            const double PHASE = forder > 7 ? SCO(7) : 0.0;
            double tmax = 1e99;
            time -= TD;
            if (PHASE > 0.0) tmax = PHASE * PER;      /* 8th parameter: number of pulses */
            if (time > tmax) {
                value = V1;
            } else {
                if (time > PER) {
                    double basetime = PER * floor(time / PER);
                    time -= basetime;
                }
                if (time <= 0 || time >= TR + PW + TF) value = V1;
                else if (time >= TR && time <= TR + PW) value = V2;
                else if (time > 0 && time < TR) value = V1 + (V2 - V1) * (time) / TR;
                else value = V2 + (V1 - V2) * (time - (TR + PW)) / TF;
            }
        } break;
        case NGB_FN_SINE: {
            const double PHASE = forder > 5 ? SCO(5) : 0.0;
            const double phase = PHASE * M_PI / 180.0;
            const double VO = SCO(0), VA = SCO(1);
            const double FREQ = (forder > 2 && SCO(2) != 0.0) ? SCO(2) : (1 / c->tstop);
            const double TD = forder > 3 ? SCO(3) : 0.0;
            const double THETA = forder > 4 ? SCO(4) : 0.0;
            time -= TD;
            if (time <= 0) value = VO + VA * sin(phase);
            else value = VO + VA * sin(FREQ * time * 2.0 * M_PI + phase) * ngb_exp(-time * THETA);
        } break;
        case NGB_FN_EXP: {          /* vsrcload.c:203-231 */
            const double V1 = SCO(0), V2 = SCO(1);
            const double TD1 = (forder > 2 && SCO(2) != 0.0) ? SCO(2) : c->tstep;
            const double TAU1 = (forder > 3 && SCO(3) != 0.0) ? SCO(3) : c->tstep;
            const double TD2 = (forder > 4 && SCO(4) != 0.0) ? SCO(4) : TD1 + c->tstep;
            const double TAU2 = (forder > 5 && SCO(5) != 0.0) ? SCO(5) : c->tstep;
            if (time <= TD1) value = V1;
            else if (time <= TD2) value = V1 + (V2 - V1) * (1 - ngb_exp(-(time - TD1) / TAU1));
            else value = V1 + (V2 - V1) * (1 - ngb_exp(-(time - TD1) / TAU1)) + (V1 - V2) * (1 - ngb_exp(-(time - TD2) / TAU2));
        } break;
        case NGB_FN_SFFM: {         /* vsrcload.c:233-287 */
            if (c->is_current) {    /* isrcload.c:206-254 differs: no delay is applied (TD is read and not used), the
                                     * modulation phase is coefficient 5 and the carrier phase coefficient 6 */
                const double VO = SCO(0), VA = SCO(1);
                const double FC = forder > 2 ? SCO(2) : (5. / c->tstop);
                double MDI = forder > 3 ? SCO(3) : 90.0;
                const double FM = (forder > 4 && SCO(4) != 0.0) ? SCO(4) : (500. / c->tstop);
                const double phasem = (forder > 5 ? SCO(5) : 0.0) * M_PI / 180.0;
                const double phasec = (forder > 6 ? SCO(6) : 0.0) * M_PI / 180.0;
                if (MDI > FC / FM) MDI = FC / FM; else if (MDI < 0) MDI = 0;
                value = VO + VA * sin((2.0 * M_PI * FC * time + phasec) + MDI * sin(2.0 * M_PI * FM * time + phasem));
                break;
            }
            const double VO = SCO(0), VA = SCO(1);
            const double FC = forder > 2 ? SCO(2) : (5. / c->tstop);
            double MDI = forder > 3 ? SCO(3) : 90.0;
            const double FM = (forder > 4 && SCO(4) != 0.0) ? SCO(4) : (500. / c->tstop);
            const double TD = forder > 5 ? SCO(5) : 0;
            const double phasem = (forder > 6 ? SCO(6) : 0.0) * M_PI / 180.0;
            const double phasec = (forder > 7 ? SCO(7) : 0.0) * M_PI / 180.0;
            if (MDI > FC / FM) MDI = FC / FM; else if (MDI < 0) MDI = 0;
            time -= TD;
            if (time <= 0) value = 0;
            else value = VO + VA * sin((2.0 * M_PI * FC * time + phasec) + MDI * sin(2.0 * M_PI * FM * time + phasem));
        } break;
        case NGB_FN_AM: {           /* vsrcload.c:288-322 */
            const double VO = SCO(0), VMO = SCO(1);
            const double VMA = forder > 2 ? SCO(2) : 1.;
            const double FM = forder > 3 ? SCO(3) : (5. / c->tstop);
            const double FC = forder > 4 ? SCO(4) : (500. / c->tstop);
            const double TD = forder > 5 ? SCO(5) : 0.0;
            const double phasem = (forder > 6 ? SCO(6) : 0.0) * M_PI / 180.0;
            const double phasec = (forder > 7 ? SCO(7) : 0.0) * M_PI / 180.0;
            time -= TD;
            if (time <= 0) value = 0;
            else value = VO + (VMO + VMA * sin(2.0 * M_PI * FM * time + phasem)) * sin(2.0 * M_PI * FC * time + phasec);
        } break;
        case NGB_FN_PWL: {          /* vsrcload.c:324-367; forder = number of list entries */
            const double *co = c->pwl + NGB_LDG(&c->pwl_ptr[inst]);
            if (c->is_current) {    /* isrcload.c:291-314: the older search, no delay, no repetition */
                if (time < NGB_LDG(&co[0])) { value = NGB_LDG(&co[1]); break; }
                value = NGB_LDG(&co[forder - 1]);
                for (int i = 0; i < forder / 2 - 1; i++) {
                    const double t0 = NGB_LDG(&co[2 * i]), t1 = NGB_LDG(&co[2 * (i + 1)]);
                    if (t0 == time) { value = NGB_LDG(&co[2 * i + 1]); break; }
                    if (t0 < time && t1 > time) {
                        value = NGB_LDG(&co[2 * i + 1]) + (((time - t0) / (t1 - t0)) * (NGB_LDG(&co[2 * i + 3]) - NGB_LDG(&co[2 * i + 1])));
                        break;
                    }
                }
                break;
            }
            const int rep = NGB_LDG(&c->pwl_rep[inst]);
            time -= NGB_LDG(&c->pwl_rdelay[inst]);
            if (time <= NGB_LDG(&co[0])) { value = NGB_LDG(&co[1]); break; }
            const double end_time = NGB_LDG(&co[forder - 2]);
            if (time > end_time) {
                if (rep >= 0) {
                    const double period = end_time - NGB_LDG(&co[rep]);
                    time -= NGB_LDG(&co[rep]);
                    time -= period * floor(time / period);
                    time += NGB_LDG(&co[rep]);
                } else { value = NGB_LDG(&co[forder - 1]); break; }
            }
            value = dc;
            for (int i = 2; i < forder; i += 2) {
                const double itime = NGB_LDG(&co[i]);
                if (itime >= time) {
                    time -= NGB_LDG(&co[i - 2]);
                    time /= NGB_LDG(&co[i]) - NGB_LDG(&co[i - 2]);
                    value = NGB_LDG(&co[i - 1]);
                    value += time * (NGB_LDG(&co[i + 1]) - NGB_LDG(&co[i - 1]));
                    break;
                }
            }
        } break;
        }
    }
#undef SCO
    if (mode & NGB_MODETRANOP) value *= NGB_LDG(&c->ctl.srcfact[s]);
    return value;
}

NGB_HD int ngb_src_thread(const NgbSrcCtx *c, size_t t)
{
    const int S = c->S;
    const int inst = (int)(t / (size_t)S);
    const int s = (int)(t - (size_t)inst * S);
    if (!NGB_LDG(&c->ctl.active[s])) return NGB_OK;
    const int mode = NGB_LDG(&c->ctl.mode[s]);
    if (!c->is_current) {
        const double value = ngb_src_value(c, t, inst, s, mode, 1);
        int r = NGB_LDG(&c->spos[inst]);
        if (r >= 0) c->stamp[(size_t)r * S + s] = value;
    } else {
        const double m = NGB_LDG(&c->par[(size_t)1 * c->T + t]);
        const double value = ngb_src_value(c, t, inst, s, mode, 2);
        int r = NGB_LDG(&c->spos[inst]);
        if (r >= 0) c->stamp[(size_t)r * S + s] = m * value;           /* rhs[pos] += m*value */
        r = NGB_LDG(&c->spos[c->ninst + inst]);
        if (r >= 0) c->stamp[(size_t)r * S + s] = -(m * value);        /* rhs[neg] -= m*value */
    }
    return NGB_OK;
}

/* ------------------------------------------------------------------ assembly */
/* one thread per (target, sample): u = target * S + s */
NGB_HD void ngb_asm_store(const NgbAsmCtx *c, int tg, int s, double acc);
NGB_HD void ngb_asm_thread(const NgbAsmCtx *c, size_t u)
{
    const int S = c->S;
    const int tg = (int)(u / (size_t)S);
    const int s = (int)(u - (size_t)tg * S);
    if (!NGB_LDG(&c->ctl.active[s])) return;
    double acc = 0.0;
    const int lo = NGB_LDG(&c->tgt_ptr[tg]), hi = NGB_LDG(&c->tgt_ptr[tg + 1]);
#ifdef __CUDA_ARCH__
    if (hi - lo > NGB_ASM_LONG && c->nlong) return;       /* ngb_asm_long_group's job */
#endif
    /* a gather: four independent loads in flight, added in list order (the summation order is the contract) */
    int p = lo;
    for (; p + 4 <= hi; p += 4) {
        const int r0 = NGB_LDG(&c->tgt_rows[p]), r1 = NGB_LDG(&c->tgt_rows[p + 1]), r2 = NGB_LDG(&c->tgt_rows[p + 2]), r3 = NGB_LDG(&c->tgt_rows[p + 3]);
        const double a0 = NGB_LDG(&c->stamp[(size_t)r0 * S + s]), a1 = NGB_LDG(&c->stamp[(size_t)r1 * S + s]);
        const double a2 = NGB_LDG(&c->stamp[(size_t)r2 * S + s]), a3 = NGB_LDG(&c->stamp[(size_t)r3 * S + s]);
        acc += a0; acc += a1; acc += a2; acc += a3;
    }
    for (; p < hi; p++)
        acc += NGB_LDG(&c->stamp[(size_t)NGB_LDG(&c->tgt_rows[p]) * S + s]);
    ngb_asm_store(c, tg, s, acc);
}

NGB_HD void ngb_asm_store(const NgbAsmCtx *c, int tg, int s, double acc)
{
    const int S = c->S;
    if (tg < c->nnz) {
        /* LoadGmin_CSC: CKTdiagGmin on every present diagonal, applied with the factor call */
        if (c->add_diag_gmin && NGB_LDG(&c->slot_diag[tg])) {
            const double dg = NGB_LDG(&c->ctl.diag_gmin[s]);
            if (dg != 0.0) acc += dg;
        }
        c->Ax[(size_t)s * c->nnz + tg] = acc;
    } else {
        const int eq = tg - c->nnz;
        double *rhs = c->x + (size_t)(1 - NGB_LDG(&c->ctl.xsel[s])) * c->neq1 * S;
        rhs[(size_t)eq * S + s] = acc;
    }
}

/* nodeset / ic assignments at the end of CKTload (cktload.c:118-172), one thread per sample, rows in
 * reference order.  LoadGmin_CSC runs after CKTload in the reference; the assembly has already added
 * CKTdiagGmin to the diagonals, so a rewritten diagonal gets it again here. */
NGB_HD void ngb_override_thread(const NgbAsmCtx *c, int s)
{
    const int S = c->S;
    if (!NGB_LDG(&c->ctl.active[s])) return;
    const int mode = NGB_LDG(&c->ctl.mode[s]);
    if (!(mode & NGB_MODEDC)) return;
    const double srcfact = NGB_LDG(&c->ctl.srcfact[s]);
    const double dg = c->add_diag_gmin ? NGB_LDG(&c->ctl.diag_gmin[s]) : 0.0;
    double *Ax = c->Ax + (size_t)s * c->nnz;
    double *rhs = c->x + (size_t)(1 - NGB_LDG(&c->ctl.xsel[s])) * c->neq1 * S;
    for (int i = 0; i < c->nov; i++) {
        const int kind = NGB_LDG(&c->ov_kind[i]);
        if (kind == 0 ? !(mode & (NGB_MODEINITJCT | NGB_MODEINITFIX)) : (!(mode & NGB_MODETRANOP) || (mode & NGB_MODEUIC))) continue;
        const int eq = NGB_LDG(&c->ov_eq[i]), d = NGB_LDG(&c->ov_diag[i]);
        const double v = NGB_LDG(&c->ov_val[(size_t)i * S + s]);
        for (int p = NGB_LDG(&c->ov_zptr[i]); p < NGB_LDG(&c->ov_zptr[i + 1]); p++) Ax[NGB_LDG(&c->ov_zslot[p])] = 0.0;
        if (NGB_LDG(&c->ov_cur[i])) {
            rhs[(size_t)eq * S + s] = 1.0e10 * v * srcfact;
            if (d >= 0) { if (kind == 0) Ax[d] = 1e10; else Ax[d] += 1.0e10; }
        } else {
            rhs[(size_t)eq * S + s] = v * srcfact;
            if (d >= 0) Ax[d] = 1;
        }
        if (d >= 0 && dg != 0.0) Ax[d] += dg;
    }
}

/* ------------------------------------------------------------------ LU */
#ifndef NGB_GROUP_SYNC
#define NGB_GROUP_SYNC() ((void)0)     /* hostsim: one "thread" per group */
#endif

/* A pivoting event (SMPreorder) is due for sample s and the refactor has just run on the recorded order of that event: is
 * this order the one a pivoting factor of THIS matrix would produce?  lpivot (klu_kernel.c:370-470) takes the diagonal when
 * |x_d| >= tol * max|x_i|, else the first largest entry; on the normalised L entries l = x / pivot that reads |l| * tol <= 1
 * (code 1), |l| < 1 (code 2: another pivot won, ties go to the host) and |l_d| < tol (code 3: the diagonal lost).  Decisions
 * within 1e-9 of the threshold, infinities and NaN count as "no": the sample then reports E_SINGULAR with singular_col = -2
 * and gets the host's pivoting factor (ngb_tran.c: repivot_suspended), exactly like a zero pivot.  V: the group's value array,
 * ext: internal -> schedule numbering of the packed kernels (NULL: identity) */
NGB_HD void ngb_lu_verify_order(const NgbLuCtx *c, int s, int lane, int nl, const double *V, const int *ext)
{
    const int *chk = c->sch.vchk;
    const int nV = c->sch.nV;
    const double tol = c->pivtol, one = 1.0 - 1e-9;
    int bad = 0;
    if (!chk) bad = (lane == 0);
    else
        for (int e = lane; e < nV; e += nl) {
            const int code = NGB_LDG(&chk[ext ? NGB_LDG(&ext[e]) : e]);
            if (!code) continue;
            const double a = fabs(V[e]);
            if (code == 1 ? !(a * tol <= one) : (code == 2 ? !(a <= one) : !(a <= tol * one))) bad = 1;
        }
    if (bad) { c->singular_col[s] = -2; c->ctl.err[s] = NGB_E_SINGULAR; }
}

/* Whole SMPluFac/SMPsolve/NIconvTest sequence for sample s, executed by a group of `nl`
 * threads (`lane` = index in the group).  V, Rs, Z are group-private scratch (shared memory
 * on the device): V[nV], Rs[n], Z[ntask]. */
NGB_HD void ngb_lu_sample(const NgbLuCtx *c, int s, int lane, int nl, double *V, double *Rs, double *Z)
{
    const unsigned ngb_gsync_mask = 0xffffffffu; (void)ngb_gsync_mask;
    const NgbLuSched *h = &c->sch;
    const int S = c->S, n = h->n, nV = h->nV;
    const int active = NGB_LDG(&c->ctl.active[s]);
    if (!active) return;
    if (c->ctl.lusel && NGB_LDG(&c->ctl.lusel[s]) != c->which) return;
    const double *Ax = c->Ax + (size_t)s * h->nnz;

    if (c->do_factor) {
        /* row scale factors */
        for (int i = lane; i < n; i += nl) {
            double r = 0.0;
            const int lo = NGB_LDG(&h->row_ptr[i]), hi = NGB_LDG(&h->row_ptr[i + 1]);
            for (int p = lo; p < hi; p++) {
                double a = fabs(Ax[NGB_LDG(&h->row_slot[p])]);
                r = (r > a) ? r : a;
            }
            if (r == 0.0) r = 1.0;
            Rs[i] = r;
        }
        NGB_GROUP_SYNC();
        for (int lev = 0; lev < h->nlev; lev++) {
            const int lo = NGB_LDG(&h->lev_ptr[lev]), hi = NGB_LDG(&h->lev_ptr[lev + 1]);
            for (int q = lo + lane; q < hi; q += nl) {
                const int e = NGB_LDG(&h->lev_ent[q]);
                const int as = NGB_LDG(&h->e_aslot[e]);
                double v = 0.0;
                if (as >= 0) v = Ax[as] / Rs[NGB_LDG(&h->e_arow[e])];
                const int p0 = NGB_LDG(&h->e_pptr[e]), p1 = NGB_LDG(&h->e_pptr[e + 1]);
                for (int p = p0; p < p1; p++) {
                    /* X[i] -= Lx * ujk : product and difference rounded separately */
#ifdef __CUDA_ARCH__
                    v = __dsub_rn(v, __dmul_rn(V[NGB_LDG(&h->pair_l[p])], V[NGB_LDG(&h->pair_u[p])]));
#else
                    { volatile double pr = V[h->pair_l[p]] * V[h->pair_u[p]]; v = v - pr; }
#endif
                }
                const int dv = NGB_LDG(&h->e_div[e]);
                if (dv >= 0) v = v / V[dv];
                V[e] = v;
            }
            NGB_GROUP_SYNC();
        }
        /* zero pivot -> E_SINGULAR (klu_refactor.c:390-404) */
#ifdef __CUDA_ARCH__
        if (nl > 1024) {             /* grid-wide group: the scan over n pivots is shared out */
            if (lane == 0) c->singular_col[s] = 0x7fffffff;
            NGB_GROUP_SYNC();
            for (int k = lane; k < n; k += nl)
                if (V[NGB_LDG(&h->diag_v[k])] == 0.0) atomicMin(&c->singular_col[s], k);
            NGB_GROUP_SYNC();
            if (lane == 0) {
                const int sc = c->singular_col[s];
                c->singular_col[s] = (sc == 0x7fffffff) ? -1 : sc;
                if (sc != 0x7fffffff) c->ctl.err[s] = NGB_E_SINGULAR;
            }
        } else
#endif
        if (lane == 0) {
            int sc = -1;
            for (int k = 0; k < n; k++)
                if (V[NGB_LDG(&h->diag_v[k])] == 0.0) { sc = k; break; }
            c->singular_col[s] = sc;
            if (sc >= 0) c->ctl.err[s] = NGB_E_SINGULAR;
        }
        if (c->verify && NGB_LDG(&c->verify[s])) { NGB_GROUP_SYNC(); ngb_lu_verify_order(c, s, lane, nl, V, NULL); }
        if (c->V && c->V + (size_t)s * nV != V) {       /* (the grid-wide LU works in these arrays) */
            double *Vg = c->V + (size_t)s * nV;
            for (int e = lane; e < nV; e += nl) Vg[e] = V[e];
            double *Rg = c->Rs + (size_t)s * n;
            for (int i = lane; i < n; i += nl) Rg[i] = Rs[i];
        }
        NGB_GROUP_SYNC();
    } else if (c->V + (size_t)s * nV != V) {
        const double *Vg = c->V + (size_t)s * nV;
        for (int e = lane; e < nV; e += nl) V[e] = Vg[e];
        const double *Rg = c->Rs + (size_t)s * n;
        for (int i = lane; i < n; i += nl) Rs[i] = Rg[i];
        NGB_GROUP_SYNC();
    }

    if (c->do_solve) {
        const int xs = NGB_LDG(&c->ctl.xsel[s]);
        double *rhs = c->x + (size_t)(1 - xs) * c->neq1 * S;
        const double *old = c->x + (size_t)xs * c->neq1 * S;
        for (int lev = 0; lev < h->nslev; lev++) {
            const int lo = NGB_LDG(&h->slev_ptr[lev]), hi = NGB_LDG(&h->slev_ptr[lev + 1]);
            for (int q = lo + lane; q < hi; q += nl) {
                const int tk = NGB_LDG(&h->slev_task[q]);
                const int kind = NGB_LDG(&h->t_kind[tk]);
                const int ini = NGB_LDG(&h->t_init[tk]);
                double z;
                if (kind == 0) {
                    const int eq = NGB_LDG(&h->b_eq[ini]);
                    z = rhs[(size_t)eq * S + s] / Rs[ini];
                } else {
                    z = Z[ini];
                }
                const int p0 = NGB_LDG(&h->t_pptr[tk]), p1 = NGB_LDG(&h->t_pptr[tk + 1]);
                for (int p = p0; p < p1; p++) {
#ifdef __CUDA_ARCH__
                    z = __dsub_rn(z, __dmul_rn(V[NGB_LDG(&h->t_val[p])], Z[NGB_LDG(&h->t_src[p])]));
#else
                    { volatile double pr = V[h->t_val[p]] * Z[h->t_src[p]]; z = z - pr; }
#endif
                }
                if (kind == 1) z = z / V[NGB_LDG(&h->t_div[tk])];
                Z[tk] = z;
            }
            NGB_GROUP_SYNC();
        }
        /* SMPsolve: zero the RHS, scatter the solution back through the column permutation */
        for (int i = lane; i < c->neq1; i += nl) rhs[(size_t)i * S + s] = 0.0;
        NGB_GROUP_SYNC();
        for (int k = lane; k < n; k += nl) {
            const int eq = NGB_LDG(&h->out_eq[k]);
            if (eq != 0) rhs[(size_t)eq * S + s] = Z[NGB_LDG(&h->out_task[k])];
        }
        NGB_GROUP_SYNC();
        /* NIconvTest node loop (niconv.c:41-77): result 1 = not converged */
        if (c->nodeconv) {
            int bad = 0;
            for (int i = 1 + lane; i <= n; i += nl) {
                const double nw = rhs[(size_t)i * S + s], od = old[(size_t)i * S + s];
                if (nw != nw) { bad = 1; continue; }
                const double mx = (fabs(od) > fabs(nw)) ? fabs(od) : fabs(nw);
                const double tol = c->reltol * mx
                                 + ((NGB_LDG(&c->node_type[i]) == NGB_SP_VOLTAGE) ? c->vntol : c->abstol);
                if (fabs(nw - od) > tol) bad = 1;
            }
#ifdef __CUDA_ARCH__
            if (bad) atomicOr(&c->nodeconv[s], 1);
#else
            if (bad) c->nodeconv[s] = 1;
#endif
        }
    }
}


/* Packed variant: `sb` is the schedule blob (shared memory on the device), V/Rs/Z group-private.
 * Same arithmetic, same order; only the bookkeeping differs:
 *   1. row scale factors, 2. every value initialised to A/Rs in one parallel pass,
 *   3. levels (contiguous ranges) touch shared memory only. */
/* Each level runs in two phases: the products l*u of ALL its pairs, spread evenly over the lanes
 * (P), then per entry the subtractions in KLU's order.  The long rows (a supply node touches every
 * stage) would otherwise keep one lane busy with index loads and multiplies while 31 wait. */
NGB_HD void ngb_lu_sample_packed(const NgbLuCtx *c, const unsigned short *sb, int s, int lane, int nl,
                                 double *V, double *Rs, double *Z, double *As, double *P, unsigned ngb_gsync_mask)
{
    (void)ngb_gsync_mask;
    const NgbLuPacked *h = &c->pk;
    const int S = c->S, n = h->n, nV = h->nV;
    if (!NGB_LDG(&c->ctl.active[s])) return;
    if (c->ctl.lusel && NGB_LDG(&c->ctl.lusel[s]) != c->which) return;

    if (c->do_factor) {
        /* the sample's matrix is read where it lies (L2: the assembly kernel has just written it); keeping
         * a copy in shared memory cost a third of the samples an SM can hold */
        const double *Ax = c->Ax + (size_t)s * h->nnz;
        if (lane == 0) c->singular_col[s] = -1;
        for (int i = lane; i < n; i += nl) {
            double r = 0.0;
            const int lo = sb[h->o_rowptr + i], hi = sb[h->o_rowptr + i + 1];
            for (int p = lo; p < hi; p++) {
                double a = fabs(NGB_LDG(&Ax[sb[h->o_rowslot + p]]));
                r = (r > a) ? r : a;
            }
            if (r == 0.0) r = 1.0;
            Rs[i] = r;
        }
        NGB_GROUP_SYNC();
        for (int e = lane; e < nV; e += nl) {
            const int as = sb[h->o_aslot + e];
            V[e] = (as != 0xFFFF) ? NGB_LDG(&Ax[as]) / Rs[sb[h->o_arow + e]] : 0.0;
        }
        NGB_GROUP_SYNC();
        for (int lev = 0; lev < h->nlev; lev++) {
            const int lo = sb[h->o_lev_ptr + lev], hi = sb[h->o_lev_ptr + lev + 1];
            const int pbase = sb[h->o_pptr + lo], pend = sb[h->o_pptr + hi];
            if (pend > pbase) {
                for (int p = pbase + lane; p < pend; p += nl) {
#ifdef __CUDA_ARCH__
                    P[p - pbase] = __dmul_rn(V[sb[h->o_pl + p]], V[sb[h->o_pu + p]]);
#else
                    { volatile double pr = V[sb[h->o_pl + p]] * V[sb[h->o_pu + p]]; P[p - pbase] = pr; }
#endif
                }
                NGB_GROUP_SYNC();
            }
            for (int e = lo + lane; e < hi; e += nl) {
                double v = V[e];
                const int p0 = sb[h->o_pptr + e], p1 = sb[h->o_pptr + e + 1];
                for (int p = p0; p < p1; p++) {
#ifdef __CUDA_ARCH__
                    v = __dsub_rn(v, P[p - pbase]);
#else
                    v = v - P[p - pbase];
#endif
                }
                const int dv = sb[h->o_div + e];
                if (dv != 0xFFFF) v = v / V[dv];
                V[e] = v;
            }
            NGB_GROUP_SYNC();
        }
        for (int k = lane; k < n; k += nl)
            if (V[sb[h->o_diag + k]] == 0.0) { c->singular_col[s] = k; c->ctl.err[s] = NGB_E_SINGULAR; }
        if (c->verify && NGB_LDG(&c->verify[s])) ngb_lu_verify_order(c, s, lane, nl, V, h->ext);
        if (c->V) {
            double *Vg = c->V + (size_t)s * nV;
            for (int e = lane; e < nV; e += nl) Vg[NGB_LDG(&h->ext[e])] = V[e];
            double *Rg = c->Rs + (size_t)s * n;
            for (int i = lane; i < n; i += nl) Rg[i] = Rs[i];
        }
    } else {
        const double *Vg = c->V + (size_t)s * nV;
        for (int e = lane; e < nV; e += nl) V[e] = Vg[NGB_LDG(&h->ext[e])];
        const double *Rg = c->Rs + (size_t)s * n;
        for (int i = lane; i < n; i += nl) Rs[i] = Rg[i];
        NGB_GROUP_SYNC();
    }

    if (c->do_solve) {
        const int xs = NGB_LDG(&c->ctl.xsel[s]);
        double *rhs = c->x + (size_t)(1 - xs) * c->neq1 * S;
        const double *old = c->x + (size_t)xs * c->neq1 * S;
        /* y tasks start from b/Rs: fetch all right-hand sides in one parallel pass */
        for (int tk = lane; tk < h->ntask; tk += nl) {
            if (sb[h->o_kind + tk] == 0) {
                const int row = sb[h->o_init + tk];
                Z[tk] = rhs[(size_t)NGB_LDG(&h->b_eq[row]) * S + s] / Rs[row];
            }
        }
        NGB_GROUP_SYNC();
        for (int lev = 0; lev < h->nslev; lev++) {
            const int lo = sb[h->o_slev_ptr + lev], hi = sb[h->o_slev_ptr + lev + 1];
            const int pbase = sb[h->o_tpptr + lo], pend = sb[h->o_tpptr + hi];
            if (pend > pbase) {
                for (int p = pbase + lane; p < pend; p += nl) {
#ifdef __CUDA_ARCH__
                    P[p - pbase] = __dmul_rn(V[sb[h->o_tval + p]], Z[sb[h->o_tsrc + p]]);
#else
                    { volatile double pr = V[sb[h->o_tval + p]] * Z[sb[h->o_tsrc + p]]; P[p - pbase] = pr; }
#endif
                }
                NGB_GROUP_SYNC();
            }
            for (int tk = lo + lane; tk < hi; tk += nl) {
                const int kind = sb[h->o_kind + tk];
                double z = (kind == 0) ? Z[tk] : Z[sb[h->o_init + tk]];
                const int p0 = sb[h->o_tpptr + tk], p1 = sb[h->o_tpptr + tk + 1];
                for (int p = p0; p < p1; p++) {
#ifdef __CUDA_ARCH__
                    z = __dsub_rn(z, P[p - pbase]);
#else
                    z = z - P[p - pbase];
#endif
                }
                if (kind == 1) z = z / V[sb[h->o_tdiv + tk]];
                Z[tk] = z;
            }
            NGB_GROUP_SYNC();
        }
        for (int i = lane; i < c->neq1; i += nl) rhs[(size_t)i * S + s] = 0.0;
        NGB_GROUP_SYNC();
        for (int k = lane; k < n; k += nl) {
            const int eq = NGB_LDG(&h->out_eq[k]);
            if (eq != 0) rhs[(size_t)eq * S + s] = Z[sb[h->o_out + k]];
        }
        NGB_GROUP_SYNC();
        if (c->nodeconv) {
            int bad = 0;
            for (int i = 1 + lane; i <= n; i += nl) {
                const double nw = rhs[(size_t)i * S + s], od = old[(size_t)i * S + s];
                if (nw != nw) { bad = 1; continue; }
                const double mx = (fabs(od) > fabs(nw)) ? fabs(od) : fabs(nw);
                const double tol = c->reltol * mx
                                 + ((NGB_LDG(&c->node_type[i]) == NGB_SP_VOLTAGE) ? c->vntol : c->abstol);
                if (fabs(nw - od) > tol) bad = 1;
            }
#ifdef __CUDA_ARCH__
            if (bad) atomicOr(&c->nodeconv[s], 1);
#else
            if (bad) c->nodeconv[s] = 1;
#endif
        }
    }
}

/* ---- second packing: records (NgbLuPacked o2_*) ------------------------------------------------
 * Same arithmetic, same order of operations per value as ngb_lu_sample_packed (and therefore as
 * klu_refactor / klu_solve); what changes is how a warp gets to its operands:
 *   - the sample's matrix is read once, slot-parallel and coalesced, straight into the value array;
 *     row maxima and the division by them then work on shared memory only (the first packing walked
 *     the CSR rows with one dependent global load per entry);
 *   - a level, a value and a solve task are one 8-byte record each and a product is one 4-byte word,
 *     so every step of a level costs one index load instead of two to four dependent ones;
 *   - levels whose entries need no arithmetic (level 0 of the factorisation and of the solve) are
 *     not visited;
 *   - the solution is written, zero-filled and tested for convergence in one pass over the equations. */
#ifdef __CUDACC__
#define NGB_UNROLL _Pragma("unroll")
#define NGB_UNROLL4 _Pragma("unroll 4")
#else
#define NGB_UNROLL
#define NGB_UNROLL4
#endif
#ifdef __CUDA_ARCH__
#define NGB_DMUL(a, b) __dmul_rn((a), (b))
#define NGB_DSUB(a, b) __dsub_rn((a), (b))
#else
static inline double ngb_host_dmul(double a, double b) { volatile double r = a * b; return r; }
#define NGB_DMUL(a, b) ngb_host_dmul((a), (b))
#define NGB_DSUB(a, b) ((a) - (b))
#endif
typedef struct { unsigned x, y; } NgbRec;     /* four 16-bit fields */

NGB_HD void ngb_lu_sample_pk2(const NgbLuCtx *c, const unsigned short *sb, int s, int lane, int nl,
                              double *V, double *Rs, double *Z, double *P, unsigned ngb_gsync_mask)
{
    (void)ngb_gsync_mask;
    const NgbLuPacked *h = &c->pk;
    const int S = c->S, n = h->n, nV = h->nV;
    const NgbRec *levd = (const NgbRec *)(sb + h->o2_levd);
    const NgbRec *emeta = (const NgbRec *)(sb + h->o2_emeta);
    const unsigned *pair = (const unsigned *)(sb + h->o2_pair);
    const unsigned short *diag = sb + h->o2_diag;
    if (!NGB_LDG(&c->ctl.active[s])) return;
    if (c->ctl.lusel && NGB_LDG(&c->ctl.lusel[s]) != c->which) return;

    if (c->do_factor) {
        const double *Ax = c->Ax + (size_t)s * h->nnz;
        const unsigned *slotmap = (const unsigned *)(sb + h->o2_slotmap);
        const unsigned short *rowptr = sb + h->o2_rowptr, *rowv = sb + h->o2_rowv;
        const int nnz = h->nnz;
        if (lane == 0) c->singular_col[s] = -1;
        for (int e = lane; e < nV; e += nl) V[e] = 0.0;
        NGB_GROUP_SYNC();
        /* A -> V, four independent coalesced loads in flight per lane */
        for (int j0 = lane; j0 < nnz; j0 += 4 * nl) {
            double a[4];
NGB_UNROLL
            for (int u = 0; u < 4; u++) { const int j = j0 + u * nl; a[u] = (j < nnz) ? NGB_LDG(&Ax[j]) : 0.0; }
NGB_UNROLL
            for (int u = 0; u < 4; u++) { const int j = j0 + u * nl; if (j < nnz) V[slotmap[j] & 0xFFFFu] = a[u]; }
        }
        NGB_GROUP_SYNC();
        /* KLU_scale: largest magnitude of every row (1 for an empty row) */
        for (int i = lane; i < n; i += nl) {
            double r = 0.0;
            const int lo = rowptr[i], hi = rowptr[i + 1];
            for (int q = lo; q < hi; q++) {
                const double a = fabs(V[rowv[q]]);
                r = (r > a) ? r : a;
            }
            if (r == 0.0) r = 1.0;
            Rs[i] = r;
        }
        NGB_GROUP_SYNC();
        NGB_UNROLL4
        for (int j = lane; j < nnz; j += nl) {           /* independent divisions: four in flight per lane */
            const unsigned w = slotmap[j];
            const int e = (int)(w & 0xFFFFu);
            V[e] = V[e] / Rs[w >> 16];
        }
        NGB_GROUP_SYNC();
        {
            NgbRec d = levd[h->lev0 < h->nlev ? h->lev0 : 0];
            for (int lev = h->lev0; lev < h->nlev; lev++) {
                const int lo = (int)(d.x & 0xFFFFu), hi = (int)(d.x >> 16);
                const int pbase = (int)(d.y & 0xFFFFu), pend = (int)(d.y >> 16);
                if (lev + 1 < h->nlev) d = levd[lev + 1];
                if (pend > pbase) {
                    for (int q = pbase + lane; q < pend; q += nl) {
                        const unsigned w = pair[q];
                        P[q - pbase] = NGB_DMUL(V[w & 0xFFFFu], V[w >> 16]);
                    }
                    NGB_GROUP_SYNC();
                }
                for (int it = lo + lane; it < hi; it += nl) {           /* items: own rows first, hoisted prefixes after */
                    const NgbRec m = emeta[it];
                    const double *pp = P + ((int)(m.x & 0xFFFFu) - pbase);
                    const int cnt = (int)(m.x >> 16) - (int)(m.x & 0xFFFFu);
                    const unsigned dv = m.y & 0xFFFFu;
                    const int e = (int)(m.y >> 16);
                    double v = V[e];
                    if (pend > pbase) {
NGB_UNROLL4
                        for (int k = 0; k < cnt; k++) v = NGB_DSUB(v, pp[k]);
                    } else if (cnt) {                 /* level of single products: no first phase, multiply here */
                        const unsigned w = pair[m.x & 0xFFFFu];
                        v = NGB_DSUB(v, NGB_DMUL(V[w & 0xFFFFu], V[w >> 16]));
                    }
                    if (dv != 0xFFFFu) v = v / V[dv];
                    V[e] = v;
                }
                NGB_GROUP_SYNC();
            }
        }
        for (int k = lane; k < n; k += nl)
            if (V[diag[k]] == 0.0) { c->singular_col[s] = k; c->ctl.err[s] = NGB_E_SINGULAR; }
        if (c->verify && NGB_LDG(&c->verify[s])) ngb_lu_verify_order(c, s, lane, nl, V, h->ext);
        if (c->V) {
            double *Vg = c->V + (size_t)s * nV;
            for (int e = lane; e < nV; e += nl) Vg[NGB_LDG(&h->ext[e])] = V[e];
            double *Rg = c->Rs + (size_t)s * n;
            for (int i = lane; i < n; i += nl) Rg[i] = Rs[i];
        }
    } else {
        const double *Vg = c->V + (size_t)s * nV;
        for (int e = lane; e < nV; e += nl) V[e] = Vg[NGB_LDG(&h->ext[e])];
        const double *Rg = c->Rs + (size_t)s * n;
        for (int i = lane; i < n; i += nl) Rs[i] = Rg[i];
        NGB_GROUP_SYNC();
    }

    if (c->do_solve) {
        const NgbRec *slevd = (const NgbRec *)(sb + h->o2_slevd);
        const NgbRec *tmeta = (const NgbRec *)(sb + h->o2_tmeta);
        const unsigned *tpair = (const unsigned *)(sb + h->o2_tpair);
        const unsigned short *ttgt = sb + h->o2_ttgt;
        const NgbRec *yinit = (const NgbRec *)(sb + h->o2_yinit);
        const unsigned short *eqtask = sb + h->o2_eqtask;
        const int xs = NGB_LDG(&c->ctl.xsel[s]);
        double *rhs = c->x + (size_t)(1 - xs) * c->neq1 * S;
        const double *old = c->x + (size_t)xs * c->neq1 * S;
        /* forward-solve tasks start from b/Rs: all right-hand sides fetched with four loads in flight */
        for (int q0 = lane; q0 < n; q0 += 4 * nl) {
            double bq[4];
NGB_UNROLL
            for (int u = 0; u < 4; u++) {
                const int q = q0 + u * nl;
                bq[u] = (q < n) ? rhs[(size_t)(yinit[q].y & 0xFFFFu) * S + s] : 0.0;
            }
NGB_UNROLL
            for (int u = 0; u < 4; u++) {
                const int q = q0 + u * nl;
                if (q < n) { const NgbRec y = yinit[q]; Z[y.x & 0xFFFFu] = bq[u] / Rs[y.x >> 16]; }
            }
        }
        NGB_GROUP_SYNC();
        {
            NgbRec d = slevd[h->slev0 < h->nslev ? h->slev0 : 0];
            for (int lev = h->slev0; lev < h->nslev; lev++) {
                const int lo = (int)(d.x & 0xFFFFu), hi = (int)(d.x >> 16);
                const int pbase = (int)(d.y & 0xFFFFu), pend = (int)(d.y >> 16);
                if (lev + 1 < h->nslev) d = slevd[lev + 1];
                if (pend > pbase) {
                    for (int q = pbase + lane; q < pend; q += nl) {
                        const unsigned w = tpair[q];
                        P[q - pbase] = NGB_DMUL(V[w & 0xFFFFu], Z[w >> 16]);
                    }
                    NGB_GROUP_SYNC();
                }
                for (int it = lo + lane; it < hi; it += nl) {
                    const NgbRec m = tmeta[it];
                    const int tk = ttgt[it];
                    const double *pp = P + ((int)(m.x & 0xFFFFu) - pbase);
                    const int cnt = (int)(m.x >> 16) - (int)(m.x & 0xFFFFu);
                    const unsigned dv = m.y >> 16;
                    double z = Z[m.y & 0xFFFFu];
                    if (pend > pbase) {
NGB_UNROLL4
                        for (int k = 0; k < cnt; k++) z = NGB_DSUB(z, pp[k]);
                    } else if (cnt) {
                        const unsigned w = tpair[m.x & 0xFFFFu];
                        z = NGB_DSUB(z, NGB_DMUL(V[w & 0xFFFFu], Z[w >> 16]));
                    }
                    if (dv != 0xFFFFu) z = z / V[dv];
                    Z[tk] = z;
                }
                NGB_GROUP_SYNC();
            }
        }
        /* solution out (equations without a column get 0, row 0 is ground) + node test of NIconvTest */
        {
            int bad = 0;
            const int neq1 = c->neq1;
            for (int i0 = lane; i0 < neq1; i0 += 4 * nl) {
                double od[4];
                if (c->nodeconv) {
NGB_UNROLL
                    for (int u = 0; u < 4; u++) { const int i = i0 + u * nl; od[u] = (i < neq1 && i <= n) ? old[(size_t)i * S + s] : 0.0; }
                }
NGB_UNROLL
                for (int u = 0; u < 4; u++) {
                    const int i = i0 + u * nl;
                    if (i < neq1) {
                        const unsigned tk = eqtask[i];
                        const double nw = (tk != 0xFFFFu) ? Z[tk] : 0.0;
                        rhs[(size_t)i * S + s] = nw;
                        if (c->nodeconv && i >= 1 && i <= n) {
                            if (nw != nw) { bad = 1; continue; }
                            const double mx = (fabs(od[u]) > fabs(nw)) ? fabs(od[u]) : fabs(nw);
                            const double tol = c->reltol * mx
                                             + ((NGB_LDG(&c->node_type[i]) == NGB_SP_VOLTAGE) ? c->vntol : c->abstol);
                            if (fabs(nw - od[u]) > tol) bad = 1;
                        }
                    }
                }
            }
            if (c->nodeconv) {
#ifdef __CUDA_ARCH__
                if (bad) atomicOr(&c->nodeconv[s], 1);
#else
                if (bad) c->nodeconv[s] = 1;
#endif
            }
        }
    }
}

#endif
