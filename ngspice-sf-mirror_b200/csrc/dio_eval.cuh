/* dio_eval.cuh -- junction diode load, one thread per (instance, sample).
 *
 * Follows DIOload (src/spicelib/devices/dio/dioload.c:17-865) for the configuration without
 * self-heating and soft reverse recovery (those are refused at upload, dio_fields.h): initial-voltage selection :139-221, pnjlim with the
 * breakdown mirror :296-329, bottom / sidewall / tunnel currents with high-injection knees
 * :360-522, depletion + diffusion + overlap charge :530-604, NIintegrate :667-680, convergence
 * flag :713-726, state stores :727-742 and the stamps :757-790.  DIOtrunc (diotrunc.c:22-29) is
 * folded in: the LTE bound of the junction charge is reduced into ctl.lte.
 */
#ifndef NGB_DIO_EVAL_CUH
#define NGB_DIO_EVAL_CUH
#include "ngb_types.h"
#include "dio_fields.h"
#include "devsup.cuh"

typedef struct NgbDioCtx {
    int ninst, S, T;
    const int *nodes;       /* [4][ninst] pos, neg, posPrime, posSwPrime               */
    const int *flags;       /* [ninst] DIOF_*                                          */
    const double *par;      /* [DIOP_COUNT][T]                                         */
    const int *spos;        /* [DIOS_COUNT][ninst] stamp rows, -1 = ground             */
    double *state;          /* [nhist][DIOST_COUNT][T]                                 */
    double *stamp;
    const double *x; int neq1;
    double reltol, abstol, vntol, chgtol, trtol;
    NgbCtl ctl;
} NgbDioCtx;

#define NGB_CONSTKoverQ (1.38064852e-23 / 1.6021766208e-19)   /* CONSTboltz / CHARGE (const.h:32,37; main.c:492) */
#define NGB_CONSTe 2.7182818284590452354                      /* CONSTnap (const.h) */

NGB_HD int ngb_dio_thread(const NgbDioCtx *c, size_t t)
{
    const int S = c->S;
    const int inst = (int)(t / (size_t)S);
    const int s = (int)(t - (size_t)inst * S);
    if (!NGB_LDG(&c->ctl.active[s])) return NGB_OK;
    const int mode = NGB_LDG(&c->ctl.mode[s]);
    const int head = NGB_LDG(&c->ctl.head[s]);
    const int fl = NGB_LDG(&c->flags[inst]);
    const int nh = c->ctl.nhist;
#define P(n) NGB_LDG(&c->par[(size_t)DIOP_##n * c->T + t])
#define ST(h, k) c->state[((size_t)(((head) + (h)) % nh) * DIOST_COUNT + (k)) * c->T + t]
    {   /* deferred whole-vector state copies of DCtran (dctran.c:319-322, 711-716) */
        const int sop = NGB_LDG(&c->ctl.stateop[s]);
        if (sop) {
            for (int k = 0; k < DIOST_COUNT; k++) {
                if (sop & NGB_OP_COPY01) ST(1, k) = ST(0, k);
                if (sop & NGB_OP_COPY1_23) { const double v = ST(1, k); ST(2, k) = v; if (nh > 3) ST(3, k) = v; }
                if ((sop & NGB_OP_COPY23) && nh > 3) ST(3, k) = ST(2, k);
            }
        }
    }
    const double gmin = NGB_LDG(&c->ctl.gmin[s]);
    const double Temp = P(temp);
    const double vt = NGB_CONSTKoverQ * Temp;
    const double vte = P(emissionCoeff) * vt;
    const double vtesw = P(swEmissionCoeff) * vt;
    const double vtebrk = P(brkdEmissionCoeff) * vt;
    const double vterec = P(recEmissionCoeff) * vt;
    const double gspr = P(tConductance), gsprsw = P(tConductanceSW);
    const double tBV = P(tBrkdwnV);
    const int sepsw = (fl & DIOF_RESISTSW) != 0;       /* sidewall diode behind its own series resistance */
    double vd, vdsw = 0.0, cd, gd, cdb, gdb, cdb_dT, cdsw = 0.0, gdsw = 0.0, cdsw_dT = 0.0, dIdio_dT, dIdioSw_dT = 0.0;
    double cdres, gdres;
    int Check = 1, Check_sw = 1;

    if (mode & NGB_MODEINITSMSIG) {
        vd = ST(0, DIOST_voltage);
        if (sepsw) vdsw = ST(0, DIOST_voltageSW);
    } else if (mode & NGB_MODEINITTRAN) {
        vd = ST(1, DIOST_voltage);
        if (sepsw) vdsw = ST(1, DIOST_voltageSW);
    } else if ((mode & NGB_MODEINITJCT) && (mode & NGB_MODETRANOP) && (mode & NGB_MODEUIC)) {
        vd = P(initCond);
        if (sepsw) vdsw = P(initCond);
    } else if ((mode & NGB_MODEINITJCT) && (fl & DIOF_OFF)) {
        vd = vdsw = 0.0;
    } else if (mode & NGB_MODEINITJCT) {
        vd = P(tVcrit);
        vdsw = P(tVcritSW);
    } else if ((mode & NGB_MODEINITFIX) && (fl & DIOF_OFF)) {
        vd = vdsw = 0.0;
    } else {
        if (mode & NGB_MODEINITPRED) {
            /* DEVpred (devsup.c): extrapolation from the two previous points */
            const double d0 = NGB_LDG(&c->ctl.delta[s]), d1 = NGB_LDG(&c->ctl.delta_old[(size_t)1 * S + s]);
            const double xfact = d0 / d1;
            ST(0, DIOST_voltage) = ST(1, DIOST_voltage);
            vd = (1 + xfact) * ST(1, DIOST_voltage) - xfact * ST(2, DIOST_voltage);
            ST(0, DIOST_current) = ST(1, DIOST_current);
            ST(0, DIOST_conduct) = ST(1, DIOST_conduct);
            ST(0, DIOST_deltemp) = ST(1, DIOST_deltemp);
            ST(0, DIOST_dIdio_dT) = ST(1, DIOST_dIdio_dT);
            ST(0, DIOST_qth) = ST(1, DIOST_qth);
            if (sepsw) {
                vdsw = (1 + xfact) * ST(1, DIOST_voltageSW) - xfact * ST(2, DIOST_voltageSW);
                ST(0, DIOST_dIdioSW_dT) = ST(1, DIOST_dIdioSW_dT);
            }
            ST(0, DIOST_resCurrent) = ST(1, DIOST_resCurrent);
            ST(0, DIOST_resConduct) = ST(1, DIOST_resConduct);
            ST(0, DIOST_cqcsr) = ST(1, DIOST_cqcsr);
            ST(0, DIOST_gqcsr) = ST(1, DIOST_gqcsr);
        } else {
            const double *xo = c->x + (size_t)NGB_LDG(&c->ctl.xsel[s]) * c->neq1 * S;
            const double vneg = NGB_LDG(&xo[(size_t)NGB_LDG(&c->nodes[c->ninst + inst]) * S + s]);
            vd = NGB_LDG(&xo[(size_t)NGB_LDG(&c->nodes[2 * c->ninst + inst]) * S + s]) - vneg;
            if (sepsw) vdsw = NGB_LDG(&xo[(size_t)NGB_LDG(&c->nodes[3 * c->ninst + inst]) * S + s]) - vneg;
            ST(0, DIOST_qth) = 0.0;                        /* cth0 * delTemp, no self-heating */
            if (mode & NGB_MODEINITTRAN) ST(1, DIOST_qth) = 0.0;
        }
        /* limit the new junction voltage */
        {
            const double lim = -tBV + 10 * vtebrk;
            if ((fl & DIOF_BV) && vd < NGB_MIN(0.0, lim)) {
                double vdtemp = -(vd + tBV);
                vdtemp = ngb_pnjlim(vdtemp, -(ST(0, DIOST_voltage) + tBV), vtebrk, P(tVcrit), &Check);
                vd = -(vdtemp + tBV);
            } else {
                vd = ngb_pnjlim(vd, ST(0, DIOST_voltage), vte, P(tVcrit), &Check);
            }
            if (sepsw) {
                if ((fl & DIOF_BV) && vdsw < NGB_MIN(0.0, lim)) {
                    double vdtemp = -(vdsw + tBV);
                    vdtemp = ngb_pnjlim(vdtemp, -(ST(0, DIOST_voltageSW) + tBV), vtebrk, P(tVcritSW), &Check_sw);
                    vdsw = -(vdtemp + tBV);
                } else {
                    vdsw = ngb_pnjlim(vdsw, ST(0, DIOST_voltageSW), vtesw, P(tVcritSW), &Check_sw);
                }
            }
        }
    }

    /* dc current and derivatives */
    {
        const double csat = P(tSatCur), csat_dT = P(tSatCur_dT);
        const double csatsw = P(tSatSWCur), csatsw_dT = P(tSatSWCur_dT);
        if ((fl & DIOF_SATSW) && (fl & DIOF_NSW)) {           /* sidewall with its own characteristic */
            const double vds = sepsw ? vdsw : vd;
            if (vds >= -3 * vtesw) {
                const double evd = ngb_exp(vds / vtesw);
                cdsw = csatsw * (evd - 1);
                gdsw = csatsw * evd / vtesw;
                cdsw_dT = csatsw_dT * (evd - 1) - csatsw * vds * evd / (vtesw * Temp);
            } else if (!(fl & DIOF_BV) || vds >= -tBV) {
                double argsw = 3 * vtesw / (vds * NGB_CONSTe), argsw_dT;
                argsw = argsw * argsw * argsw;
                argsw_dT = 3 * argsw / Temp;
                cdsw = -csatsw * (1 + argsw);
                gdsw = csatsw * 3 * argsw / vds;
                cdsw_dT = -csatsw_dT - (csatsw_dT * argsw + csatsw * argsw_dT);
            } else if (!sepsw) {                              /* no breakdown for a separate sidewall diode */
                const double evrev = ngb_exp(-(tBV + vds) / vtebrk);
                const double evrev_dT = (tBV + vds) * evrev / (vtebrk * Temp);
                cdsw = -csatsw * evrev;
                gdsw = csatsw * evrev / vtebrk;
                cdsw_dT = -(csatsw_dT * evrev + csatsw * evrev_dT);
            }
        }
        if (vd >= -3 * vte) {
            const double evd = ngb_exp(vd / vte);
            cdb = csat * (evd - 1);
            gdb = csat * evd / vte;
            cdb_dT = csat_dT * (evd - 1) - csat * vd * evd / (vte * Temp);
            if ((fl & DIOF_SATSW) && !(fl & DIOF_NSW)) {
                cdsw = csatsw * (evd - 1);
                gdsw = csatsw * evd / vte;
                cdsw_dT = csatsw_dT * (evd - 1) - csatsw * vd * evd / (vte * Temp);
            }
            if (fl & DIOF_RECSAT) {                           /* recombination current */
                const double isr = P(tRecSatCur), mjh = P(tGradingCoeff) / 2;
                const double evd_rec = ngb_exp(vd / vterec);
                double cdb_rec = isr * (evd_rec - 1);
                double gdb_rec = isr * evd_rec / vterec;
                const double cdb_rec_dT = P(tRecSatCur_dT) * (evd_rec - 1) - isr * vd * evd_rec / (vterec * Temp);
                const double t1 = ngb_pow((1 - vd / P(tJctPot)), 2) + 0.005;
                const double gen_fac = ngb_pow(t1, mjh);
                const double gen_fac_vd = -P(tGradingCoeff) * (1 - vd / P(tJctPot)) * ngb_pow(t1, (mjh - 1));
                cdb_rec = cdb_rec * gen_fac;
                gdb_rec = gdb_rec * gen_fac + cdb_rec * gen_fac_vd;
                cdb = cdb + cdb_rec;
                gdb = gdb + gdb_rec;
                cdb_dT = cdb_dT + cdb_rec_dT * gen_fac;
            }
        } else if (!(fl & DIOF_BV) || vd >= -tBV) {
            double arg = 3 * vte / (vd * NGB_CONSTe), darg_dT;
            arg = arg * arg * arg;
            darg_dT = 3 * arg / Temp;
            if (fl & DIOF_RECSAT) {
                const double evd_rec = ngb_exp((-3 * vte) / vterec);
                const double cdb_rec = P(tRecSatCur) * (evd_rec - 1);
                const double t1 = ngb_pow((1 - (-3 * vte) / P(tJctPot)), 2) + 0.005;
                const double gen_fac = ngb_pow(t1, P(tGradingCoeff) / 2);
                cdb = -csat * (1 + arg) + gen_fac * cdb_rec;
            } else {
                cdb = -csat * (1 + arg);
            }
            gdb = csat * 3 * arg / vd;
            cdb_dT = -csat_dT - (csat_dT * arg + csat * darg_dT);
            if ((fl & DIOF_SATSW) && !(fl & DIOF_NSW)) {
                cdsw = -csatsw * (1 + arg);
                gdsw = csatsw * 3 * arg / vd;
                cdsw_dT = -csatsw_dT - (csatsw_dT * arg + csatsw * darg_dT);
            }
        } else {
            double evrev = ngb_exp(-(tBV + vd) / vtebrk);
            double evrev_dT = (tBV + vd) * evrev / (vtebrk * Temp);
            if (fl & DIOF_RECSAT) {
                const double evd_rec = ngb_exp((-3 * vte) / vterec);
                const double cdb_rec = P(tRecSatCur) * (evd_rec - 1);
                const double t1 = ngb_pow((1 - (-3 * vte) / P(tJctPot)), 2) + 0.005;
                const double gen_fac = ngb_pow(t1, P(tGradingCoeff) / 2);
                cdb = -csat * evrev + gen_fac * cdb_rec;
            } else {
                cdb = -csat * evrev;
            }
            gdb = csat * evrev / vtebrk;
            cdb_dT = -(csat_dT * evrev + csat * evrev_dT);
            if ((fl & DIOF_SATSW) && !(fl & DIOF_NSW) && !sepsw) {
                /* vdsw is 0 without a separate sidewall diode (dioload.c:56, 459) */
                evrev = ngb_exp(-(tBV + 0.0) / vtebrk);
                evrev_dT = (tBV + 0.0) * evrev / (vtebrk * Temp);
                cdsw = -csatsw * evrev;
                gdsw = csatsw * evrev / vtebrk;
                cdsw_dT = -(csatsw_dT * evrev + csatsw * evrev_dT);
            }
        }
        if (fl & DIOF_TUNSW) {
            const double vtetun = P(tunEmissionCoeff) * vt;
            const double evd = ngb_exp(-vd / vtetun);
            const double is = P(tTunSatSWCur);
            cdsw = cdsw - is * (evd - 1);
            gdsw = gdsw + is * evd / vtetun;
            cdsw_dT = cdsw_dT - P(tTunSatSWCur_dT) * (evd - 1) - is * vd * evd / (vtetun * Temp);
        }
        if (fl & DIOF_TUN) {
            const double vtetun = P(tunEmissionCoeff) * vt;
            const double evd = ngb_exp(-vd / vtetun);
            const double is = P(tTunSatCur);
            cdb = cdb - is * (evd - 1);
            gdb = gdb + is * evd / vtetun;
            cdb_dT = cdb_dT - P(tTunSatCur_dT) * (evd - 1) - is * vd * evd / (vtetun * Temp);
        }
        if (vd >= -3 * vte) {
            if ((fl & DIOF_IKF) && cdb > 1.0e-18) {
                const double ik = P(forwardKneeCurrent);
                const double sq = sqrt(cdb / ik);
                gdb = ((1 + sq) * gdb - cdb * gdb / (2 * sq * ik)) / (1 + 2 * sq + cdb / ik);
                cdb = cdb / (1 + sq);
            }
        } else {
            if ((fl & DIOF_IKR) && cdb < -1.0e-18) {
                const double ik = P(reverseKneeCurrent);
                const double sq = sqrt(cdb / (-ik));
                gdb = ((1 + sq) * gdb + cdb * gdb / (2 * sq * ik)) / (1 + 2 * sq - cdb / ik);
                cdb = cdb / (1 + sq);
            }
        }
        if ((fl & DIOF_IKP) && cdsw > 1.0e-18) {
            const double ik = P(forwardSWKneeCurrent);
            const double sq = sqrt(cdsw / ik);
            gdsw = ((1 + sq) * gdsw - cdsw * gdsw / (2 * sq * ik)) / (1 + 2 * sq + cdsw / ik);
            cdsw = cdsw / (1 + sq);
        }
        if (!sepsw) {
            cd = cdb + cdsw + gmin * vd;
            gd = gdb + gdsw + gmin;
            dIdio_dT = cdb_dT + cdsw_dT;
        } else {
            cd = cdb + gmin * vd;
            gd = gdb + gmin;
            cdsw = cdsw + gmin * vdsw;
            gdsw = gdsw + gmin;
            dIdio_dT = cdb_dT;
            dIdioSw_dT = cdsw_dT;
        }
    }
    cdres = cd; gdres = gd;

    if ((mode & (NGB_MODEDCTRANCURVE | NGB_MODETRAN | NGB_MODEAC | NGB_MODEINITSMSIG)) ||
        ((mode & NGB_MODETRANOP) && (mode & NGB_MODEUIC))) {
        /* charge storage */
        const double czero = P(tJctCap), mj = P(tGradingCoeff), pb = P(tJctPot), fcpb = P(tDepCap);
        const double czeroSW = P(tJctSWCap), mjsw = P(gradingSWCoeff), pbsw = P(tJctSWPot), fcpbsw = P(tDepSWCap);
        const double cov = P(cmetal) + P(cpoly);
        const double tt = P(tTransitTime);
        const double vdx = sepsw ? vdsw : vd;
        double deplcharge, deplcap, deplchargeSW, deplcapSW, capd, capdsw = 0.0;
        if (vd < fcpb) {
            const double arg = 1 - vd / pb;
            const double sarg = ngb_exp(-mj * ngb_log(arg));
            deplcharge = pb * czero * (1 - arg * sarg) / (1 - mj);
            deplcap = czero * sarg;
        } else {
            const double czof2 = czero / P(tF2);
            deplcharge = czero * P(tF1) + czof2 * (P(tF3) * (vd - fcpb) + (mj / (pb + pb)) * (vd * vd - fcpb * fcpb));
            deplcap = czof2 * (P(tF3) + mj * vd / pb);
        }
        if (vdx < fcpbsw) {
            const double argSW = 1 - vdx / pbsw;
            const double sargSW = ngb_exp(-mjsw * ngb_log(argSW));
            deplchargeSW = pbsw * czeroSW * (1 - argSW * sargSW) / (1 - mjsw);
            deplcapSW = czeroSW * sargSW;
        } else {
            const double czof2SW = czeroSW / P(tF2SW);
            deplchargeSW = czeroSW * P(tF1) + czof2SW * (P(tF3SW) * (vdx - fcpbsw) + (mjsw / (pbsw + pbsw)) * (vdx * vdx - fcpbsw * fcpbsw));
            deplcapSW = czof2SW * (P(tF3SW) + mjsw * vdx / pbsw);
        }
        {
            const double diffcharge = tt * cd, diffcap = tt * gd;
            if (!sepsw) {
                ST(0, DIOST_capCharge) = diffcharge + deplcharge + deplchargeSW + cov * vd;
                capd = diffcap + deplcap + deplcapSW + P(cmetal) + P(cpoly);
            } else {
                ST(0, DIOST_capCharge) = diffcharge + deplcharge + cov * vd;
                capd = diffcap + deplcap + P(cmetal) + P(cpoly);
                ST(0, DIOST_capChargeSW) = deplcapSW;          /* sic: dioload.c:596-597 stores the capacitance */
                capdsw = deplcapSW;
            }
            ST(0, DIOST_srcapCharge) = 0.0;
        }
        if (!(mode & NGB_MODETRANOP) || !(mode & NGB_MODEUIC)) {
            if (mode & NGB_MODEINITSMSIG) {
                ST(0, DIOST_capCurrent) = capd;
                if (sepsw) ST(0, DIOST_capCurrentSW) = capdsw;
                return NGB_OK;                             /* `continue` of dioload.c:628 */
            }
            {
                const int order = NGB_LDG(&c->ctl.order[s]);
                const double ag0 = NGB_LDG(&c->ctl.ag0[s]), ag1 = NGB_LDG(&c->ctl.ag1[s]);
                double q0, q1, cc, geq;
                const int gear = c->ctl.gear;
                const double ag2 = gear ? NGB_LDG(&c->ctl.ag2[s]) : 0.0;
                if (order != 1 && order != 2) return NGB_E_ORDER;
                if (mode & NGB_MODEINITTRAN) {
                    ST(1, DIOST_capCharge) = ST(0, DIOST_capCharge);
                    if (sepsw) ST(1, DIOST_capChargeSW) = ST(0, DIOST_capChargeSW);
                }
                q0 = ST(0, DIOST_capCharge); q1 = ST(1, DIOST_capCharge);
                cc = ngb_integrate(gear, order, ag0, ag1, ag2, q0, q1, (gear && order == 2) ? ST(2, DIOST_capCharge) : 0.0, (order == 2) ? ST(1, DIOST_capCurrent) : 0.0);
                ST(0, DIOST_capCurrent) = cc;
                geq = ag0 * capd;
                gd = gd + geq;
                cd = cd + cc;
                if (sepsw) {
                    const double qs0 = ST(0, DIOST_capChargeSW), qs1 = ST(1, DIOST_capChargeSW);
                    const double ccs = ngb_integrate(gear, order, ag0, ag1, ag2, qs0, qs1, (gear && order == 2) ? ST(2, DIOST_capChargeSW) : 0.0, (order == 2) ? ST(1, DIOST_capCurrentSW) : 0.0);
                    ST(0, DIOST_capCurrentSW) = ccs;
                    gdsw = gdsw + ag0 * capdsw;
                    cdsw = cdsw + ccs;
                    if (mode & NGB_MODEINITTRAN) ST(1, DIOST_capCurrentSW) = ccs;
                    if (c->ctl.lte)
                        ngb_lte_state(&c->ctl, s, c->state, DIOST_COUNT, (size_t)c->T, t, head, DIOST_capChargeSW, order);
                }
                if (mode & NGB_MODEINITTRAN) ST(1, DIOST_capCurrent) = cc;
                if (c->ctl.lte)                            /* DIOtrunc -> CKTterr on the junction charge */
                    ngb_lte_state(&c->ctl, s, c->state, DIOST_COUNT, (size_t)c->T, t, head, DIOST_capCharge, order);
            }
        }
    }

    /* convergence flag */
    if (!(mode & NGB_MODEINITFIX) || !(fl & DIOF_OFF)) {
        if (Check == 1 || (sepsw && Check_sw == 1)) {
#ifdef __CUDA_ARCH__
            atomicAdd(&c->ctl.noncon[s], 1);
#else
            c->ctl.noncon[s] += 1;
#endif
        }
    }
    ST(0, DIOST_voltage) = vd;
    ST(0, DIOST_current) = cd;
    ST(0, DIOST_conduct) = gd;
    ST(0, DIOST_deltemp) = 0.0;
    ST(0, DIOST_dIdio_dT) = dIdio_dT;
    if (sepsw) {
        ST(0, DIOST_voltageSW) = vdsw;
        ST(0, DIOST_currentSW) = cdsw;
        ST(0, DIOST_conductSW) = gdsw;
        ST(0, DIOST_dIdioSW_dT) = dIdioSw_dT;
    }
    ST(0, DIOST_qp) = 0.0;                                 /* rhsOld[qpNode = 0] */
    ST(0, DIOST_resCurrent) = cdres;
    ST(0, DIOST_resConduct) = gdres;
    ST(0, DIOST_cqcsr) = 0.0;
    ST(0, DIOST_gqcsr) = 0.0;

    /* stamps, in the statement order of dioload.c:757-790 */
    {
        const double cdeq = cd - gd * vd;
#define STAMP(k, v) do { const int r_ = NGB_LDG(&c->spos[(k) * c->ninst + inst]); if (r_ >= 0) c->stamp[(size_t)r_ * S + s] = (v); } while (0)
        STAMP(DIOS_rhsNeg, cdeq);
        STAMP(DIOS_rhsPosPrime, -cdeq);
        STAMP(DIOS_posPos, gspr);
        STAMP(DIOS_negNeg, gd);
        STAMP(DIOS_ppPp, gd + gspr);
        STAMP(DIOS_posPp, -gspr);
        STAMP(DIOS_negPp, -gd);
        STAMP(DIOS_ppPos, -gspr);
        STAMP(DIOS_ppNeg, -gd);
        if (sepsw) {
            const double cdeqsw = cdsw - gdsw * vdsw;
            STAMP(DIOS_rhsNegSw, cdeqsw);
            STAMP(DIOS_rhsPosSwPrime, -cdeqsw);
            STAMP(DIOS_posPosSw, gsprsw);
            STAMP(DIOS_negNegSw, gdsw);
            STAMP(DIOS_pspPsp, gdsw + gsprsw);
            STAMP(DIOS_posPsp, -gsprsw);
            STAMP(DIOS_negPsp, -gdsw);
            STAMP(DIOS_pspPos, -gsprsw);
            STAMP(DIOS_pspNeg, -gdsw);
        }
#undef STAMP
    }
#undef P
#undef ST
    return NGB_OK;
}
#endif
