/* dio_eval.cuh -- junction diode load, one thread per (instance, sample).
 *
 * Follows DIOload (src/spicelib/devices/dio/dioload.c:17-865): initial-voltage selection :139-221, pnjlim with the
 * breakdown mirror :296-329, bottom / sidewall / tunnel currents with high-injection knees
 * :360-522, depletion + diffusion + overlap charge :530-604, NIintegrate :667-680, convergence
 * flag :713-726, state stores :727-742 and the stamps :757-790.  DIOtrunc (diotrunc.c:22-29) is
 * folded in: the LTE bound of the junction charge is reduced into ctl.lte.
 * Self-heating (thermal node `dt`, rth0 / cth0): the temperature rise is limited with DEVlimitlog, the parameter set is
 * mapped to DIOtemp + delTemp by ngb_dio_temp_update -- DIOtempUpdate (diotemp.c:18-270) -- inside every load and written
 * back to the thread's parameter column like the reference writes its instance structure (the NEXT load limits with this
 * load's tVcrit / tBrkdwnV), the dissipated power and the d/dT terms are stamped (:744-835).  Soft reverse recovery (vp, tt):
 * the charge node qp and its subcircuit (:572-586, 677-690, 838-862).
 */
#ifndef NGB_DIO_EVAL_CUH
#define NGB_DIO_EVAL_CUH
#include "ngb_types.h"
#include "dio_fields.h"
#include "devsup.cuh"

typedef struct NgbDioCtx {
    int ninst, S, T;
    const int *nodes;       /* [6][ninst] pos, neg, posPrime, posSwPrime, temp, qp     */
    const int *flags;       /* [ninst] DIOF_*                                          */
    double *par;            /* [DIOP_COUNT][T]; instances with self-heating rewrite their temperature-mapped entries */
    const int *spos;        /* [DIOS_COUNT][ninst] stamp rows, -1 = ground             */
    double *state;          /* [nhist][DIOST_COUNT][T]                                 */
    double *stamp;
    const double *x; int neq1;
    double reltol, abstol, vntol, chgtol, trtol;
    NgbCtl ctl;
} NgbDioCtx;

#define NGB_CONSTKoverQ (1.38064852e-23 / 1.6021766208e-19)   /* CONSTboltz / CHARGE (const.h:32,37; main.c:492) */
#define NGB_CONSTe 2.7182818284590452354                      /* CONSTnap (const.h) */

/* DIOtempUpdate (diotemp.c:18-270): the temperature-dependent parameters of one instance at temperature Temp, with the
 * d/dT the reference carries (saturation currents and series conductances only); q is the thread's parameter column */
NGB_HD void ngb_dio_temp_update(double *q, int fl, double Temp, double reltol)
{
#define Q(n) q[DIOP_##n]
    const double kb = 1.38064852e-23, chg = 1.6021766208e-19, reftemp = 27.0 + 273.15, root2 = 1.4142135623730950488016887242097;
    double vt, vte, vts, vtt, vtr, vtnom, dt, factor, lnTRatio, egfet, egfet1, egfet_dT = 0.0, fact2, arg, pbfact, arg1, fact1, pbfact1;
    double vte_dT, vts_dT, vtt_dT, vtr_dT, arg0, arg1_dT, arg2, arg2_dT, xfc, xfcs, totalSatCur;
    const double nom = Q(nomTemp), eg = Q(activationEnergy), xti = Q(saturationCurrentExp);
    const int tlev = (int)Q(tlev), tlevc = (int)Q(tlevc);
    vt = NGB_CONSTKoverQ * Temp;
    vte = Q(emissionCoeff) * vt;       vte_dT = NGB_CONSTKoverQ * Q(emissionCoeff);
    vts = Q(swEmissionCoeff) * vt;     vts_dT = NGB_CONSTKoverQ * Q(swEmissionCoeff);
    vtt = Q(tunEmissionCoeff) * vt;    vtt_dT = NGB_CONSTKoverQ * Q(tunEmissionCoeff);
    vtr = Q(recEmissionCoeff) * vt;    vtr_dT = NGB_CONSTKoverQ * Q(recEmissionCoeff);
    vtnom = NGB_CONSTKoverQ * nom;
    dt = Temp - nom;
    lnTRatio = ngb_log(Temp / nom);
    factor = 1.0 + (Q(gradCoeffTemp1) * dt) + (Q(gradCoeffTemp2) * dt * dt);
    Q(tGradingCoeff) = Q(gradingCoeff) * factor;
    if ((tlev == 0) || (tlev == 1)) {
        egfet = 1.16 - (7.02e-4 * Temp * Temp) / (Temp + 1108);
        egfet1 = 1.16 - (7.02e-4 * nom * nom) / (nom + 1108);
    } else {
        egfet = eg - (Q(firstBGcorrFactor) * Temp * Temp) / (Temp + Q(secndBGcorrFactor));
        egfet_dT = (Q(firstBGcorrFactor) * Temp * Temp) / ((Temp + Q(secndBGcorrFactor)) * (Temp + Q(secndBGcorrFactor)))
                   - 2 * Q(firstBGcorrFactor) * Temp / (Temp + Q(secndBGcorrFactor));
        egfet1 = eg - (Q(firstBGcorrFactor) * nom * nom) / (nom + Q(secndBGcorrFactor));
    }
    fact2 = Temp / reftemp;
    arg = -egfet / (2 * kb * Temp) + 1.1150877 / (kb * (reftemp + reftemp));
    pbfact = -2 * vt * (1.5 * ngb_log(fact2) + chg * arg);
    arg1 = -egfet1 / (kb * 2 * nom) + 1.1150877 / (2 * kb * reftemp);
    fact1 = nom / reftemp;
    pbfact1 = -2 * vtnom * (1.5 * ngb_log(fact1) + chg * arg1);
    if (tlevc == 0) {
        const double pbo = (Q(junctionPot) - pbfact1) / fact1;
        const double gmaold = (Q(junctionPot) - pbo) / pbo;
        double gmanew;
        Q(tJctCap) = Q(junctionCap) / (1 + Q(tGradingCoeff) * (400e-6 * (nom - reftemp) - gmaold));
        Q(tJctPot) = pbfact + fact2 * pbo;
        gmanew = (Q(tJctPot) - pbo) / pbo;
        Q(tJctCap) *= 1 + Q(tGradingCoeff) * (400e-6 * (Temp - reftemp) - gmanew);
    } else if (tlevc == 1) {
        Q(tJctPot) = Q(junctionPot) - Q(tpb) * (Temp - reftemp);
        Q(tJctCap) = Q(junctionCap) * (1 + Q(cta) * (Temp - reftemp));
    }
    if (tlevc == 0) {
        const double pboSW = (Q(junctionSWPot) - pbfact1) / fact1;
        const double gmaSWold = (Q(junctionSWPot) - pboSW) / pboSW;
        double gmaSWnew;
        Q(tJctSWCap) = Q(junctionSWCap) / (1 + Q(gradingSWCoeff) * (400e-6 * (nom - reftemp) - gmaSWold));
        Q(tJctSWPot) = pbfact + fact2 * pboSW;
        gmaSWnew = (Q(tJctSWPot) - pboSW) / pboSW;
        Q(tJctSWCap) *= 1 + Q(gradingSWCoeff) * (400e-6 * (Temp - reftemp) - gmaSWnew);
    } else if (tlevc == 1) {
        Q(tJctSWPot) = Q(junctionSWPot) - Q(tphp) * (Temp - reftemp);
        Q(tJctSWCap) = Q(junctionSWCap) * (1 + Q(ctp) * (Temp - reftemp));
    }
    /* saturation currents: is(T) = is * scale * exp(argument), d/dT = is(T) * d(argument)/dT */
    if ((tlev == 0) || (tlev == 1)) {
        arg1 = ((Temp / nom) - 1) * eg / vte;
        arg1_dT = eg / (vte * nom) - eg * (Temp / nom - 1) / (vte * Temp);
        arg2 = xti / Q(emissionCoeff) * lnTRatio; arg2_dT = xti / Q(emissionCoeff) / Temp;
        Q(tSatCur) = Q(satCur) * Q(area) * Q(m) * ngb_exp(arg1 + arg2);
        Q(tSatCur_dT) = Q(tSatCur) * (arg1_dT + arg2_dT);
        arg1 = ((Temp / nom) - 1) * eg / vts;
        arg1_dT = eg / (vts * nom) - eg * (Temp / nom - 1) / (vts * Temp);
        arg2 = xti / Q(swEmissionCoeff) * lnTRatio; arg2_dT = xti / Q(swEmissionCoeff) / Temp;
        Q(tSatSWCur) = Q(satSWCur) * Q(pj) * Q(m) * ngb_exp(arg1 + arg2);
        Q(tSatSWCur_dT) = Q(tSatSWCur) * (arg1_dT + arg2_dT);
        arg1 = ((Temp / nom) - 1) * Q(tunEGcorrectionFactor) * eg / vtt;
        arg1_dT = Q(tunEGcorrectionFactor) * eg / (vtt * nom) - eg * (Temp / nom - 1) / (vtt * Temp);
        arg2 = Q(tunSaturationCurrentExp) / Q(tunEmissionCoeff) * lnTRatio; arg2_dT = Q(tunSaturationCurrentExp) / Q(tunEmissionCoeff) / Temp;
        Q(tTunSatCur) = Q(tunSatCur) * Q(area) * Q(m) * ngb_exp(arg1 + arg2);
        Q(tTunSatCur_dT) = Q(tTunSatCur) * (arg1_dT + arg2_dT);
        Q(tTunSatSWCur) = Q(tunSatSWCur) * Q(pj) * Q(m) * ngb_exp(arg1 + arg2);
        Q(tTunSatSWCur_dT) = Q(tTunSatSWCur) * (arg1_dT + arg2_dT);
        arg1 = ((Temp / nom) - 1) * eg / vtr;
        arg1_dT = eg / (vtr * nom) - eg * (Temp / nom - 1) / (vtr * Temp);
        arg2 = xti / Q(recEmissionCoeff) * lnTRatio; arg2_dT = xti / Q(recEmissionCoeff) / Temp;
        Q(tRecSatCur) = Q(recSatCur) * Q(area) * Q(m) * ngb_exp(arg1 + arg2);
        Q(tRecSatCur_dT) = Q(tRecSatCur) * (arg1_dT + arg2_dT);
    } else {
        arg0 = egfet1 / (Q(emissionCoeff) * vtnom);
        arg1 = egfet / vte; arg1_dT = (egfet_dT * vte - egfet * vte_dT) / (egfet * egfet);
        arg2 = xti / Q(emissionCoeff) * lnTRatio; arg2_dT = xti / Q(emissionCoeff) / Temp;
        Q(tSatCur) = Q(satCur) * Q(area) * Q(m) * ngb_exp(arg0 - arg1 + arg2);
        Q(tSatCur_dT) = Q(tSatCur) * (-arg1_dT + arg2_dT);
        arg0 = egfet1 / (Q(swEmissionCoeff) * vtnom);
        arg1 = egfet / vts; arg1_dT = (egfet_dT * vts - egfet * vts_dT) / (egfet * egfet);
        arg2 = xti / Q(swEmissionCoeff) * lnTRatio; arg2_dT = xti / Q(swEmissionCoeff) / Temp;
        Q(tSatSWCur) = Q(satSWCur) * Q(pj) * Q(m) * ngb_exp(arg0 - arg1 + arg2);
        Q(tSatSWCur_dT) = Q(tSatSWCur) * (-arg1_dT + arg2_dT);
        arg0 = Q(tunEGcorrectionFactor) * egfet1 / (Q(tunEmissionCoeff) * vtnom);
        arg1 = Q(tunEGcorrectionFactor) * egfet / vtt; arg1_dT = Q(tunEGcorrectionFactor) * (egfet_dT * vtt - egfet * vtt_dT) / (egfet * egfet);
        arg2 = Q(tunSaturationCurrentExp) / Q(tunEmissionCoeff) * lnTRatio; arg2_dT = Q(tunSaturationCurrentExp) / Q(tunEmissionCoeff) / Temp;
        Q(tTunSatCur) = Q(tunSatCur) * Q(area) * Q(m) * ngb_exp(arg0 - arg1 + arg2);
        Q(tTunSatCur_dT) = Q(tTunSatCur) * (-arg1_dT + arg2_dT);
        Q(tTunSatSWCur) = Q(tunSatSWCur) * Q(pj) * Q(m) * ngb_exp(arg0 - arg1 + arg2);
        Q(tTunSatSWCur_dT) = Q(tTunSatSWCur) * (-arg1_dT + arg2_dT);
        arg0 = egfet1 / (Q(recEmissionCoeff) * vtnom);
        arg1 = egfet / vtr; arg1_dT = (egfet_dT * vtr - egfet * vtr_dT) / (egfet * egfet);
        arg2 = xti / Q(recEmissionCoeff) * lnTRatio; arg2_dT = xti / Q(recEmissionCoeff) / Temp;
        Q(tRecSatCur) = Q(recSatCur) * Q(area) * Q(m) * ngb_exp(arg0 - arg1 + arg2);
        Q(tRecSatCur_dT) = Q(tRecSatCur) * (-arg1_dT + arg2_dT);
    }
    xfc = ngb_log(1 - Q(depletionCapCoeff));
    xfcs = ngb_log(1 - Q(depletionSWcapCoeff));
    Q(tF1) = Q(tJctPot) * (1 - ngb_exp((1 - Q(tGradingCoeff)) * xfc)) / (1 - Q(tGradingCoeff));
    Q(tDepCap) = Q(depletionCapCoeff) * Q(tJctPot);
    Q(tDepSWCap) = Q(depletionSWcapCoeff) * Q(tJctSWPot);
    totalSatCur = Q(tSatCur) + Q(tSatSWCur);
    if (fl & DIOF_RESISTSW) {
        Q(tVcrit) = vte * ngb_log(vte / (root2 * Q(tSatCur)));
        Q(tVcritSW) = vts * ngb_log(vts / (root2 * Q(tSatSWCur)));
    } else {
        Q(tVcrit) = vte * ngb_log(vte / (root2 * totalSatCur));
        Q(tVcritSW) = vts * ngb_log(vts / (root2 * Q(tSatSWCur)));
    }
    if (fl & DIOF_BV) {          /* breakdown voltage matched to the saturation current at this temperature */
        double tBV, cbv, xbv, xcbv, tol;
        int iter;
        if (tlev == 0) tBV = Q(breakdownVoltage) - Q(tcv) * dt;
        else tBV = Q(breakdownVoltage) * (1 - Q(tcv) * dt);
        if ((int)Q(level) == 1) cbv = Q(m) * Q(breakdownCurrent);
        else cbv = Q(breakdownCurrent) * Q(area) * Q(m);
        if (cbv < totalSatCur * tBV / vt) {
            xbv = tBV;
        } else {
            tol = reltol * cbv;
            xbv = tBV - Q(brkdEmissionCoeff) * vt * ngb_log(1 + cbv / totalSatCur);
            for (iter = 0; iter < 25; iter++) {
                xbv = tBV - Q(brkdEmissionCoeff) * vt * ngb_log(cbv / totalSatCur + 1 - xbv / vt);
                xcbv = totalSatCur * (ngb_exp((tBV - xbv) / (Q(brkdEmissionCoeff) * vt)) - 1 + xbv / vt);
                if (fabs(xcbv - cbv) <= tol) break;
            }
        }
        Q(tBrkdwnV) = xbv;
    }
    factor = 1.0 + (Q(tranTimeTemp1) * dt) + (Q(tranTimeTemp2) * dt * dt);
    Q(tTransitTime) = Q(transitTime) * factor;
    Q(tConductance) = Q(conductance) * Q(area) * Q(m);
    if ((fl & DIOF_RESIST) && Q(resist) != 0.0) {
        factor = 1.0 + (Q(resistTemp1)) * dt + (Q(resistTemp2) * dt * dt);
        Q(tConductance) = Q(conductance) * Q(area) * Q(m) / factor;
        Q(tConductance_dT) = -Q(conductance) * Q(area) * Q(m) * (Q(resistTemp1) + Q(resistTemp2) * dt) / (factor * factor);
    }
    Q(tConductanceSW) = Q(conductanceSW) * Q(pj) * Q(m);
    if ((fl & DIOF_RESISTSW) && Q(resistSW) != 0.0) {
        factor = 1.0 + (Q(resistTemp1)) * dt + (Q(resistTemp2) * dt * dt);
        Q(tConductanceSW) = Q(conductanceSW) * Q(pj) * Q(m) / factor;
        Q(tConductanceSW_dT) = -Q(conductanceSW) * Q(pj) * Q(m) * (Q(resistTemp1) + Q(resistTemp2) * dt) / (factor * factor);
    }
    Q(tF2) = ngb_exp((1 + Q(tGradingCoeff)) * xfc);
    Q(tF3) = 1 - Q(depletionCapCoeff) * (1 + Q(tGradingCoeff));
    Q(tF2SW) = ngb_exp((1 + Q(gradingSWCoeff)) * xfcs);
    Q(tF3SW) = 1 - Q(depletionSWcapCoeff) * (1 + Q(gradingSWCoeff));
#undef Q
}

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
    /* the thread's parameter column in local storage: an instance with self-heating re-maps it to its own temperature below */
    double q[DIOP_COUNT];
    for (int k = 0; k < DIOP_COUNT_V1; k++) q[k] = c->par[(size_t)k * c->T + t];
    if (fl & DIOF_SELFHEAT)          /* the raw rows DIOtempUpdate starts from */
        for (int k = DIOP_COUNT_V1; k < DIOP_COUNT; k++) q[k] = c->par[(size_t)k * c->T + t];
    else { q[DIOP_rth0] = 0.0; q[DIOP_cth0] = 0.0; q[DIOP_softRevRecParam] = c->par[(size_t)DIOP_softRevRecParam * c->T + t]; }
#define P(n) q[DIOP_##n]
#define ST(h, k) c->state[((size_t)(((head) + (h)) % nh) * DIOST_COUNT + (k)) * c->T + t]
    {   /* deferred whole-vector state copies of DCtran (dctran.c:319-322, 711-716) */
        const int sop = NGB_LDG(&c->ctl.stateop[s]);
        if (sop) {
            for (int k = 0; k < DIOST_COUNT; k++) {
                if (sop & NGB_OP_COPY01) ST(1, k) = ST(0, k);
                if (sop & NGB_OP_COPY1_23) { const double v = ST(1, k); ST(2, k) = v; if (nh > 3) ST(3, k) = v; }
                if (sop & NGB_OP_COPY23) { const double v = ST(2, k); ST(0, k) = v; if (nh > 3) ST(3, k) = v; }
            }
        }
    }
    const double gmin = NGB_LDG(&c->ctl.gmin[s]);
    const int selfheat = (fl & DIOF_SELFHEAT) != 0, revrec = (fl & DIOF_REVREC) != 0;
    double Temp = P(temp);
    double vt = NGB_CONSTKoverQ * Temp;
    double vte = P(emissionCoeff) * vt;
    const double vtesw = P(swEmissionCoeff) * vt;      /* not re-evaluated at the raised temperature (dioload.c:320-322 updates vt, vte, vtebrk) */
    double vtebrk = P(brkdEmissionCoeff) * vt;
    const double vterec = P(recEmissionCoeff) * vt;
    double gspr = P(tConductance), gsprsw = P(tConductanceSW);
    double tBV = P(tBrkdwnV);                          /* with self-heating: the value the PREVIOUS load left (limiting uses it) */
    const int sepsw = (fl & DIOF_RESISTSW) != 0;       /* sidewall diode behind its own series resistance */
    double delTemp = 0.0, vqp = 0.0, cqcsr = 0.0, gqcsr = 0.0, gcTt = 0.0, ceqqth = 0.0;
    int Check_th = selfheat ? 1 : 0;
    const double *xo_ = c->x + (size_t)NGB_LDG(&c->ctl.xsel[s]) * c->neq1 * S;
#define XNODE(role) NGB_LDG(&xo_[(size_t)NGB_LDG(&c->nodes[(role) * c->ninst + inst]) * S + s])
    double vd, vdsw = 0.0, cd, gd, cdb, gdb, cdb_dT, cdsw = 0.0, gdsw = 0.0, cdsw_dT = 0.0, dIdio_dT, dIdioSw_dT = 0.0;
    double cdres, gdres;
    int Check = 1, Check_sw = 1;

    if (mode & NGB_MODEINITSMSIG) {
        vd = ST(0, DIOST_voltage);
        if (sepsw) vdsw = ST(0, DIOST_voltageSW);
        delTemp = ST(0, DIOST_deltemp); vqp = ST(0, DIOST_qp);
    } else if (mode & NGB_MODEINITTRAN) {
        vd = ST(1, DIOST_voltage);
        if (sepsw) vdsw = ST(1, DIOST_voltageSW);
        delTemp = ST(1, DIOST_deltemp); vqp = ST(1, DIOST_qp);
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
            delTemp = (1 + xfact) * ST(1, DIOST_deltemp) - xfact * ST(2, DIOST_deltemp);
            vqp = (1 + xfact) * ST(1, DIOST_qp) - xfact * ST(2, DIOST_qp);
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
            delTemp = selfheat ? XNODE(4) : 0.0;
            ST(0, DIOST_qth) = P(cth0) * delTemp;
            if (mode & NGB_MODEINITTRAN) ST(1, DIOST_qth) = ST(0, DIOST_qth);
            vqp = XNODE(5);                                /* rhsOld[qpNode]; node 0 (ground) without soft recovery */
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
        if (selfheat) {              /* DEVlimitlog (devsup.c:157-184) */
            const double told = ST(0, DIOST_deltemp);
            Check_th = 0;
            if (delTemp != delTemp || told != told) { delTemp = 0.0; Check_th = 1; }
            if (delTemp > told + 100.0) { delTemp = told + 100.0 + log10((delTemp - told) / 100.0); Check_th = 1; }
            else if (delTemp < told - 100.0) { delTemp = told - 100.0 - log10((told - delTemp) / 100.0); Check_th = 1; }
        } else delTemp = 0.0;
    }
    if (selfheat) {
        /* the parameter set at the raised temperature, kept for the next load like the reference's instance structure */
        Temp = P(temp) + delTemp;
        ngb_dio_temp_update(q, fl, Temp, c->reltol);
        vt = NGB_CONSTKoverQ * Temp;
        vte = P(emissionCoeff) * vt;
        vtebrk = P(brkdEmissionCoeff) * vt;
        gspr = P(tConductance); gsprsw = P(tConductanceSW); tBV = P(tBrkdwnV);
        for (int k = 0; k < DIOP_COUNT_V1; k++) c->par[(size_t)k * c->T + t] = q[k];
        c->par[(size_t)DIOP_tConductance_dT * c->T + t] = q[DIOP_tConductance_dT];
        c->par[(size_t)DIOP_tConductanceSW_dT * c->T + t] = q[DIOP_tConductanceSW_dT];
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
        if (revrec) {
            /* soft recovery: the diffusion charge lives on the qp node's capacitor (dioload.c:565-582) */
            ST(0, DIOST_capCharge) = deplcharge + deplchargeSW + cov * vd;
            capd = deplcap + deplcapSW + P(cmetal) + P(cpoly);
            ST(0, DIOST_srcapCharge) = tt * vqp;
        } else {
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
                if (revrec) {
                    if (mode & NGB_MODEINITTRAN) ST(1, DIOST_srcapCharge) = ST(0, DIOST_srcapCharge);
                    cqcsr = ngb_integrate(gear, order, ag0, ag1, ag2, ST(0, DIOST_srcapCharge), ST(1, DIOST_srcapCharge),
                                          (gear && order == 2) ? ST(2, DIOST_srcapCharge) : 0.0, (order == 2) ? ST(1, DIOST_srcapCurrent) : 0.0);
                    ST(0, DIOST_srcapCurrent) = cqcsr;
                    gqcsr = ag0 * tt;
                    if (mode & NGB_MODEINITTRAN) ST(1, DIOST_srcapCurrent) = cqcsr;
                    if (c->ctl.lte)                        /* DIOtrunc's third CKTterr (diotrunc.c:25-26) */
                        ngb_lte_state(&c->ctl, s, c->state, DIOST_COUNT, (size_t)c->T, t, head, DIOST_srcapCharge, order);
                }
                if (selfheat) {
                    const double cth = P(cth0);
                    const double cq = ngb_integrate(gear, order, ag0, ag1, ag2, ST(0, DIOST_qth), ST(1, DIOST_qth),
                                                    (gear && order == 2) ? ST(2, DIOST_qth) : 0.0, (order == 2) ? ST(1, DIOST_cqth) : 0.0);
                    ST(0, DIOST_cqth) = cq;
                    gcTt = ag0 * cth;
                    ceqqth = cq - ag0 * ST(0, DIOST_qth);     /* NIintegrate's ceq: ccap - ag0 * q0 (niinteg.c:77) */
                    if (mode & NGB_MODEINITTRAN) ST(1, DIOST_cqth) = cq;
                }
            }
        }
    }

    /* convergence flag */
    if (!(mode & NGB_MODEINITFIX) || !(fl & DIOF_OFF)) {
        if (Check_th == 1 || Check == 1 || (sepsw && Check_sw == 1)) {
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
    ST(0, DIOST_deltemp) = delTemp;
    ST(0, DIOST_dIdio_dT) = dIdio_dT;
    if (sepsw) {
        ST(0, DIOST_voltageSW) = vdsw;
        ST(0, DIOST_currentSW) = cdsw;
        ST(0, DIOST_conductSW) = gdsw;
        ST(0, DIOST_dIdioSW_dT) = dIdioSw_dT;
    }
    ST(0, DIOST_qp) = vqp;
    ST(0, DIOST_resCurrent) = cdres;
    ST(0, DIOST_resConduct) = gdres;
    ST(0, DIOST_cqcsr) = cqcsr;
    ST(0, DIOST_gqcsr) = gqcsr;

    /* stamps, in the statement order of dioload.c:757-790 */
    {
        const double cdeq = cd - gd * vd;
        double vrs = 0.0, vrssw = 0.0, Ith = 0.0, dIrs_dT = 0.0, dIth_dVrs = 0.0, dIth_dT = 0.0, dIth_dVdio = 0.0;
        double dIrssw_dT = 0.0, dIth_dVrssw = 0.0, dIth_dVdioSw = 0.0;
        if (selfheat) {              /* dissipated power and its derivatives (dioload.c:736-776) */
            vrs = XNODE(0) - XNODE(2);
            dIrs_dT = vrs * q[DIOP_tConductance_dT];
            Ith = vd * cd + vrs * vrs * gspr;
            dIth_dVrs = vrs * gspr;
            dIth_dVrs = dIth_dVrs + vrs * gspr;
            dIth_dT = vrs * dIrs_dT + dIdio_dT * vd;
            dIth_dVdio = cd + vd * gd;
            if (sepsw) {
                vrssw = XNODE(0) - XNODE(3);
                dIrssw_dT = vrssw * q[DIOP_tConductanceSW_dT];
                Ith = Ith + vdsw * cdsw + vrssw * vrssw * gsprsw;
                dIth_dVrssw = vrssw * gsprsw;
                dIth_dVrssw = dIth_dVrssw + vrssw * gsprsw;
                dIth_dT = dIth_dT + vrssw * dIrssw_dT + dIdioSw_dT * vdsw;
                dIth_dVdioSw = cdsw + vdsw * gdsw;
            }
        }
#define STAMP(k, v) do { const int r_ = NGB_LDG(&c->spos[(k) * c->ninst + inst]); if (r_ >= 0) c->stamp[(size_t)r_ * S + s] = (v); } while (0)
        STAMP(DIOS_rhsNeg, cdeq);
        STAMP(DIOS_rhsPosPrime, -cdeq);
        if (selfheat) {
            STAMP(DIOS_thRhsPos, dIrs_dT * delTemp);
            STAMP(DIOS_thRhsPp, dIdio_dT * delTemp - dIrs_dT * delTemp);
            STAMP(DIOS_thRhsNeg, -dIdio_dT * delTemp);
            STAMP(DIOS_thRhsTemp, Ith - dIth_dVdio * vd - dIth_dVrs * vrs - dIth_dT * delTemp - ceqqth);
        }
        if (sepsw) {
            const double cdeqsw = cdsw - gdsw * vdsw;
            STAMP(DIOS_rhsNegSw, cdeqsw);
            STAMP(DIOS_rhsPosSwPrime, -cdeqsw);
            if (selfheat) {
                STAMP(DIOS_thRhsPosSw, dIrssw_dT * delTemp);
                STAMP(DIOS_thRhsPsp, dIdioSw_dT * delTemp - dIrssw_dT * delTemp);
                STAMP(DIOS_thRhsNegSw, -dIdioSw_dT * delTemp);
                STAMP(DIOS_thRhsTempSw, -dIth_dVdioSw * vdsw - dIth_dVrssw * vrssw);
            }
        }
        STAMP(DIOS_posPos, gspr);
        STAMP(DIOS_negNeg, gd);
        STAMP(DIOS_ppPp, gd + gspr);
        STAMP(DIOS_posPp, -gspr);
        STAMP(DIOS_negPp, -gd);
        STAMP(DIOS_ppPos, -gspr);
        STAMP(DIOS_ppNeg, -gd);
        if (selfheat) {
            STAMP(DIOS_thTempPos, -dIth_dVrs);
            STAMP(DIOS_thTempPp, -dIth_dVdio + dIth_dVrs);
            STAMP(DIOS_thTempNeg, dIth_dVdio);
            STAMP(DIOS_thTempTemp, -dIth_dT + 1 / P(rth0) + gcTt);
            STAMP(DIOS_thPosTemp, dIrs_dT);
            STAMP(DIOS_thPpTemp, dIdio_dT - dIrs_dT);
            STAMP(DIOS_thNegTemp, -dIdio_dT);
        }
        if (sepsw) {
            STAMP(DIOS_posPosSw, gsprsw);
            STAMP(DIOS_negNegSw, gdsw);
            STAMP(DIOS_pspPsp, gdsw + gsprsw);
            STAMP(DIOS_posPsp, -gsprsw);
            STAMP(DIOS_negPsp, -gdsw);
            STAMP(DIOS_pspPos, -gsprsw);
            STAMP(DIOS_pspNeg, -gdsw);
            if (selfheat) {
                STAMP(DIOS_thTempPosSw, -dIth_dVrssw);
                STAMP(DIOS_thTempPsp, -dIth_dVdioSw + dIth_dVrssw);
                STAMP(DIOS_thTempNegSw, dIth_dVdioSw);
                STAMP(DIOS_thPosTempSw, dIrssw_dT);
                STAMP(DIOS_thPspTemp, dIdioSw_dT - dIrssw_dT);
                STAMP(DIOS_thNegTempSw, -dIdioSw_dT);
            }
        }
        if (revrec) {                /* qp node: ddt(Qp) + Qp/vp = tt/vp * Id, and its share of the diode current (dioload.c:836-862) */
            const double vp = P(softRevRecParam);
            const double fac = P(tTransitTime) / vp;
            const double dcrrdvd = fac * gdres;
            const double ceqrr = -fac * cdres + cqcsr + dcrrdvd * vd - gqcsr * vqp;
            const double grr = 1 / vp;
            const double qpGain = (1 - vp) / P(tTransitTime);
            const double geqrrd = qpGain * gqcsr;
            const double ceqrrd = qpGain * cqcsr - geqrrd * vqp;
            STAMP(DIOS_rrRhsQp, -ceqrr);
            STAMP(DIOS_rrQpQp, grr + gqcsr);
            STAMP(DIOS_rrQpPp, -dcrrdvd);
            STAMP(DIOS_rrQpNeg, dcrrdvd);
            STAMP(DIOS_rrRhsPp, -ceqrrd);
            STAMP(DIOS_rrRhsNeg, ceqrrd);
            STAMP(DIOS_rrPpQp, geqrrd);
            STAMP(DIOS_rrNegQp, -geqrrd);
        }
#undef STAMP
    }
#undef P
#undef ST
#undef XNODE
    return NGB_OK;
}
#endif
