/* bsim3_eval.cuh -- BSIM3v3.3.0 load, one thread per (instance, sample).
 *
 * Restates BSIM3load (src/spicelib/devices/bsim3/b3ld.c:42-3131, serial flavour) for the
 * configuration without NQS and with acmMod = 0 (other settings are refused at upload with
 * E_UNSUPP): initial voltages and limiting :176-378, junction diodes :401-494,
 * threshold / mobility / Vdsat / output resistance / drain and substrate current :497-1229,
 * intrinsic charges capMod 0 :1267-1600, capMod 1 :1770-1945, capMod 2 :1795-1947 and capMod 3
 * (charge-thickness model) :1950-2238,
 * junction charges :2256-2431, overlap charges and capacitance matrix :2475-2790, integration
 * and equivalent currents :2793-2900, stamps :2903-3069.  BSIM3trunc (b3trunc.c:38-40) is folded
 * in: the LTE bounds of qb, qg, qd are reduced into ctl.lte.
 */
#ifndef NGB_BSIM3_EVAL_CUH
#define NGB_BSIM3_EVAL_CUH
#include "ngb_types.h"
#include "bsim3_fields.h"
#include "devsup.cuh"

#define B3_MAX_EXP 5.834617425e14
#define B3_MIN_EXP 1.713908431e-15
#define B3_EXPT 34.0
#define B3_EPSSI 1.03594e-10
#define B3_Q 1.60219e-19
#define B3_DELTA 0.02                 /* DELTA_1 .. DELTA_4 are all 0.02 */

typedef struct B3Ctx {
    int ninst, S, T;
    const double *mtab, *ptab;       /* [nrows][B3M_COUNT], [nrows][B3P_COUNT]        */
    const int *prow;                 /* [ninst]                                         */
    const double *inst;              /* [B3I_COUNT][T]                                  */
    const int *flags;                /* [ninst]                                         */
    const int *nodes;                /* [B3N_COUNT][ninst]                              */
    const int *spos;                 /* [B3S_COUNT][ninst] stamp rows, -1 = ground      */
    double *stamp;
    double *state;                   /* [nhist][B3ST_COUNT][T]                          */
    double *von;                     /* [T] previous-iterate von (DEVfetlim)            */
    const double *x; int neq1;
    NgbCtl ctl;
    double temp, vt0;                /* CKTtemp, CONSTvt0                               */
} B3Ctx;

/* values handed from the current evaluation to the charge model */
typedef struct B3W {
    double Vds, Vgs, Vbs;
    double Vbseff, dVbseff_dVb, Phis, dPhis_dVb, sqrtPhis, dsqrtPhis_dVb;
    double Vth, dVth_dVb, dVth_dVd, Vgs_eff, dVgs_eff_dVg, Vgst;
    double n, dn_dVb, dn_dVd, Abulk0, dAbulk0_dVb, Vtm;
    double cdrain, gm, gds, gmbs, gbbs, gbgs, gbds, csub;
    double qgate, qbulk, qdrn;
    double cggb, cgsb, cgdb, cdgb, cdsb, cddb, cbgb, cbsb, cbdb;
} B3W;

#define B3M(x) NGB_LDG(&mrow[B3M_##x])
#define B3P(x) NGB_LDG(&prw[B3P_##x])
#define B3I(x) NGB_LDG(&c->inst[(size_t)B3I_##x * c->T + t])

/* bulk junction current of one side: b3ld.c:437-494 */
NGB_HD void b3_junction_dc(double isat, double v, double Nvtm, double ijth, double vjm, double IsEvjm, double gmin,
                           double *g, double *cur)
{
    if (isat <= 0.0) {
        *g = gmin;
        *cur = *g * v;
    } else if (ijth == 0.0 || v < vjm) {
        const double ev = ngb_exp(v / Nvtm);
        *g = isat * ev / Nvtm + gmin;
        *cur = isat * (ev - 1.0) + gmin * v;
    } else {
        const double T0 = IsEvjm / Nvtm;
        *g = T0 + gmin;
        *cur = IsEvjm - isat + T0 * (v - vjm) + gmin * v;
    }
}

/* drain current, substrate current and their derivatives: b3ld.c:497-1229 */
NGB_HD void b3_core_dc(const B3Ctx *c, const double *mrow, const double *prw, size_t t, B3W *w, double *von_out)
{
    const double Vds = w->Vds, Vgs = w->Vgs, Vbs = w->Vbs;
    const double Leff = B3P(leff), Vtm = B3M(vtm), cox = B3M(cox), tox = B3M(tox);
    const double phi = B3P(phi), sqrtPhi = B3P(sqrtPhi), k1ox = B3P(k1ox);
    double T0, T1, T2, T3, T4, T5, T6, T7, T8, T9, T10, tmp1, tmp2, tmp3, tmp4;
    double dT0_dVg, dT0_dVd, dT0_dVb, dT1_dVg, dT1_dVd, dT1_dVb, dT2_dVg, dT2_dVd, dT2_dVb, dT3_dVg, dT3_dVd, dT3_dVb;
    double Vbseff, dVbseff_dVb, Phis, dPhis_dVb, sqrtPhis, dsqrtPhis_dVb, Xdep, dXdep_dVb;
    double lt1, dlt1_dVb, ltw, dltw_dVb, Theta0, dTheta0_dVb, Delt_vth, dDelt_vth_dVb, V0;
    double Vth, dVth_dVb, dVth_dVd, n, dn_dVb, dn_dVd, Vgs_eff, dVgs_eff_dVg, Vgst;
    double Vgsteff, dVgsteff_dVg, dVgsteff_dVd, dVgsteff_dVb;
    double Weff, dWeff_dVg, dWeff_dVb, Rds, dRds_dVg, dRds_dVb;
    double Abulk0, dAbulk0_dVb, Abulk, dAbulk_dVb, dAbulk_dVg;
    double Denomi, dDenomi_dVg, dDenomi_dVd, dDenomi_dVb, ueff, dueff_dVg, dueff_dVd, dueff_dVb;
    double WVCox, WVCoxRds, Esat, EsatL, dEsatL_dVg, dEsatL_dVd, dEsatL_dVb, Lambda, dLambda_dVg, Vgst2Vtm;
    double Vdsat, dVdsat_dVg, dVdsat_dVd, dVdsat_dVb, Vdseff, dVdseff_dVg, dVdseff_dVd, dVdseff_dVb;
    double Vasat, dVasat_dVg, dVasat_dVd, dVasat_dVb, diffVds;
    double VACLM, dVACLM_dVg, dVACLM_dVd, dVACLM_dVb, VADIBL, dVADIBL_dVg, dVADIBL_dVd, dVADIBL_dVb;
    double Va, dVa_dVg, dVa_dVd, dVa_dVb, VASCBE, dVASCBE_dVg, dVASCBE_dVd, dVASCBE_dVb;
    double CoxWovL, beta, dbeta_dVg, dbeta_dVd, dbeta_dVb, fgche1, dfgche1_dVg, dfgche1_dVd, dfgche1_dVb;
    double fgche2, dfgche2_dVg, dfgche2_dVd, dfgche2_dVb, gche, dgche_dVg, dgche_dVd, dgche_dVb;
    double Idl, dIdl_dVg, dIdl_dVd, dIdl_dVb, Idsa, dIdsa_dVg, dIdsa_dVd, dIdsa_dVb, Ids, Gm, Gds, Gmb;
    double Isub, Gbd, Gbg, Gbb, TempRatio, DIBL_Sft, dDIBL_Sft_dVd;

    /* effective bulk bias */
    T0 = Vbs - B3P(vbsc) - 0.001;
    T1 = sqrt(T0 * T0 - 0.004 * B3P(vbsc));
    Vbseff = B3P(vbsc) + 0.5 * (T0 + T1);
    dVbseff_dVb = 0.5 * (1.0 + T0 / T1);
    if (Vbseff < Vbs) Vbseff = Vbs;

    if (Vbseff > 0.0) {
        T0 = phi / (phi + Vbseff);
        Phis = phi * T0;
        dPhis_dVb = -T0 * T0;
        sqrtPhis = B3P(phis3) / (phi + 0.5 * Vbseff);
        dsqrtPhis_dVb = -0.5 * sqrtPhis * sqrtPhis / B3P(phis3);
    } else {
        Phis = phi - Vbseff;
        dPhis_dVb = -1.0;
        sqrtPhis = sqrt(Phis);
        dsqrtPhis_dVb = -0.5 / sqrtPhis;
    }
    Xdep = B3P(Xdep0) * sqrtPhis / sqrtPhi;
    dXdep_dVb = (B3P(Xdep0) / sqrtPhi) * dsqrtPhis_dVb;

    /* threshold voltage */
    T3 = sqrt(Xdep);
    V0 = B3P(vbi) - phi;

    T0 = B3P(dvt2) * Vbseff;
    if (T0 >= -0.5) { T1 = 1.0 + T0; T2 = B3P(dvt2); }
    else { T4 = 1.0 / (3.0 + 8.0 * T0); T1 = (1.0 + 3.0 * T0) * T4; T2 = B3P(dvt2) * T4 * T4; }
    lt1 = B3M(factor1) * T3 * T1;
    dlt1_dVb = B3M(factor1) * (0.5 / T3 * T1 * dXdep_dVb + T3 * T2);

    T0 = B3P(dvt2w) * Vbseff;
    if (T0 >= -0.5) { T1 = 1.0 + T0; T2 = B3P(dvt2w); }
    else { T4 = 1.0 / (3.0 + 8.0 * T0); T1 = (1.0 + 3.0 * T0) * T4; T2 = B3P(dvt2w) * T4 * T4; }
    ltw = B3M(factor1) * T3 * T1;
    dltw_dVb = B3M(factor1) * (0.5 / T3 * T1 * dXdep_dVb + T3 * T2);

    T0 = -0.5 * B3P(dvt1) * Leff / lt1;
    if (T0 > -B3_EXPT) {
        T1 = ngb_exp(T0);
        Theta0 = T1 * (1.0 + 2.0 * T1);
        dT1_dVb = -T0 / lt1 * T1 * dlt1_dVb;
        dTheta0_dVb = (1.0 + 4.0 * T1) * dT1_dVb;
    } else {
        T1 = B3_MIN_EXP;
        Theta0 = T1 * (1.0 + 2.0 * T1);
        dTheta0_dVb = 0.0;
    }
    {
        const double thetavth = B3P(dvt0) * Theta0;
        Delt_vth = thetavth * V0;
        dDelt_vth_dVb = B3P(dvt0) * dTheta0_dVb * V0;
    }

    T0 = -0.5 * B3P(dvt1w) * B3P(weff) * Leff / ltw;
    if (T0 > -B3_EXPT) {
        T1 = ngb_exp(T0);
        T2 = T1 * (1.0 + 2.0 * T1);
        dT1_dVb = -T0 / ltw * T1 * dltw_dVb;
        dT2_dVb = (1.0 + 4.0 * T1) * dT1_dVb;
    } else {
        T1 = B3_MIN_EXP;
        T2 = T1 * (1.0 + 2.0 * T1);
        dT2_dVb = 0.0;
    }
    T0 = B3P(dvt0w) * T2;
    T2 = T0 * V0;
    dT2_dVb = B3P(dvt0w) * dT2_dVb * V0;

    TempRatio = c->temp / B3M(tnom) - 1.0;
    T0 = sqrt(1.0 + B3P(nlx) / Leff);
    T1 = k1ox * (T0 - 1.0) * sqrtPhi + (B3P(kt1) + B3P(kt1l) / Leff + B3P(kt2) * Vbseff) * TempRatio;
    tmp2 = tox * phi / (B3P(weff) + B3P(w0));

    T3 = B3P(eta0) + B3P(etab) * Vbseff;
    if (T3 < 1.0e-4) { T9 = 1.0 / (3.0 - 2.0e4 * T3); T3 = (2.0e-4 - T3) * T9; T4 = T9 * T9; }
    else T4 = 1.0;
    dDIBL_Sft_dVd = T3 * B3P(theta0vb0);
    DIBL_Sft = dDIBL_Sft_dVd * Vds;

    Vth = B3M(type) * B3I(vth0) - B3P(k1) * sqrtPhi + k1ox * sqrtPhis - B3P(k2ox) * Vbseff - Delt_vth - T2
        + (B3P(k3) + B3P(k3b) * Vbseff) * tmp2 + T1 - DIBL_Sft;
    *von_out = Vth;

    dVth_dVb = k1ox * dsqrtPhis_dVb - B3P(k2ox) - dDelt_vth_dVb - dT2_dVb + B3P(k3b) * tmp2
             - B3P(etab) * Vds * B3P(theta0vb0) * T4 + B3P(kt2) * TempRatio;
    dVth_dVd = -dDIBL_Sft_dVd;

    /* subthreshold swing factor n */
    tmp2 = B3P(nfactor) * B3_EPSSI / Xdep;
    tmp3 = B3P(cdsc) + B3P(cdscb) * Vbseff + B3P(cdscd) * Vds;
    tmp4 = (tmp2 + tmp3 * Theta0 + B3P(cit)) / cox;
    if (tmp4 >= -0.5) {
        n = 1.0 + tmp4;
        dn_dVb = (-tmp2 / Xdep * dXdep_dVb + tmp3 * dTheta0_dVb + B3P(cdscb) * Theta0) / cox;
        dn_dVd = B3P(cdscd) * Theta0 / cox;
    } else {
        T0 = 1.0 / (3.0 + 8.0 * tmp4);
        n = (1.0 + 3.0 * tmp4) * T0;
        T0 *= T0;
        dn_dVb = (-tmp2 / Xdep * dXdep_dVb + tmp3 * dTheta0_dVb + B3P(cdscb) * Theta0) / cox * T0;
        dn_dVd = B3P(cdscd) * Theta0 / cox * T0;
    }

    /* poly-gate depletion */
    T0 = B3I(vfb) + phi;
    if ((B3P(ngate) > 1.e18) && (B3P(ngate) < 1.e25) && (Vgs > T0)) {
        T1 = 1.0e6 * B3_Q * B3_EPSSI * B3P(ngate) / (cox * cox);
        T4 = sqrt(1.0 + 2.0 * (Vgs - T0) / T1);
        T2 = T1 * (T4 - 1.0);
        T3 = 0.5 * T2 * T2 / T1;
        T7 = 1.12 - T3 - 0.05;
        T6 = sqrt(T7 * T7 + 0.224);
        T5 = 1.12 - 0.5 * (T7 + T6);
        Vgs_eff = Vgs - T5;
        dVgs_eff_dVg = 1.0 - (0.5 - 0.5 / T4) * (1.0 + T7 / T6);
    } else {
        Vgs_eff = Vgs;
        dVgs_eff_dVg = 1.0;
    }
    Vgst = Vgs_eff - Vth;

    /* effective Vgst */
    T10 = 2.0 * n * Vtm;
    {
        const double VgstNVt = Vgst / T10;
        const double ExpArg = (2.0 * B3P(voff) - Vgst) / T10;
        if (VgstNVt > B3_EXPT) {
            Vgsteff = Vgst;
            dVgsteff_dVg = dVgs_eff_dVg;
            dVgsteff_dVd = -dVth_dVd;
            dVgsteff_dVb = -dVth_dVb;
        } else if (ExpArg > B3_EXPT) {
            double ExpVgst;
            T0 = (Vgst - B3P(voff)) / (n * Vtm);
            ExpVgst = ngb_exp(T0);
            Vgsteff = Vtm * B3P(cdep0) / cox * ExpVgst;
            dVgsteff_dVg = Vgsteff / (n * Vtm);
            dVgsteff_dVd = -dVgsteff_dVg * (dVth_dVd + T0 * Vtm * dn_dVd);
            dVgsteff_dVb = -dVgsteff_dVg * (dVth_dVb + T0 * Vtm * dn_dVb);
            dVgsteff_dVg *= dVgs_eff_dVg;
        } else {
            const double ExpVgst = ngb_exp(VgstNVt);
            T1 = T10 * ngb_log(1.0 + ExpVgst);
            dT1_dVg = ExpVgst / (1.0 + ExpVgst);
            dT1_dVb = -dT1_dVg * (dVth_dVb + Vgst / n * dn_dVb) + T1 / n * dn_dVb;
            dT1_dVd = -dT1_dVg * (dVth_dVd + Vgst / n * dn_dVd) + T1 / n * dn_dVd;

            dT2_dVg = -cox / (Vtm * B3P(cdep0)) * ngb_exp(ExpArg);
            T2 = 1.0 - T10 * dT2_dVg;
            dT2_dVd = -dT2_dVg * (dVth_dVd - 2.0 * Vtm * ExpArg * dn_dVd) + (T2 - 1.0) / n * dn_dVd;
            dT2_dVb = -dT2_dVg * (dVth_dVb - 2.0 * Vtm * ExpArg * dn_dVb) + (T2 - 1.0) / n * dn_dVb;

            Vgsteff = T1 / T2;
            T3 = T2 * T2;
            dVgsteff_dVg = (T2 * dT1_dVg - T1 * dT2_dVg) / T3 * dVgs_eff_dVg;
            dVgsteff_dVd = (T2 * dT1_dVd - T1 * dT2_dVd) / T3;
            dVgsteff_dVb = (T2 * dT1_dVb - T1 * dT2_dVb) / T3;
        }
    }

    /* effective channel geometry, series resistance */
    T9 = sqrtPhis - sqrtPhi;
    Weff = B3P(weff) - 2.0 * (B3P(dwg) * Vgsteff + B3P(dwb) * T9);
    dWeff_dVg = -2.0 * B3P(dwg);
    dWeff_dVb = -2.0 * B3P(dwb) * dsqrtPhis_dVb;
    if (Weff < 2.0e-8) {
        T0 = 1.0 / (6.0e-8 - 2.0 * Weff);
        Weff = 2.0e-8 * (4.0e-8 - Weff) * T0;
        T0 *= T0 * 4.0e-16;
        dWeff_dVg *= T0;
        dWeff_dVb *= T0;
    }
    T0 = B3P(prwg) * Vgsteff + B3P(prwb) * T9;
    if (T0 >= -0.9) {
        Rds = B3P(rds0) * (1.0 + T0);
        dRds_dVg = B3P(rds0) * B3P(prwg);
        dRds_dVb = B3P(rds0) * B3P(prwb) * dsqrtPhis_dVb;
    } else {
        T1 = 1.0 / (17.0 + 20.0 * T0);
        Rds = B3P(rds0) * (0.8 + T0) * T1;
        T1 *= T1;
        dRds_dVg = B3P(rds0) * B3P(prwg) * T1;
        dRds_dVb = B3P(rds0) * B3P(prwb) * dsqrtPhis_dVb * T1;
    }

    /* bulk charge effect */
    T1 = 0.5 * k1ox / sqrtPhis;
    dT1_dVb = -T1 / sqrtPhis * dsqrtPhis_dVb;
    T9 = sqrt(B3P(xj) * Xdep);
    tmp1 = Leff + 2.0 * T9;
    T5 = Leff / tmp1;
    tmp2 = B3P(a0) * T5;
    tmp3 = B3P(weff) + B3P(b1);
    tmp4 = B3P(b0) / tmp3;
    T2 = tmp2 + tmp4;
    dT2_dVb = -T9 / tmp1 / Xdep * dXdep_dVb;
    T6 = T5 * T5;
    T7 = T5 * T6;

    Abulk0 = 1.0 + T1 * T2;
    dAbulk0_dVb = T1 * tmp2 * dT2_dVb + T2 * dT1_dVb;

    T8 = B3P(ags) * B3P(a0) * T7;
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
    T2 = B3P(keta) * Vbseff;
    if (T2 >= -0.9) {
        T0 = 1.0 / (1.0 + T2);
        dT0_dVb = -B3P(keta) * T0 * T0;
    } else {
        T1 = 1.0 / (0.8 + T2);
        T0 = (17.0 + 20.0 * T2) * T1;
        dT0_dVb = -B3P(keta) * T1 * T1;
    }
    dAbulk_dVg *= T0;
    dAbulk_dVb = dAbulk_dVb * T0 + Abulk * dT0_dVb;
    dAbulk0_dVb = dAbulk0_dVb * T0 + Abulk0 * dT0_dVb;
    Abulk *= T0;
    Abulk0 *= T0;

    /* mobility */
    {
        const int mobMod = (int)B3M(mobMod);
        if (mobMod == 1) {
            T0 = Vgsteff + Vth + Vth;
            T2 = B3P(ua) + B3P(uc) * Vbseff;
            T3 = T0 / tox;
            T5 = T3 * (T2 + B3P(ub) * T3);
            dDenomi_dVg = (T2 + 2.0 * B3P(ub) * T3) / tox;
            dDenomi_dVd = dDenomi_dVg * 2.0 * dVth_dVd;
            dDenomi_dVb = dDenomi_dVg * 2.0 * dVth_dVb + B3P(uc) * T3;
        } else if (mobMod == 2) {
            T5 = Vgsteff / tox * (B3P(ua) + B3P(uc) * Vbseff + B3P(ub) * Vgsteff / tox);
            dDenomi_dVg = (B3P(ua) + B3P(uc) * Vbseff + 2.0 * B3P(ub) * Vgsteff / tox) / tox;
            dDenomi_dVd = 0.0;
            dDenomi_dVb = Vgsteff * B3P(uc) / tox;
        } else {
            T0 = Vgsteff + Vth + Vth;
            T2 = 1.0 + B3P(uc) * Vbseff;
            T3 = T0 / tox;
            T4 = T3 * (B3P(ua) + B3P(ub) * T3);
            T5 = T4 * T2;
            dDenomi_dVg = (B3P(ua) + 2.0 * B3P(ub) * T3) * T2 / tox;
            dDenomi_dVd = dDenomi_dVg * 2.0 * dVth_dVd;
            dDenomi_dVb = dDenomi_dVg * 2.0 * dVth_dVb + B3P(uc) * T4;
        }
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
    ueff = B3I(u0temp) / Denomi;
    T9 = -ueff / Denomi;
    dueff_dVg = T9 * dDenomi_dVg;
    dueff_dVd = T9 * dDenomi_dVd;
    dueff_dVb = T9 * dDenomi_dVb;

    /* saturation voltage */
    WVCox = Weff * B3P(vsattemp) * cox;
    WVCoxRds = WVCox * Rds;
    Esat = 2.0 * B3P(vsattemp) / ueff;
    EsatL = Esat * Leff;
    T0 = -EsatL / ueff;
    dEsatL_dVg = T0 * dueff_dVg;
    dEsatL_dVd = T0 * dueff_dVd;
    dEsatL_dVb = T0 * dueff_dVb;
    {
        const double a1 = B3P(a1);
        if (a1 == 0.0) {
            Lambda = B3P(a2);
            dLambda_dVg = 0.0;
        } else if (a1 > 0.0) {
            T0 = 1.0 - B3P(a2);
            T1 = T0 - B3P(a1) * Vgsteff - 0.0001;
            T2 = sqrt(T1 * T1 + 0.0004 * T0);
            Lambda = B3P(a2) + T0 - 0.5 * (T1 + T2);
            dLambda_dVg = 0.5 * B3P(a1) * (1.0 + T1 / T2);
        } else {
            T1 = B3P(a2) + B3P(a1) * Vgsteff - 0.0001;
            T2 = sqrt(T1 * T1 + 0.0004 * B3P(a2));
            Lambda = 0.5 * (T1 + T2);
            dLambda_dVg = 0.5 * B3P(a1) * (1.0 + T1 / T2);
        }
    }
    Vgst2Vtm = Vgsteff + 2.0 * Vtm;
    if (Rds > 0) {
        tmp2 = dRds_dVg / Rds + dWeff_dVg / Weff;
        tmp3 = dRds_dVb / Rds + dWeff_dVb / Weff;
    } else {
        tmp2 = dWeff_dVg / Weff;
        tmp3 = dWeff_dVb / Weff;
    }
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
        dT1_dVg = (2.0 / Lambda - 1.0) - 2.0 * Vgst2Vtm * tmp1 + Abulk * dEsatL_dVg + EsatL * dAbulk_dVg
                + 3.0 * (T9 + T7 * tmp2 + T6 * dAbulk_dVg);
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
        (void)dT3_dVg; (void)dT3_dVd; (void)dT3_dVb;
        dVdsat_dVg = (dT1_dVg - (T1 * dT1_dVg - dT0_dVg * T2 - T0 * dT2_dVg) / T3 - Vdsat * dT0_dVg) / T0;
        dVdsat_dVb = (dT1_dVb - (T1 * dT1_dVb - dT0_dVb * T2 - T0 * dT2_dVb) / T3 - Vdsat * dT0_dVb) / T0;
        dVdsat_dVd = (dT1_dVd - (T1 * dT1_dVd - T0 * dT2_dVd) / T3) / T0;
    }

    /* effective Vds */
    T1 = Vdsat - Vds - B3P(delta);
    dT1_dVg = dVdsat_dVg;
    dT1_dVd = dVdsat_dVd - 1.0;
    dT1_dVb = dVdsat_dVb;
    T2 = sqrt(T1 * T1 + 4.0 * B3P(delta) * Vdsat);
    T0 = T1 / T2;
    T3 = 2.0 * B3P(delta) / T2;
    dT2_dVg = T0 * dT1_dVg + T3 * dVdsat_dVg;
    dT2_dVd = T0 * dT1_dVd + T3 * dVdsat_dVd;
    dT2_dVb = T0 * dT1_dVb + T3 * dVdsat_dVb;
    Vdseff = Vdsat - 0.5 * (T1 + T2);
    dVdseff_dVg = dVdsat_dVg - 0.5 * (dT1_dVg + dT2_dVg);
    dVdseff_dVd = dVdsat_dVd - 0.5 * (dT1_dVd + dT2_dVd);
    dVdseff_dVb = dVdsat_dVb - 0.5 * (dT1_dVb + dT2_dVb);
    if (Vds == 0.0) { Vdseff = 0.0; dVdseff_dVg = 0.0; dVdseff_dVb = 0.0; }

    /* VAsat */
    tmp4 = 1.0 - 0.5 * Abulk * Vdsat / Vgst2Vtm;
    T9 = WVCoxRds * Vgsteff;
    T8 = T9 / Vgst2Vtm;
    T0 = EsatL + Vdsat + 2.0 * T9 * tmp4;
    T7 = 2.0 * WVCoxRds * tmp4;
    dT0_dVg = dEsatL_dVg + dVdsat_dVg + T7 * (1.0 + tmp2 * Vgsteff)
            - T8 * (Abulk * dVdsat_dVg - Abulk * Vdsat / Vgst2Vtm + Vdsat * dAbulk_dVg);
    dT0_dVb = dEsatL_dVb + dVdsat_dVb + T7 * tmp3 * Vgsteff - T8 * (dAbulk_dVb * Vdsat + Abulk * dVdsat_dVb);
    dT0_dVd = dEsatL_dVd + dVdsat_dVd - T8 * Abulk * dVdsat_dVd;
    T9 = WVCoxRds * Abulk;
    T1 = 2.0 / Lambda - 1.0 + T9;
    dT1_dVg = -2.0 * tmp1 + WVCoxRds * (Abulk * tmp2 + dAbulk_dVg);
    dT1_dVb = dAbulk_dVb * WVCoxRds + T9 * tmp3;
    Vasat = T0 / T1;
    dVasat_dVg = (dT0_dVg - Vasat * dT1_dVg) / T1;
    dVasat_dVb = (dT0_dVb - Vasat * dT1_dVb) / T1;
    dVasat_dVd = dT0_dVd / T1;

    if (Vdseff > Vds) Vdseff = Vds;
    diffVds = Vds - Vdseff;

    /* VACLM */
    if ((B3P(pclm) > 0.0) && (diffVds > 1.0e-10)) {
        T0 = 1.0 / (B3P(pclm) * Abulk * B3P(litl));
        dT0_dVb = -T0 / Abulk * dAbulk_dVb;
        dT0_dVg = -T0 / Abulk * dAbulk_dVg;
        T2 = Vgsteff / EsatL;
        T1 = Leff * (Abulk + T2);
        dT1_dVg = Leff * ((1.0 - T2 * dEsatL_dVg) / EsatL + dAbulk_dVg);
        dT1_dVb = Leff * (dAbulk_dVb - T2 * dEsatL_dVb / EsatL);
        dT1_dVd = -T2 * dEsatL_dVd / Esat;
        T9 = T0 * T1;
        VACLM = T9 * diffVds;
        dVACLM_dVg = T0 * dT1_dVg * diffVds - T9 * dVdseff_dVg + T1 * diffVds * dT0_dVg;
        dVACLM_dVb = (dT0_dVb * T1 + T0 * dT1_dVb) * diffVds - T9 * dVdseff_dVb;
        dVACLM_dVd = T0 * dT1_dVd * diffVds + T9 * (1.0 - dVdseff_dVd);
    } else {
        VACLM = B3_MAX_EXP;
        dVACLM_dVd = dVACLM_dVg = dVACLM_dVb = 0.0;
    }

    /* VADIBL */
    if (B3P(thetaRout) > 0.0) {
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
        T2 = B3P(thetaRout);
        VADIBL = (Vgst2Vtm - T0 / T1) / T2;
        dVADIBL_dVg = (1.0 - dT0_dVg / T1 + T0 * dT1_dVg / T9) / T2;
        dVADIBL_dVb = (-dT0_dVb / T1 + T0 * dT1_dVb / T9) / T2;
        dVADIBL_dVd = (-dT0_dVd / T1 + T0 * dT1_dVd / T9) / T2;
        T7 = B3P(pdiblb) * Vbseff;
        if (T7 >= -0.9) {
            T3 = 1.0 / (1.0 + T7);
            VADIBL *= T3;
            dVADIBL_dVg *= T3;
            dVADIBL_dVb = (dVADIBL_dVb - VADIBL * B3P(pdiblb)) * T3;
            dVADIBL_dVd *= T3;
        } else {
            T4 = 1.0 / (0.8 + T7);
            T3 = (17.0 + 20.0 * T7) * T4;
            dVADIBL_dVg *= T3;
            dVADIBL_dVb = dVADIBL_dVb * T3 - VADIBL * B3P(pdiblb) * T4 * T4;
            dVADIBL_dVd *= T3;
            VADIBL *= T3;
        }
    } else {
        VADIBL = B3_MAX_EXP;
        dVADIBL_dVd = dVADIBL_dVg = dVADIBL_dVb = 0.0;
    }

    /* VA */
    T8 = B3P(pvag) / EsatL;
    T9 = T8 * Vgsteff;
    if (T9 > -0.9) {
        T0 = 1.0 + T9;
        dT0_dVg = T8 * (1.0 - Vgsteff * dEsatL_dVg / EsatL);
        dT0_dVb = -T9 * dEsatL_dVb / EsatL;
        dT0_dVd = -T9 * dEsatL_dVd / EsatL;
    } else {
        T1 = 1.0 / (17.0 + 20.0 * T9);
        T0 = (0.8 + T9) * T1;
        T1 *= T1;
        dT0_dVg = T8 * (1.0 - Vgsteff * dEsatL_dVg / EsatL) * T1;
        T9 *= T1 / EsatL;
        dT0_dVb = -T9 * dEsatL_dVb;
        dT0_dVd = -T9 * dEsatL_dVd;
    }
    tmp1 = VACLM * VACLM;
    tmp2 = VADIBL * VADIBL;
    tmp3 = VACLM + VADIBL;
    T1 = VACLM * VADIBL / tmp3;
    tmp3 *= tmp3;
    dT1_dVg = (tmp1 * dVADIBL_dVg + tmp2 * dVACLM_dVg) / tmp3;
    dT1_dVd = (tmp1 * dVADIBL_dVd + tmp2 * dVACLM_dVd) / tmp3;
    dT1_dVb = (tmp1 * dVADIBL_dVb + tmp2 * dVACLM_dVb) / tmp3;
    Va = Vasat + T0 * T1;
    dVa_dVg = dVasat_dVg + T1 * dT0_dVg + T0 * dT1_dVg;
    dVa_dVd = dVasat_dVd + T1 * dT0_dVd + T0 * dT1_dVd;
    dVa_dVb = dVasat_dVb + T1 * dT0_dVb + T0 * dT1_dVb;

    /* VASCBE */
    if (B3P(pscbe2) > 0.0) {
        if (diffVds > B3P(pscbe1) * B3P(litl) / B3_EXPT) {
            T0 = B3P(pscbe1) * B3P(litl) / diffVds;
            VASCBE = Leff * ngb_exp(T0) / B3P(pscbe2);
            T1 = T0 * VASCBE / diffVds;
            dVASCBE_dVg = T1 * dVdseff_dVg;
            dVASCBE_dVd = -T1 * (1.0 - dVdseff_dVd);
            dVASCBE_dVb = T1 * dVdseff_dVb;
        } else {
            VASCBE = B3_MAX_EXP * Leff / B3P(pscbe2);
            dVASCBE_dVg = dVASCBE_dVd = dVASCBE_dVb = 0.0;
        }
    } else {
        VASCBE = B3_MAX_EXP;
        dVASCBE_dVg = dVASCBE_dVd = dVASCBE_dVb = 0.0;
    }

    /* drain current */
    CoxWovL = cox * Weff / Leff;
    beta = ueff * CoxWovL;
    dbeta_dVg = CoxWovL * dueff_dVg + beta * dWeff_dVg / Weff;
    dbeta_dVd = CoxWovL * dueff_dVd;
    dbeta_dVb = CoxWovL * dueff_dVb + beta * dWeff_dVb / Weff;

    T0 = 1.0 - 0.5 * Abulk * Vdseff / Vgst2Vtm;
    dT0_dVg = -0.5 * (Abulk * dVdseff_dVg - Abulk * Vdseff / Vgst2Vtm + Vdseff * dAbulk_dVg) / Vgst2Vtm;
    dT0_dVd = -0.5 * Abulk * dVdseff_dVd / Vgst2Vtm;
    dT0_dVb = -0.5 * (Abulk * dVdseff_dVb + dAbulk_dVb * Vdseff) / Vgst2Vtm;

    fgche1 = Vgsteff * T0;
    dfgche1_dVg = Vgsteff * dT0_dVg + T0;
    dfgche1_dVd = Vgsteff * dT0_dVd;
    dfgche1_dVb = Vgsteff * dT0_dVb;

    T9 = Vdseff / EsatL;
    fgche2 = 1.0 + T9;
    dfgche2_dVg = (dVdseff_dVg - T9 * dEsatL_dVg) / EsatL;
    dfgche2_dVd = (dVdseff_dVd - T9 * dEsatL_dVd) / EsatL;
    dfgche2_dVb = (dVdseff_dVb - T9 * dEsatL_dVb) / EsatL;

    gche = beta * fgche1 / fgche2;
    dgche_dVg = (beta * dfgche1_dVg + fgche1 * dbeta_dVg - gche * dfgche2_dVg) / fgche2;
    dgche_dVd = (beta * dfgche1_dVd + fgche1 * dbeta_dVd - gche * dfgche2_dVd) / fgche2;
    dgche_dVb = (beta * dfgche1_dVb + fgche1 * dbeta_dVb - gche * dfgche2_dVb) / fgche2;

    T0 = 1.0 + gche * Rds;
    T9 = Vdseff / T0;
    Idl = gche * T9;
    dIdl_dVg = (gche * dVdseff_dVg + T9 * dgche_dVg) / T0 - Idl * gche / T0 * dRds_dVg;
    dIdl_dVd = (gche * dVdseff_dVd + T9 * dgche_dVd) / T0;
    dIdl_dVb = (gche * dVdseff_dVb + T9 * dgche_dVb - Idl * dRds_dVb * gche) / T0;

    T9 = diffVds / Va;
    T0 = 1.0 + T9;
    Idsa = Idl * T0;
    dIdsa_dVg = T0 * dIdl_dVg - Idl * (dVdseff_dVg + T9 * dVa_dVg) / Va;
    dIdsa_dVd = T0 * dIdl_dVd + Idl * (1.0 - dVdseff_dVd - T9 * dVa_dVd) / Va;
    dIdsa_dVb = T0 * dIdl_dVb - Idl * (dVdseff_dVb + T9 * dVa_dVb) / Va;

    T9 = diffVds / VASCBE;
    T0 = 1.0 + T9;
    Ids = Idsa * T0;
    Gm = T0 * dIdsa_dVg - Idsa * (dVdseff_dVg + T9 * dVASCBE_dVg) / VASCBE;
    Gds = T0 * dIdsa_dVd + Idsa * (1.0 - dVdseff_dVd - T9 * dVASCBE_dVd) / VASCBE;
    Gmb = T0 * dIdsa_dVb - Idsa * (dVdseff_dVb + T9 * dVASCBE_dVb) / VASCBE;

    Gds += Gm * dVgsteff_dVd;
    Gmb += Gm * dVgsteff_dVb;
    Gm *= dVgsteff_dVg;
    Gmb *= dVbseff_dVb;

    /* substrate current */
    {
        const double tmp = B3P(alpha0) + B3P(alpha1) * Leff;
        if ((tmp <= 0.0) || (B3P(beta0) <= 0.0)) {
            Isub = Gbd = Gbb = Gbg = 0.0;
        } else {
            T2 = tmp / Leff;
            if (diffVds > B3P(beta0) / B3_EXPT) {
                T0 = -B3P(beta0) / diffVds;
                T1 = T2 * diffVds * ngb_exp(T0);
                T3 = T1 / diffVds * (T0 - 1.0);
                dT1_dVg = T3 * dVdseff_dVg;
                dT1_dVd = T3 * (dVdseff_dVd - 1.0);
                dT1_dVb = T3 * dVdseff_dVb;
            } else {
                T3 = T2 * B3_MIN_EXP;
                T1 = T3 * diffVds;
                dT1_dVg = -T3 * dVdseff_dVg;
                dT1_dVd = T3 * (1.0 - dVdseff_dVd);
                dT1_dVb = -T3 * dVdseff_dVb;
            }
            Isub = T1 * Idsa;
            Gbg = T1 * dIdsa_dVg + Idsa * dT1_dVg;
            Gbd = T1 * dIdsa_dVd + Idsa * dT1_dVd;
            Gbb = T1 * dIdsa_dVb + Idsa * dT1_dVb;
            Gbd += Gbg * dVgsteff_dVd;
            Gbb += Gbg * dVgsteff_dVb;
            Gbg *= dVgsteff_dVg;
            Gbb *= dVbseff_dVb;
        }
    }
    w->cdrain = Ids; w->gds = Gds; w->gm = Gm; w->gmbs = Gmb;
    w->gbbs = Gbb; w->gbgs = Gbg; w->gbds = Gbd; w->csub = Isub;
    w->Vbseff = Vbseff; w->dVbseff_dVb = dVbseff_dVb; w->Phis = Phis; w->dPhis_dVb = dPhis_dVb;
    w->sqrtPhis = sqrtPhis; w->dsqrtPhis_dVb = dsqrtPhis_dVb;
    w->Vth = Vth; w->dVth_dVb = dVth_dVb; w->dVth_dVd = dVth_dVd; w->Vgs_eff = Vgs_eff; w->dVgs_eff_dVg = dVgs_eff_dVg;
    w->Vgst = Vgst; w->n = n; w->dn_dVb = dn_dVb; w->dn_dVd = dn_dVd; w->Abulk0 = Abulk0; w->dAbulk0_dVb = dAbulk0_dVb;
    w->Vtm = Vtm;
}

/* intrinsic charges, capMod 0 (b3ld.c:1267-1600): piecewise Meyer-like model on Vfbcv */
NGB_HD void b3_charges_cm0(const B3Ctx *c, const double *mrow, const double *prw, size_t t, B3W *w)
{
    const double Vds = w->Vds, cox = B3M(cox), k1ox = B3P(k1ox), phi = B3P(phi), xpart = B3M(xpart);
    const double Vgs_eff = w->Vgs_eff, dVgs_eff_dVg = w->dVgs_eff_dVg;
    double Vbseff, dVbseff_dVb, Vfb, Vth, Vgst, dVth_dVb, CoxWL, Arg1;
    double qgate, qbulk, qdrn, cggb, cgdb, cgsb, cdgb, cddb, cdsb, cbgb, cbdb, cbsb;
    double T0, T1, T2, T3, T4, T5, T6, T7, T8, T9, T10, T11, T12, tmp, tmp1;
    (void)c; (void)t;
    if (w->Vbseff < 0.0) { Vbseff = w->Vbs; dVbseff_dVb = 1.0; }
    else { Vbseff = phi - w->Phis; dVbseff_dVb = -w->dPhis_dVb; }
    Vfb = B3P(vfbcv);
    Vth = Vfb + phi + k1ox * w->sqrtPhis;
    Vgst = Vgs_eff - Vth;
    dVth_dVb = k1ox * w->dsqrtPhis_dVb;
    CoxWL = cox * B3P(weffCV) * B3P(leffCV);
    Arg1 = Vgs_eff - Vbseff - Vfb;

    if (Arg1 <= 0.0) {
        qgate = CoxWL * Arg1;
        qbulk = -qgate;
        qdrn = 0.0;
        cggb = CoxWL * dVgs_eff_dVg;
        cgdb = 0.0;
        cgsb = CoxWL * (dVbseff_dVb - dVgs_eff_dVg);
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
        cgsb = T0 * (dVbseff_dVb - dVgs_eff_dVg);
        cdgb = 0.0; cddb = 0.0; cdsb = 0.0;
        cbgb = -cggb;
        cbdb = 0.0;
        cbsb = -cgsb;
    } else {
        const double One_Third_CoxWL = CoxWL / 3.0;
        const double Two_Third_CoxWL = 2.0 * One_Third_CoxWL;
        const double AbulkCV = w->Abulk0 * B3P(abulkCVfactor);
        const double dAbulkCV_dVb = B3P(abulkCVfactor) * w->dAbulk0_dVb;
        const double Vdsat = Vgst / AbulkCV;
        const double dVdsat_dVg = dVgs_eff_dVg / AbulkCV;
        const double dVdsat_dVb = -(Vdsat * dAbulkCV_dVb + dVth_dVb) / AbulkCV;
        double Alphaz, dAlphaz_dVg, dAlphaz_dVb;
        if (xpart > 0.5) {
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
                cddb = T4 * (2.0 - (1.0 / (3.0 * T1 * T1) + 2.0 * tmp) * T6 + T8 * (6.0 * Vdsat - 2.4 * Vds));
                cdsb = -(cdgb + T10 + cddb);
                T7 = 2.0 * (T1 + T3);
                qbulk = -(qgate - T4 * T7);
                T7 *= T9;
                T0 = 4.0 * T4 * (1.0 - T5);
                T12 = (-T7 * dAlphaz_dVg - cdgb - T0 * dVdsat_dVg) * dVgs_eff_dVg;
                T11 = -T7 * dAlphaz_dVb - T10 - T0 * dVdsat_dVb;
                T10 = -4.0 * T4 * (T2 - 0.5 + 0.5 * T5) - cddb;
                tmp = -(T10 + T11 + T12);
                cbgb = -(cggb + cdgb + T12);
                cbdb = -(cgdb + cddb + T10);
                cbsb = -(cgsb + cdsb + tmp);
            }
        } else {
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
    w->qgate = qgate; w->qbulk = qbulk; w->qdrn = qdrn;
    w->cggb = cggb; w->cgsb = cgsb; w->cgdb = cgdb; w->cdgb = cdgb; w->cdsb = cdsb; w->cddb = cddb;
    w->cbgb = cbgb; w->cbsb = cbsb; w->cbdb = cbdb;
}

/* intrinsic charges and capacitances, capMod 2 (b3ld.c:1730-1947) and capMod 3, the
 * charge-thickness model (:1950-2238); both start from the CV version of Vgsteff (:1735-1768) */
NGB_HD void b3_charges(const B3Ctx *c, const double *mrow, const double *prw, size_t t, B3W *w)
{
    const double Vds = w->Vds, Vtm = w->Vtm, cox = B3M(cox), k1ox = B3P(k1ox), phi = B3P(phi);
    const double xpart = B3M(xpart);
    const int capMod = (int)B3M(capMod);
    const double Vgs_eff = w->Vgs_eff, dVgs_eff_dVg = w->dVgs_eff_dVg, Vgst = w->Vgst, Vth = w->Vth;
    const double dVth_dVd = w->dVth_dVd, dVth_dVb = w->dVth_dVb, dVbseff_dVb = w->dVbseff_dVb;
    const double vfbzb = B3I(vfbzb);
    double VbseffCV, dVbseffCV_dVb, CoxWL, noff, dnoff_dVd, dnoff_dVb, voffcv, VgstNVt;
    double Vgsteff, dVgsteff_dVg, dVgsteff_dVd, dVgsteff_dVb;
    double T0, T1, T2, T3, T4, T5, T6, T7, T8, T9, T10, T11, T12, V3, V4, tmp;
    double Vfbeff, dVfbeff_dVg, dVfbeff_dVb, Qac0, dQac0_dVg, dQac0_dVb, Qsub0, dQsub0_dVg, dQsub0_dVd, dQsub0_dVb;
    double AbulkCV, dAbulkCV_dVb, VdsatCV, VdseffCV, dVdseffCV_dVg, dVdseffCV_dVd, dVdseffCV_dVb;
    double qgate, qbulk, qsrc, qdrn, Cgg1, Cgd1, Cgb1, Cbg1, Cbd1, Cbb1, Csg, Csd, Csb;
    double Cgg, Cgd, Cgb, Cbg, Cbd, Cbb;

    if (w->Vbseff < 0.0) { VbseffCV = w->Vbseff; dVbseffCV_dVb = 1.0; }
    else { VbseffCV = phi - w->Phis; dVbseffCV_dVb = -w->dPhis_dVb; }
    CoxWL = cox * B3P(weffCV) * B3P(leffCV);

    noff = w->n * B3P(noff);
    dnoff_dVd = B3P(noff) * w->dn_dVd;
    dnoff_dVb = B3P(noff) * w->dn_dVb;
    T0 = Vtm * noff;
    voffcv = B3P(voffcv);
    VgstNVt = (Vgst - voffcv) / T0;
    if (VgstNVt > B3_EXPT) {
        Vgsteff = Vgst - voffcv;
        dVgsteff_dVg = dVgs_eff_dVg;
        dVgsteff_dVd = -dVth_dVd;
        dVgsteff_dVb = -dVth_dVb;
    } else if (VgstNVt < -B3_EXPT) {
        Vgsteff = T0 * ngb_log(1.0 + B3_MIN_EXP);
        dVgsteff_dVg = 0.0;
        dVgsteff_dVd = Vgsteff / noff;
        dVgsteff_dVb = dVgsteff_dVd * dnoff_dVb;
        dVgsteff_dVd *= dnoff_dVd;
    } else {
        const double ExpVgst = ngb_exp(VgstNVt);
        Vgsteff = T0 * ngb_log(1.0 + ExpVgst);
        dVgsteff_dVg = ExpVgst / (1.0 + ExpVgst);
        dVgsteff_dVd = -dVgsteff_dVg * (dVth_dVd + (Vgst - voffcv) / noff * dnoff_dVd) + Vgsteff / noff * dnoff_dVd;
        dVgsteff_dVb = -dVgsteff_dVg * (dVth_dVb + (Vgst - voffcv) / noff * dnoff_dVb) + Vgsteff / noff * dnoff_dVb;
        dVgsteff_dVg *= dVgs_eff_dVg;
    }

    if (capMod == 1) {
        /* b3ld.c:1770-1945 */
        const double Vfb = vfbzb;
        const double Arg1 = Vgs_eff - VbseffCV - Vfb - Vgsteff;
        double One_Third_CoxWL, Two_Third_CoxWL, dVdsatCV_dVg, dVdsatCV_dVb, dT0_dVg, dT0_dVb, dT3_dVg, dT3_dVd, dT3_dVb;
        if (Arg1 <= 0.0) {
            qgate = CoxWL * Arg1;
            Cgg = CoxWL * (dVgs_eff_dVg - dVgsteff_dVg);
            Cgd = -CoxWL * dVgsteff_dVd;
            Cgb = -CoxWL * (dVbseffCV_dVb + dVgsteff_dVb);
        } else {
            T0 = 0.5 * k1ox;
            T1 = sqrt(T0 * T0 + Arg1);
            T2 = CoxWL * T0 / T1;
            qgate = CoxWL * k1ox * (T1 - T0);
            Cgg = T2 * (dVgs_eff_dVg - dVgsteff_dVg);
            Cgd = -T2 * dVgsteff_dVd;
            Cgb = -T2 * (dVbseffCV_dVb + dVgsteff_dVb);
        }
        qbulk = -qgate;
        Cbg = -Cgg;
        Cbd = -Cgd;
        Cbb = -Cgb;

        One_Third_CoxWL = CoxWL / 3.0;
        Two_Third_CoxWL = 2.0 * One_Third_CoxWL;
        AbulkCV = w->Abulk0 * B3P(abulkCVfactor);
        dAbulkCV_dVb = B3P(abulkCVfactor) * w->dAbulk0_dVb;
        VdsatCV = Vgsteff / AbulkCV;
        if (VdsatCV < Vds) {
            dVdsatCV_dVg = 1.0 / AbulkCV;
            dVdsatCV_dVb = -VdsatCV * dAbulkCV_dVb / AbulkCV;
            T0 = Vgsteff - VdsatCV / 3.0;
            dT0_dVg = 1.0 - dVdsatCV_dVg / 3.0;
            dT0_dVb = -dVdsatCV_dVb / 3.0;
            qgate += CoxWL * T0;
            Cgg1 = CoxWL * dT0_dVg;
            Cgb1 = CoxWL * dT0_dVb + Cgg1 * dVgsteff_dVb;
            Cgd1 = Cgg1 * dVgsteff_dVd;
            Cgg1 *= dVgsteff_dVg;
            Cgg += Cgg1;
            Cgb += Cgb1;
            Cgd += Cgd1;

            T0 = VdsatCV - Vgsteff;
            dT0_dVg = dVdsatCV_dVg - 1.0;
            dT0_dVb = dVdsatCV_dVb;
            qbulk += One_Third_CoxWL * T0;
            Cbg1 = One_Third_CoxWL * dT0_dVg;
            Cbb1 = One_Third_CoxWL * dT0_dVb + Cbg1 * dVgsteff_dVb;
            Cbd1 = Cbg1 * dVgsteff_dVd;
            Cbg1 *= dVgsteff_dVg;
            Cbg += Cbg1;
            Cbb += Cbb1;
            Cbd += Cbd1;

            if (xpart > 0.5) T0 = -Two_Third_CoxWL;
            else if (xpart < 0.5) T0 = -0.4 * CoxWL;
            else T0 = -One_Third_CoxWL;
            qsrc = T0 * Vgsteff;
            Csg = T0 * dVgsteff_dVg;
            Csb = T0 * dVgsteff_dVb;
            Csd = T0 * dVgsteff_dVd;
        } else {
            T0 = AbulkCV * Vds;
            T1 = 12.0 * (Vgsteff - 0.5 * T0 + 1.e-20);
            T2 = Vds / T1;
            T3 = T0 * T2;
            dT3_dVg = -12.0 * T2 * T2 * AbulkCV;
            dT3_dVd = 6.0 * T0 * (4.0 * Vgsteff - T0) / T1 / T1 - 0.5;
            dT3_dVb = 12.0 * T2 * T2 * dAbulkCV_dVb * Vgsteff;

            qgate += CoxWL * (Vgsteff - 0.5 * Vds + T3);
            Cgg1 = CoxWL * (1.0 + dT3_dVg);
            Cgb1 = CoxWL * dT3_dVb + Cgg1 * dVgsteff_dVb;
            Cgd1 = CoxWL * dT3_dVd + Cgg1 * dVgsteff_dVd;
            Cgg1 *= dVgsteff_dVg;
            Cgg += Cgg1;
            Cgb += Cgb1;
            Cgd += Cgd1;

            qbulk += CoxWL * (1.0 - AbulkCV) * (0.5 * Vds - T3);
            Cbg1 = -CoxWL * ((1.0 - AbulkCV) * dT3_dVg);
            Cbb1 = -CoxWL * ((1.0 - AbulkCV) * dT3_dVb + (0.5 * Vds - T3) * dAbulkCV_dVb) + Cbg1 * dVgsteff_dVb;
            Cbd1 = -CoxWL * (1.0 - AbulkCV) * dT3_dVd + Cbg1 * dVgsteff_dVd;
            Cbg1 *= dVgsteff_dVg;
            Cbg += Cbg1;
            Cbb += Cbb1;
            Cbd += Cbd1;

            if (xpart > 0.5) {
                T1 = T1 + T1;
                qsrc = -CoxWL * (0.5 * Vgsteff + 0.25 * T0 - T0 * T0 / T1);
                Csg = -CoxWL * (0.5 + 24.0 * T0 * Vds / T1 / T1 * AbulkCV);
                Csb = -CoxWL * (0.25 * Vds * dAbulkCV_dVb - 12.0 * T0 * Vds / T1 / T1 * (4.0 * Vgsteff - T0) * dAbulkCV_dVb)
                    + Csg * dVgsteff_dVb;
                Csd = -CoxWL * (0.25 * AbulkCV - 12.0 * AbulkCV * T0 / T1 / T1 * (4.0 * Vgsteff - T0)) + Csg * dVgsteff_dVd;
                Csg *= dVgsteff_dVg;
            } else if (xpart < 0.5) {
                T1 = T1 / 12.0;
                T2 = 0.5 * CoxWL / (T1 * T1);
                T3 = Vgsteff * (2.0 * T0 * T0 / 3.0 + Vgsteff * (Vgsteff - 4.0 * T0 / 3.0)) - 2.0 * T0 * T0 * T0 / 15.0;
                qsrc = -T2 * T3;
                T4 = 4.0 / 3.0 * Vgsteff * (Vgsteff - T0) + 0.4 * T0 * T0;
                Csg = -2.0 * qsrc / T1 - T2 * (Vgsteff * (3.0 * Vgsteff - 8.0 * T0 / 3.0) + 2.0 * T0 * T0 / 3.0);
                Csb = (qsrc / T1 * Vds + T2 * T4 * Vds) * dAbulkCV_dVb + Csg * dVgsteff_dVb;
                Csd = (qsrc / T1 + T2 * T4) * AbulkCV + Csg * dVgsteff_dVd;
                Csg *= dVgsteff_dVg;
            } else {
                qsrc = -0.5 * (qgate + qbulk);
                Csg = -0.5 * (Cgg1 + Cbg1);
                Csb = -0.5 * (Cgb1 + Cbb1);
                Csd = -0.5 * (Cgd1 + Cbd1);
            }
        }
        qdrn = -(qgate + qbulk + qsrc);
        goto store;
    }

    /* accumulation charge through the smoothed flat-band voltage (common to both models) */
    V3 = vfbzb - Vgs_eff + VbseffCV - B3_DELTA;
    if (vfbzb <= 0.0) { T0 = sqrt(V3 * V3 - 4.0 * B3_DELTA * vfbzb); T2 = -B3_DELTA / T0; }
    else { T0 = sqrt(V3 * V3 + 4.0 * B3_DELTA * vfbzb); T2 = B3_DELTA / T0; }
    T1 = 0.5 * (1.0 + V3 / T0);
    Vfbeff = vfbzb - 0.5 * (V3 + T0);
    dVfbeff_dVg = T1 * dVgs_eff_dVg;
    dVfbeff_dVb = -T1 * dVbseffCV_dVb;

    if (capMod == 2) {
        Qac0 = CoxWL * (Vfbeff - vfbzb);
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

        AbulkCV = w->Abulk0 * B3P(abulkCVfactor);
        dAbulkCV_dVb = B3P(abulkCVfactor) * w->dAbulk0_dVb;
        VdsatCV = Vgsteff / AbulkCV;

        V4 = VdsatCV - Vds - B3_DELTA;
        T0 = sqrt(V4 * V4 + 4.0 * B3_DELTA * VdsatCV);
        VdseffCV = VdsatCV - 0.5 * (V4 + T0);
        T1 = 0.5 * (1.0 + V4 / T0);
        T2 = B3_DELTA / T0;
        T3 = (1.0 - T1 - T2) / AbulkCV;
        dVdseffCV_dVg = T3;
        dVdseffCV_dVd = T1;
        dVdseffCV_dVb = -T3 * VdsatCV * dAbulkCV_dVb;
        if (Vds == 0.0) { VdseffCV = 0.0; dVdseffCV_dVg = 0.0; dVdseffCV_dVb = 0.0; }

        T0 = AbulkCV * VdseffCV;
        T1 = 12.0 * (Vgsteff - 0.5 * T0 + 1e-20);
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
            T3 = Vgsteff * (2.0 * T0 * T0 / 3.0 + Vgsteff * (Vgsteff - 4.0 * T0 / 3.0)) - 2.0 * T0 * T0 * T0 / 15.0;
            qsrc = -T2 * T3;
            T7 = 4.0 / 3.0 * Vgsteff * (Vgsteff - T0) + 0.4 * T0 * T0;
            T4 = -2.0 * qsrc / T1 - T2 * (Vgsteff * (3.0 * Vgsteff - 8.0 * T0 / 3.0) + 2.0 * T0 * T0 / 3.0);
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
    } else {
        /* capMod 3: finite charge-layer thickness */
        const double ldeb = B3P(ldeb);
        double Tox, Tcen, dTcen_dVg, dTcen_dVd, dTcen_dVb, LINK, Ccen, Coxeff, dCoxeff_dVg, dCoxeff_dVd, dCoxeff_dVb;
        double CoxWLcen, QovCox, Denomi, DeltaPhi, dDeltaPhi_dVg, VgDP, dVgDP_dVg, dT0_dVg, dT0_dVd, dT0_dVb;
        double dT1_dVg, dT1_dVd, dT1_dVb;

        Tox = 1.0e8 * B3M(tox);
        T0 = (Vgs_eff - VbseffCV - vfbzb) / Tox;
        dT0_dVg = dVgs_eff_dVg / Tox;
        dT0_dVb = -dVbseffCV_dVb / Tox;

        tmp = T0 * B3P(acde);
        if ((-B3_EXPT < tmp) && (tmp < B3_EXPT)) {
            Tcen = ldeb * ngb_exp(tmp);
            dTcen_dVg = B3P(acde) * Tcen;
            dTcen_dVb = dTcen_dVg * dT0_dVb;
            dTcen_dVg *= dT0_dVg;
        } else if (tmp <= -B3_EXPT) {
            Tcen = ldeb * B3_MIN_EXP;
            dTcen_dVg = dTcen_dVb = 0.0;
        } else {
            Tcen = ldeb * B3_MAX_EXP;
            dTcen_dVg = dTcen_dVb = 0.0;
        }

        LINK = 1.0e-3 * B3M(tox);
        V3 = ldeb - Tcen - LINK;
        V4 = sqrt(V3 * V3 + 4.0 * LINK * ldeb);
        Tcen = ldeb - 0.5 * (V3 + V4);
        T1 = 0.5 * (1.0 + V3 / V4);
        dTcen_dVg *= T1;
        dTcen_dVb *= T1;

        Ccen = B3_EPSSI / Tcen;
        T2 = cox / (cox + Ccen);
        Coxeff = T2 * Ccen;
        T3 = -Ccen / Tcen;
        dCoxeff_dVg = T2 * T2 * T3;
        dCoxeff_dVb = dCoxeff_dVg * dTcen_dVb;
        dCoxeff_dVg *= dTcen_dVg;
        CoxWLcen = CoxWL * Coxeff / cox;

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

        /* gate-bias dependent surface-potential increase */
        if (k1ox <= 0.0) { Denomi = 0.25 * B3P(moin) * Vtm; T0 = 0.5 * B3P(sqrtPhi); }
        else { Denomi = B3P(moin) * Vtm * k1ox * k1ox; T0 = k1ox * B3P(sqrtPhi); }
        T1 = 2.0 * T0 + Vgsteff;
        DeltaPhi = Vtm * ngb_log(1.0 + T1 * Vgsteff / Denomi);
        dDeltaPhi_dVg = 2.0 * Vtm * (T1 - T0) / (Denomi + T1 * Vgsteff);

        T0 = Vgsteff - DeltaPhi - 0.001;
        dT0_dVg = 1.0 - dDeltaPhi_dVg;
        T1 = sqrt(T0 * T0 + Vgsteff * 0.004);
        VgDP = 0.5 * (T0 + T1);
        dVgDP_dVg = 0.5 * (dT0_dVg + (T0 * dT0_dVg + 0.002) / T1);

        T3 = 4.0 * (Vth - vfbzb - phi);
        Tox += Tox;
        if (T3 >= 0.0) {
            T0 = (Vgsteff + T3) / Tox;
            dT0_dVd = (dVgsteff_dVd + 4.0 * dVth_dVd) / Tox;
            dT0_dVb = (dVgsteff_dVb + 4.0 * dVth_dVb) / Tox;
        } else {
            T0 = (Vgsteff + 1.0e-20) / Tox;
            dT0_dVd = dVgsteff_dVd / Tox;
            dT0_dVb = dVgsteff_dVb / Tox;
        }
        tmp = ngb_exp(0.7 * ngb_log(T0));
        T1 = 1.0 + tmp;
        T2 = 0.7 * tmp / (T0 * Tox);
        Tcen = 1.9e-9 / T1;
        dTcen_dVg = -1.9e-9 * T2 / T1 / T1;
        dTcen_dVd = Tox * dTcen_dVg;
        dTcen_dVb = dTcen_dVd * dT0_dVb;
        dTcen_dVd *= dT0_dVd;
        dTcen_dVg *= dVgsteff_dVg;

        Ccen = B3_EPSSI / Tcen;
        T0 = cox / (cox + Ccen);
        Coxeff = T0 * Ccen;
        T1 = -Ccen / Tcen;
        dCoxeff_dVg = T0 * T0 * T1;
        dCoxeff_dVd = dCoxeff_dVg * dTcen_dVd;
        dCoxeff_dVb = dCoxeff_dVg * dTcen_dVb;
        dCoxeff_dVg *= dTcen_dVg;
        CoxWLcen = CoxWL * Coxeff / cox;

        AbulkCV = w->Abulk0 * B3P(abulkCVfactor);
        dAbulkCV_dVb = B3P(abulkCVfactor) * w->dAbulk0_dVb;
        VdsatCV = VgDP / AbulkCV;
        T0 = VdsatCV - Vds - B3_DELTA;
        dT0_dVg = dVgDP_dVg / AbulkCV;
        dT0_dVb = -VdsatCV * dAbulkCV_dVb / AbulkCV;
        T1 = sqrt(T0 * T0 + 4.0 * B3_DELTA * VdsatCV);
        dT1_dVg = (T0 + B3_DELTA + B3_DELTA) / T1;
        dT1_dVd = -T0 / T1;
        dT1_dVb = dT1_dVg * dT0_dVb;
        dT1_dVg *= dT0_dVg;
        if (T0 >= 0.0) {
            VdseffCV = VdsatCV - 0.5 * (T0 + T1);
            dVdseffCV_dVg = 0.5 * (dT0_dVg - dT1_dVg);
            dVdseffCV_dVd = 0.5 * (1.0 - dT1_dVd);
            dVdseffCV_dVb = 0.5 * (dT0_dVb - dT1_dVb);
        } else {
            T3 = (B3_DELTA + B3_DELTA) / (T1 - T0);
            T4 = 1.0 - T3;
            T5 = VdsatCV * T3 / (T1 - T0);
            VdseffCV = VdsatCV * T4;
            dVdseffCV_dVg = dT0_dVg * T4 + T5 * (dT1_dVg - dT0_dVg);
            dVdseffCV_dVd = T5 * (dT1_dVd + 1.0);
            dVdseffCV_dVb = dT0_dVb * (1.0 - T5) + T5 * dT1_dVb;
        }
        if (Vds == 0.0) { VdseffCV = 0.0; dVdseffCV_dVg = 0.0; dVdseffCV_dVb = 0.0; }

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
        Cgb1 = CoxWLcen * (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Cgg1 * dVgsteff_dVb + QovCox * dCoxeff_dVb;
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
        Cbb1 = CoxWLcen * (T11 * dVdseffCV_dVb + T12 * dAbulkCV_dVb) + Cbg1 * dVgsteff_dVb + QovCox * dCoxeff_dVb;
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
            Csb = CoxWLcen * (T5 * dVdseffCV_dVb + T6 * dAbulkCV_dVb) + Csg * dVgsteff_dVb + QovCox * dCoxeff_dVb;
            Csg = Csg * dVgsteff_dVg + QovCox * dCoxeff_dVg;
        } else if (xpart < 0.5) {
            T2 = T2 / 12.0;
            T3 = 0.5 * CoxWLcen / (T2 * T2);
            T4 = T1 * (2.0 * T0 * T0 / 3.0 + T1 * (T1 - 4.0 * T0 / 3.0)) - 2.0 * T0 * T0 * T0 / 15.0;
            qsrc = -T3 * T4;
            QovCox = qsrc / Coxeff;
            T8 = 4.0 / 3.0 * T1 * (T1 - T0) + 0.4 * T0 * T0;
            T5 = -2.0 * qsrc / T2 - T3 * (T1 * (3.0 * T1 - 8.0 * T0 / 3.0) + 2.0 * T0 * T0 / 3.0);
            T6 = AbulkCV * (qsrc / T2 + T3 * T8);
            T7 = T6 * VdseffCV / AbulkCV;
            Csg = T5 * dVgDP_dVg + T6 * dVdseffCV_dVg;
            Csd = Csg * dVgsteff_dVd + T6 * dVdseffCV_dVd + QovCox * dCoxeff_dVd;
            Csb = Csg * dVgsteff_dVb + T6 * dVdseffCV_dVb + T7 * dAbulkCV_dVb + QovCox * dCoxeff_dVb;
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
    }
store:
    Cgb *= dVbseff_dVb;
    Cbb *= dVbseff_dVb;
    Csb *= dVbseff_dVb;

    w->qgate = qgate; w->qbulk = qbulk; w->qdrn = qdrn;
    w->cggb = Cgg;
    w->cgsb = -(Cgg + Cgd + Cgb);
    w->cgdb = Cgd;
    w->cdgb = -(Cgg + Cbg + Csg);
    w->cdsb = (Cgg + Cgd + Cgb + Cbg + Cbd + Cbb + Csg + Csd + Csb);
    w->cddb = -(Cgd + Cbd + Csd);
    w->cbgb = Cbg;
    w->cbsb = -(Cbg + Cbd + Cbb);
    w->cbdb = Cbd;
}

/* depletion charge and capacitance of one bulk junction: b3ld.c:2333-2431 */
NGB_HD void b3_junction_cv(double v, double cz, double czsw, double czswg, const double *mrow, double *q, double *cap)
{
    const double MJ = B3M(bulkJctBotGradingCoeff), MJSW = B3M(bulkJctSideGradingCoeff), MJSWG = B3M(bulkJctGateSideGradingCoeff);
    if (v == 0.0) {
        *q = 0.0;
        *cap = cz + czsw + czswg;
    } else if (v < 0.0) {
        double arg, sarg;
        if (cz > 0.0) {
            arg = 1.0 - v / B3M(PhiB);
            sarg = (MJ == 0.5) ? 1.0 / sqrt(arg) : ngb_exp(-MJ * ngb_log(arg));
            *q = B3M(PhiB) * cz * (1.0 - arg * sarg) / (1.0 - MJ);
            *cap = cz * sarg;
        } else {
            *q = 0.0;
            *cap = 0.0;
        }
        if (czsw > 0.0) {
            arg = 1.0 - v / B3M(PhiBSW);
            sarg = (MJSW == 0.5) ? 1.0 / sqrt(arg) : ngb_exp(-MJSW * ngb_log(arg));
            *q += B3M(PhiBSW) * czsw * (1.0 - arg * sarg) / (1.0 - MJSW);
            *cap += czsw * sarg;
        }
        if (czswg > 0.0) {
            arg = 1.0 - v / B3M(PhiBSWG);
            sarg = (MJSWG == 0.5) ? 1.0 / sqrt(arg) : ngb_exp(-MJSWG * ngb_log(arg));
            *q += B3M(PhiBSWG) * czswg * (1.0 - arg * sarg) / (1.0 - MJSWG);
            *cap += czswg * sarg;
        }
    } else {
        const double T0 = cz + czsw + czswg;
        const double T1 = v * (cz * MJ / B3M(PhiB) + czsw * MJSW / B3M(PhiBSW) + czswg * MJSWG / B3M(PhiBSWG));
        *q = v * (T0 + 0.5 * T1);
        *cap = T0 + T1;
    }
}

/* bias-dependent overlap capacitance and charge of one side (capMod 2 and 3): b3ld.c:2529-2551 */
NGB_HD void b3_overlap(double v, double cgo, double cgl_w, double ckappa, double *cap, double *q)
{
    const double T0 = v + B3_DELTA;
    const double T1 = sqrt(T0 * T0 + 4.0 * B3_DELTA);
    const double T2 = 0.5 * (T0 - T1);
    const double T3 = cgl_w;
    const double T4 = sqrt(1.0 - 4.0 * T2 / ckappa);
    *cap = cgo + T3 - T3 * (1.0 - 1.0 / T4) * (0.5 - 0.5 * T0 / T1);
    *q = (cgo + T3) * v - T3 * (T2 + 0.5 * ckappa * (T4 - 1.0));
}

/* overlap capacitance and charge of one side, capMod 1: b3ld.c:2505-2532 */
NGB_HD void b3_overlap_cm1(double v, double cgo, double weffCV, double cgl, double ckappa, double *cap, double *q)
{
    if (v < 0.0) {
        const double T1 = sqrt(1.0 - 4.0 * v / ckappa);
        *cap = cgo + weffCV * cgl / T1;
        *q = cgo * v - weffCV * 0.5 * cgl * ckappa * (T1 - 1.0);
    } else {
        *cap = cgo + weffCV * cgl;
        *q = (weffCV * cgl + cgo) * v;
    }
}

NGB_HD int b3_load_thread(const B3Ctx *c, size_t t)
{
    const int S = c->S;
    const int inst = (int)(t / (size_t)S);
    const int s = (int)(t - (size_t)inst * S);
    if (!NGB_LDG(&c->ctl.active[s])) return NGB_OK;
    const int mode = NGB_LDG(&c->ctl.mode[s]);
    const int head = NGB_LDG(&c->ctl.head[s]);
    const int nh = c->ctl.nhist;
    const int fl = NGB_LDG(&c->flags[inst]);
    const int off = fl & B3F_OFF;
    const double *mrow = c->mtab + (size_t)NGB_LDG(&c->prow[inst]) * B3M_COUNT;
    const double *prw = c->ptab + (size_t)NGB_LDG(&c->prow[inst]) * B3P_COUNT;
    const double type = B3M(type);
    const double gmin = NGB_LDG(&c->ctl.gmin[s]);
    const int ChargeComputationNeeded =
        ((mode & (NGB_MODEDCTRANCURVE | NGB_MODEAC | NGB_MODETRAN | NGB_MODEINITSMSIG)) ||
         ((mode & NGB_MODETRANOP) && (mode & NGB_MODEUIC))) ? 1 : 0;
    B3W w;
    double vbs, vgs, vds, vbd, vgd, vgb, qdef;
    double gbs, cbs, gbd, cbd, capbs = 0.0, capbd = 0.0;
    int Check = 1, b3mode;
#define B3ST(h, k) c->state[((size_t)(((head) + (h)) % nh) * B3ST_COUNT + (k)) * c->T + t]
    {   /* deferred whole-vector state copies of DCtran (dctran.c:319-322, 711-716) */
        const int sop = NGB_LDG(&c->ctl.stateop[s]);
        if (sop) {
            for (int k = 0; k < B3ST_COUNT; k++) {
                if (sop & NGB_OP_COPY01) B3ST(1, k) = B3ST(0, k);
                if (sop & NGB_OP_COPY1_23) { const double v = B3ST(1, k); B3ST(2, k) = v; if (nh > 3) B3ST(3, k) = v; }
                if (sop & NGB_OP_COPY23) { const double v = B3ST(2, k); B3ST(0, k) = v; if (nh > 3) B3ST(3, k) = v; }
            }
        }
    }

    /* terminal voltages of this iteration */
    if (mode & NGB_MODEINITSMSIG) {
        vbs = B3ST(0, B3ST_vbs); vgs = B3ST(0, B3ST_vgs); vds = B3ST(0, B3ST_vds); qdef = B3ST(0, B3ST_qdef);
    } else if (mode & NGB_MODEINITTRAN) {
        vbs = B3ST(1, B3ST_vbs); vgs = B3ST(1, B3ST_vgs); vds = B3ST(1, B3ST_vds); qdef = B3ST(1, B3ST_qdef);
    } else if ((mode & NGB_MODEINITJCT) && !off) {
        vds = type * B3I(icVDS);
        vgs = type * B3I(icVGS);
        vbs = type * B3I(icVBS);
        qdef = 0.0;
        if ((vds == 0.0) && (vgs == 0.0) && (vbs == 0.0) &&
            ((mode & (NGB_MODETRAN | NGB_MODEAC | NGB_MODEDCOP | NGB_MODEDCTRANCURVE)) || !(mode & NGB_MODEUIC))) {
            vbs = 0.0;
            vgs = type * B3I(vth0) + 0.1;
            vds = 0.1;
        }
    } else if ((mode & (NGB_MODEINITJCT | NGB_MODEINITFIX)) && off) {
        qdef = vbs = vgs = vds = 0.0;
    } else {
        double vgdo;
        if (mode & NGB_MODEINITPRED) {
            const double xfact = NGB_LDG(&c->ctl.delta[s]) / NGB_LDG(&c->ctl.delta_old[(size_t)1 * S + s]);
            B3ST(0, B3ST_vbs) = B3ST(1, B3ST_vbs);
            vbs = (1.0 + xfact) * B3ST(1, B3ST_vbs) - (xfact * B3ST(2, B3ST_vbs));
            B3ST(0, B3ST_vgs) = B3ST(1, B3ST_vgs);
            vgs = (1.0 + xfact) * B3ST(1, B3ST_vgs) - (xfact * B3ST(2, B3ST_vgs));
            B3ST(0, B3ST_vds) = B3ST(1, B3ST_vds);
            vds = (1.0 + xfact) * B3ST(1, B3ST_vds) - (xfact * B3ST(2, B3ST_vds));
            B3ST(0, B3ST_vbd) = B3ST(0, B3ST_vbs) - B3ST(0, B3ST_vds);
            B3ST(0, B3ST_qdef) = B3ST(1, B3ST_qdef);
            qdef = (1.0 + xfact) * B3ST(1, B3ST_qdef) - (xfact * B3ST(2, B3ST_qdef));
        } else {
            const double *xo = c->x + (size_t)NGB_LDG(&c->ctl.xsel[s]) * c->neq1 * S;
#define XV(role) NGB_LDG(&xo[(size_t)NGB_LDG(&c->nodes[(role) * c->ninst + inst]) * S + s])
            const double vsp = XV(B3N_sp);
            vbs = type * (XV(B3N_b) - vsp);
            vgs = type * (XV(B3N_g) - vsp);
            vds = type * (XV(B3N_dp) - vsp);
            qdef = type * 0.0;                         /* rhsOld[qNode = 0] */
#undef XV
        }
        vbd = vbs - vds;
        vgd = vgs - vds;
        vgdo = B3ST(0, B3ST_vgs) - B3ST(0, B3ST_vds);
        {
            const double von = NGB_LDG(&c->von[t]);
            if (B3ST(0, B3ST_vds) >= 0.0) {
                vgs = ngb_fetlim(vgs, B3ST(0, B3ST_vgs), von);
                vds = vgs - vgd;
                vds = ngb_limvds(vds, B3ST(0, B3ST_vds));
                vgd = vgs - vds;
            } else {
                vgd = ngb_fetlim(vgd, vgdo, von);
                vds = vgs - vgd;
                vds = -ngb_limvds(-vds, -(B3ST(0, B3ST_vds)));
                vgs = vgd + vds;
            }
        }
        if (vds >= 0.0) {
            vbs = ngb_pnjlim(vbs, B3ST(0, B3ST_vbs), c->vt0, B3M(vcrit), &Check);
            vbd = vbs - vds;
        } else {
            vbd = ngb_pnjlim(vbd, B3ST(0, B3ST_vbd), c->vt0, B3M(vcrit), &Check);
            vbs = vbd + vds;
        }
    }
    vbd = vbs - vds;
    vgd = vgs - vds;
    vgb = vgs - vbs;

    /* source / drain junction diodes */
    {
        const double Nvtm = B3M(vtm) * B3M(jctEmissionCoeff);
        const double as = B3I(sourceArea), ps = B3I(sourcePerimeter), ad = B3I(drainArea), pd = B3I(drainPerimeter);
        double isats, isatd;
        if ((as <= 0.0) && (ps <= 0.0)) isats = 1.0e-14;
        else isats = as * B3M(jctTempSatCurDensity) + ps * B3M(jctSidewallTempSatCurDensity);
        if ((ad <= 0.0) && (pd <= 0.0)) isatd = 1.0e-14;
        else isatd = ad * B3M(jctTempSatCurDensity) + pd * B3M(jctSidewallTempSatCurDensity);
        b3_junction_dc(isats, vbs, Nvtm, B3M(ijth), B3I(vjsm), B3I(IsEvjsm), gmin, &gbs, &cbs);
        b3_junction_dc(isatd, vbd, Nvtm, B3M(ijth), B3I(vjdm), B3I(IsEvjdm), gmin, &gbd, &cbd);
    }

    if (vds >= 0.0) { b3mode = 1; w.Vds = vds; w.Vgs = vgs; w.Vbs = vbs; }
    else { b3mode = -1; w.Vds = -vds; w.Vgs = vgd; w.Vbs = vbd; }
    {
        double von_new;
        b3_core_dc(c, mrow, prw, t, &w, &von_new);
        c->von[t] = von_new;
    }

    w.qgate = w.qbulk = w.qdrn = 0.0;
    w.cggb = w.cgsb = w.cgdb = w.cdgb = w.cdsb = w.cddb = w.cbgb = w.cbsb = w.cbdb = 0.0;
    if (!((B3M(xpart) < 0) || !ChargeComputationNeeded)) {
        if ((int)B3M(capMod) == 0) b3_charges_cm0(c, mrow, prw, t, &w);
        else b3_charges(c, mrow, prw, t, &w);
    }

    if (ChargeComputationNeeded) {
        const double weff = B3P(weff);
        const double pd = B3I(drainPerimeter), ps = B3I(sourcePerimeter);
        const double czbd = B3M(unitAreaTempJctCap) * B3I(drainArea);
        const double czbs = B3M(unitAreaTempJctCap) * B3I(sourceArea);
        double czbdsw, czbdswg, czbssw, czbsswg, q;
        if (pd < weff) { czbdswg = B3M(unitLengthGateSidewallTempJctCap) * pd; czbdsw = 0.0; }
        else { czbdsw = B3M(unitLengthSidewallTempJctCap) * (pd - weff); czbdswg = B3M(unitLengthGateSidewallTempJctCap) * weff; }
        if (ps < weff) { czbssw = 0.0; czbsswg = B3M(unitLengthGateSidewallTempJctCap) * ps; }
        else { czbssw = B3M(unitLengthSidewallTempJctCap) * (ps - weff); czbsswg = B3M(unitLengthGateSidewallTempJctCap) * weff; }
        b3_junction_cv(vbs, czbs, czbssw, czbsswg, mrow, &q, &capbs);
        B3ST(0, B3ST_qbs) = q;
        b3_junction_cv(vbd, czbd, czbdsw, czbdswg, mrow, &q, &capbd);
        B3ST(0, B3ST_qbd) = q;
    }

    /* convergence flag (the cdhat test is compiled out: NEWCONV, macros.h:19) */
    if (!off || !(mode & NGB_MODEINITFIX)) {
        if (Check == 1) {
#ifdef __CUDA_ARCH__
            atomicAdd(&c->ctl.noncon[s], 1);
#else
            c->ctl.noncon[s] += 1;
#endif
        }
    }
    B3ST(0, B3ST_vbs) = vbs;
    B3ST(0, B3ST_vbd) = vbd;
    B3ST(0, B3ST_vgs) = vgs;
    B3ST(0, B3ST_vds) = vds;
    B3ST(0, B3ST_qdef) = qdef;

    {
        double gcdgb = 0.0, gcddb = 0.0, gcdsb = 0.0, gcsgb = 0.0, gcsdb = 0.0, gcssb = 0.0;
        double gcggb = 0.0, gcgdb = 0.0, gcgsb = 0.0, gcbgb = 0.0, gcbdb = 0.0, gcbsb = 0.0;
        double ceqqg = 0.0, ceqqb = 0.0, ceqqd = 0.0;
        int integrate = 0;

        if (ChargeComputationNeeded) {
            const double ag0 = NGB_LDG(&c->ctl.ag0[s]);
            const double cgbo = B3P(cgbo);
            double cgdo, qgdo, cgso, qgso, qgate = w.qgate, qbulk = w.qbulk, qdrn = w.qdrn, qgd, qgs, qgb;
            {
                const int capMod = (int)B3M(capMod);
                if (capMod == 0) {                               /* b3ld.c:2497-2503 */
                    cgdo = B3P(cgdo); qgdo = B3P(cgdo) * vgd;
                    cgso = B3P(cgso); qgso = B3P(cgso) * vgs;
                } else if (capMod == 1) {                        /* b3ld.c:2504-2532 */
                    b3_overlap_cm1(vgd, B3P(cgdo), B3P(weffCV), B3P(cgdl), B3P(ckappa), &cgdo, &qgdo);
                    b3_overlap_cm1(vgs, B3P(cgso), B3P(weffCV), B3P(cgsl), B3P(ckappa), &cgso, &qgso);
                } else {
                    b3_overlap(vgd, B3P(cgdo), B3P(weffCV) * B3P(cgdl), B3P(ckappa), &cgdo, &qgdo);
                    b3_overlap(vgs, B3P(cgso), B3P(weffCV) * B3P(cgsl), B3P(ckappa), &cgso, &qgso);
                }
            }
            if (b3mode > 0) {
                gcggb = (w.cggb + cgdo + cgso + cgbo) * ag0;
                gcgdb = (w.cgdb - cgdo) * ag0;
                gcgsb = (w.cgsb - cgso) * ag0;
                gcdgb = (w.cdgb - cgdo) * ag0;
                gcddb = (w.cddb + capbd + cgdo) * ag0;
                gcdsb = w.cdsb * ag0;
                gcsgb = -(w.cggb + w.cbgb + w.cdgb + cgso) * ag0;
                gcsdb = -(w.cgdb + w.cbdb + w.cddb) * ag0;
                gcssb = (capbs + cgso - (w.cgsb + w.cbsb + w.cdsb)) * ag0;
                gcbgb = (w.cbgb - cgbo) * ag0;
                gcbdb = (w.cbdb - capbd) * ag0;
                gcbsb = (w.cbsb - capbs) * ag0;
                qgd = qgdo; qgs = qgso; qgb = cgbo * vgb;
                qgate += qgd + qgs + qgb;
                qbulk -= qgb;
                qdrn -= qgd;
            } else {
                double qsrc;
                gcggb = (w.cggb + cgdo + cgso + cgbo) * ag0;
                gcgdb = (w.cgsb - cgdo) * ag0;
                gcgsb = (w.cgdb - cgso) * ag0;
                gcdgb = -(w.cggb + w.cbgb + w.cdgb + cgdo) * ag0;
                gcddb = (capbd + cgdo - (w.cgsb + w.cbsb + w.cdsb)) * ag0;
                gcdsb = -(w.cgdb + w.cbdb + w.cddb) * ag0;
                gcsgb = (w.cdgb - cgso) * ag0;
                gcsdb = w.cdsb * ag0;
                gcssb = (w.cddb + capbs + cgso) * ag0;
                gcbgb = (w.cbgb - cgbo) * ag0;
                gcbdb = (w.cbsb - capbd) * ag0;
                gcbsb = (w.cbdb - capbs) * ag0;
                qgd = qgdo; qgs = qgso; qgb = cgbo * vgb;
                qgate += qgd + qgs + qgb;
                qbulk -= qgb;
                qsrc = qdrn - qgs;
                qdrn = -(qgate + qbulk + qsrc);
            }
            B3ST(0, B3ST_qg) = qgate;
            B3ST(0, B3ST_qd) = qdrn - B3ST(0, B3ST_qbd);
            B3ST(0, B3ST_qb) = qbulk + B3ST(0, B3ST_qbd) + B3ST(0, B3ST_qbs);

            if (mode & NGB_MODEINITSMSIG) return NGB_OK;            /* line1000 */
            if (!(mode & NGB_MODEDCTRANCURVE)) integrate = 1;
        }
        if (integrate) {
            const int order = NGB_LDG(&c->ctl.order[s]);
            const double ag0 = NGB_LDG(&c->ctl.ag0[s]), ag1 = NGB_LDG(&c->ctl.ag1[s]);
            static const int qk[3] = { B3ST_qb, B3ST_qg, B3ST_qd };
            double cq[3];
            if (order != 1 && order != 2) return NGB_E_ORDER;
            if (mode & NGB_MODEINITTRAN) {
                B3ST(1, B3ST_qb) = B3ST(0, B3ST_qb);
                B3ST(1, B3ST_qg) = B3ST(0, B3ST_qg);
                B3ST(1, B3ST_qd) = B3ST(0, B3ST_qd);
            }
            for (int k = 0; k < 3; k++) {
                const int q = qk[k];
                cq[k] = ngb_integrate(c->ctl.gear, order, ag0, ag1, c->ctl.gear ? NGB_LDG(&c->ctl.ag2[s]) : 0.0, B3ST(0, q), B3ST(1, q),
                                      (c->ctl.gear && order == 2) ? B3ST(2, q) : 0.0, (order == 2) ? B3ST(1, q + 1) : 0.0);
                B3ST(0, q + 1) = cq[k];
                if (c->ctl.lte)
                    ngb_lte_state(&c->ctl, s, c->state, B3ST_COUNT, (size_t)c->T, t, head, q, order);
            }
            ceqqg = cq[1] - gcggb * vgb + gcgdb * vbd + gcgsb * vbs;
            ceqqb = cq[0] - gcbgb * vgb + gcbdb * vbd + gcbsb * vbs;
            ceqqd = cq[2] - gcdgb * vgb + gcddb * vbd + gcdsb * vbs;
            if (mode & NGB_MODEINITTRAN) {
                B3ST(1, B3ST_cqb) = B3ST(0, B3ST_cqb);
                B3ST(1, B3ST_cqg) = B3ST(0, B3ST_cqg);
                B3ST(1, B3ST_cqd) = B3ST(0, B3ST_cqd);
            }
        } else {
            /* line850: no charge currents and no capacitive conductances */
            gcdgb = gcddb = gcdsb = 0.0;
            gcsgb = gcsdb = gcssb = 0.0;
            gcggb = gcgdb = gcgsb = 0.0;
            gcbgb = gcbdb = gcbsb = 0.0;
        }

        /* line900: equivalent currents and stamps */
        {
            double Gm, Gmbs, FwdSum, RevSum, cdreq, ceqbd, ceqbs;
            double gbbdp, gbbsp, gbdpg, gbdpdp, gbdpb, gbdpsp, gbspg, gbspdp, gbspb, gbspsp;
            const double m = B3I(m);
            const double gdpr = B3I(drainConductance), gspr = B3I(sourceConductance);
            if (b3mode >= 0) {
                Gm = w.gm; Gmbs = w.gmbs;
                FwdSum = Gm + Gmbs; RevSum = 0.0;
                cdreq = type * (w.cdrain - w.gds * vds - Gm * vgs - Gmbs * vbs);
                ceqbd = -type * (w.csub - w.gbds * vds - w.gbgs * vgs - w.gbbs * vbs);
                ceqbs = 0.0;
                gbbdp = -w.gbds;
                gbbsp = (w.gbds + w.gbgs + w.gbbs);
                gbdpg = w.gbgs; gbdpdp = w.gbds; gbdpb = w.gbbs;
                gbdpsp = -(gbdpg + gbdpdp + gbdpb);
                gbspg = 0.0; gbspdp = 0.0; gbspb = 0.0; gbspsp = 0.0;
            } else {
                Gm = -w.gm; Gmbs = -w.gmbs;
                FwdSum = 0.0; RevSum = -(Gm + Gmbs);
                cdreq = -type * (w.cdrain + w.gds * vds + Gm * vgd + Gmbs * vbd);
                ceqbs = -type * (w.csub + w.gbds * vds - w.gbgs * vgd - w.gbbs * vbd);
                ceqbd = 0.0;
                gbbsp = -w.gbds;
                gbbdp = (w.gbds + w.gbgs + w.gbbs);
                gbdpg = 0.0; gbdpsp = 0.0; gbdpb = 0.0; gbdpdp = 0.0;
                gbspg = w.gbgs; gbspsp = w.gbds; gbspb = w.gbbs;
                gbspdp = -(gbspg + gbspsp + gbspb);
            }
            if (type > 0) {
                ceqbs += (cbs - gbs * vbs);
                ceqbd += (cbd - gbd * vbd);
            } else {
                ceqbs -= (cbs - gbs * vbs);
                ceqbd -= (cbd - gbd * vbd);
                ceqqg = -ceqqg; ceqqb = -ceqqb; ceqqd = -ceqqd;
            }
#define B3_STAMP(k, v) do { const int r_ = NGB_LDG(&c->spos[(k) * c->ninst + inst]); if (r_ >= 0) c->stamp[(size_t)r_ * S + s] = (v); } while (0)
            /* rhs: `-=` statements are stored negated so that assembly only adds */
            B3_STAMP(B3S_rG, -(m * ceqqg));
            B3_STAMP(B3S_rB, -(m * (ceqbs + ceqbd + ceqqb)));
            B3_STAMP(B3S_rDP, m * (ceqbd - cdreq - ceqqd));
            B3_STAMP(B3S_rSP, m * (cdreq + ceqbs + ceqqg + ceqqb + ceqqd));
            /* matrix; without NQS the terms dxpart*ggt*, T1*ddxpart_* and ggt* of the reference are exact
             * zeros, and x + 0.0 == x, so they are left out */
            B3_STAMP(B3S_Dd, m * gdpr);
            B3_STAMP(B3S_Gg, m * gcggb);
            B3_STAMP(B3S_Ss, m * gspr);
            B3_STAMP(B3S_Bb, m * (gbd + gbs - gcbgb - gcbdb - gcbsb - w.gbbs));
            B3_STAMP(B3S_DPdp, m * (gdpr + w.gds + gbd + RevSum + gcddb + gbdpdp));
            B3_STAMP(B3S_SPsp, m * (gspr + w.gds + gbs + FwdSum + gcssb + gbspsp));
            B3_STAMP(B3S_Ddp, -(m * gdpr));
            B3_STAMP(B3S_Gb, -(m * (gcggb + gcgdb + gcgsb)));
            B3_STAMP(B3S_Gdp, m * gcgdb);
            B3_STAMP(B3S_Gsp, m * gcgsb);
            B3_STAMP(B3S_Ssp, -(m * gspr));
            B3_STAMP(B3S_Bg, m * (gcbgb - w.gbgs));
            B3_STAMP(B3S_Bdp, m * (gcbdb - gbd + gbbdp));
            B3_STAMP(B3S_Bsp, m * (gcbsb - gbs + gbbsp));
            B3_STAMP(B3S_DPd, -(m * gdpr));
            B3_STAMP(B3S_DPg, m * (Gm + gcdgb + gbdpg));
            B3_STAMP(B3S_DPb, -(m * (gbd - Gmbs + gcdgb + gcddb + gcdsb - gbdpb)));
            B3_STAMP(B3S_DPsp, -(m * (w.gds + FwdSum - gcdsb - gbdpsp)));
            B3_STAMP(B3S_SPg, m * (gcsgb - Gm + gbspg));
            B3_STAMP(B3S_SPs, -(m * gspr));
            B3_STAMP(B3S_SPb, -(m * (gbs + Gmbs + gcsgb + gcsdb + gcssb - gbspb)));
            B3_STAMP(B3S_SPdp, -(m * (w.gds + RevSum - gcsdb - gbspdp)));
#undef B3_STAMP
        }
    }
#undef B3ST
    return NGB_OK;
}
#undef B3M
#undef B3P
#undef B3I
#endif
