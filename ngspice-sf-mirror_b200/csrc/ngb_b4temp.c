/* ngb_b4temp.c -- BSIM4temp inside the library: the model / bin / instance tables the
 * BSIM4 load reads, computed from a model card and instance geometry for ANY parameter value, so that a Monte-Carlo
 * batch can draw model parameters (oxide thickness, doping, ...) continuously instead of picking among tables the
 * reference was run on beforehand.
 *
 * What it restates: BSIM4temp (src/spicelib/devices/bsim4/b4temp.c:69-2408) with the value clamps of BSIM4checkModel
 * (b4check.c:356-366, 518-596, 711-725) and the geometry helpers BSIM4NumFingerDiff / PAeffGeo / RdseffGeo / RdsEndIso /
 * RdsEndSha (b4geo.c:29-390), operation for operation: the results are the same bits as the reference's (tests/
 * test_bsim4_temp.py compares every table entry of every BSIM4 fixture).
 *
 * How it is organised (not the reference's one 2 300-line function over three linked structures):
 *   - three flat double tables indexed by the generated name lists of bsim4_temp_fields.h: model cards [nmodel][B4TM],
 *     size-dependent parameter sets [nsize][B4TS] (one per distinct (model, l, w, nf), found by a linear scan like the
 *     reference's knot list), instances [ninst][B4TI];
 *   - the 160 binned parameters are one table-driven loop (b4t_bin), not 160 statements;
 *   - model level, size level and instance level are separate functions (b4t_model, b4t_size, b4t_instance); the
 *     source / drain junction halves of the instance level are one function called twice (b4t_junction).
 * Warnings the reference prints while clamping are dropped (the clamps themselves are kept); conditions it treats as
 * fatal return NGB_E_PANIC with the message in ngbLastError.
 * This is host code on purpose: it runs once per parameter draw, not per Newton iteration. */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "ngb_host.h"
#include "bsim4_temp_fields.h"
#include "../../include/ngb200.h"

#define T_KBOQ 8.617087e-5
#define T_EPS0 8.85418e-12
#define T_EPSSI 1.03594e-10
#define T_PI 3.141592654
#define T_MAX_EXP 5.834617425e14
#define T_MIN_EXP 1.713908431e-15
#define T_EXP_THRESHOLD 34.0
#define T_Q 1.60219e-19
#define T_CHARGE 1.6021766208e-19
#define T_DELTA 1.0E-9
#define T_SQRT2 1.4142135623730950488016887242097

#define M(f) mdl[B4TM_##f]
#define P(f) sz[B4TS_##f]
#define H(f) in[B4TI_##f]

static const char *const model_names[] = {
#define X(n) #n,
    NGB_B4T_MODEL_FIELDS(X)
#undef X
};
static const char *const size_names[] = {
#define X(n) #n,
    NGB_B4T_SIZE_FIELDS(X)
#undef X
};
static const char *const inst_names[] = {
#define X(n) #n,
    NGB_B4T_INST_FIELDS(X)
#undef X
};
/* binned parameter: its slot in the size table and the four slots of the model table */
static const struct { int s, m0, ml, mw, mp; } binned[] = {
#define X(n) { B4TS_##n, B4TM_##n, B4TM_l##n, B4TM_w##n, B4TM_p##n },
    NGB_B4T_BINNED(X)
#undef X
};

/* layout[0..3] = number of model, size, instance fields and of binned parameters */
void ngbBsim4TempLayout(int layout[4])
{
    layout[0] = B4TM_COUNT; layout[1] = B4TS_COUNT; layout[2] = B4TI_COUNT; layout[3] = (int)(sizeof binned / sizeof binned[0]);
}
const char *ngbBsim4TempFieldName(int list, int i)
{
    if (list == 0 && i >= 0 && i < B4TM_COUNT) return model_names[i];
    if (list == 1 && i >= 0 && i < B4TS_COUNT) return size_names[i];
    if (list == 2 && i >= 0 && i < B4TI_COUNT) return inst_names[i];
    return NULL;
}

static double b4t_dexp(double a)          /* DEXP, b4temp.c:42-50 */
{
    if (a > T_EXP_THRESHOLD) return T_MAX_EXP * (1.0 + (a) - T_EXP_THRESHOLD);
    if (a < -T_EXP_THRESHOLD) return T_MIN_EXP;
    return exp(a);
}
/* e^x / ((e^x - 1)^2 + 2 e^x MIN_EXP), saturating: the short-channel roll-off factor that appears six times */
static double b4t_theta(double x)
{
    if (x < T_EXP_THRESHOLD) {
        const double e = exp(x), d = e - 1.0, d2 = d * d;
        return e / (d2 + 2.0 * e * T_MIN_EXP);
    }
    return 1.0 / (T_MAX_EXP - 2.0);
}

/* what does not depend on the circuit's devices: derived once per run of a model card */
typedef struct { double Temp, Tnom, TRatio, delTemp, Vtm0, Eg0, ni, epsrox, toxe, epssub, vt0; } B4TEnv;

/* ------------------------------------------------------------------ model level, b4temp.c:95-416 */
static void b4t_model(double *mdl, B4TEnv *e)
{
    const double Temp = e->Temp;
    double T0, T1, T2, T3, Eg;
    if (M(SbulkJctPotential) < 0.1) M(SbulkJctPotential) = 0.1;
    if (M(SsidewallJctPotential) < 0.1) M(SsidewallJctPotential) = 0.1;
    if (M(SGatesidewallJctPotential) < 0.1) M(SGatesidewallJctPotential) = 0.1;
    if (M(DbulkJctPotential) < 0.1) M(DbulkJctPotential) = 0.1;
    if (M(DsidewallJctPotential) < 0.1) M(DsidewallJctPotential) = 0.1;
    if (M(DGatesidewallJctPotential) < 0.1) M(DGatesidewallJctPotential) = 0.1;

    /* electrical / physical oxide thickness bookkeeping */
    if (M(mtrlMod) == 0) {
        double tol = fabs(M(toxe));
        if (fabs(M(toxp)) > tol) tol = fabs(M(toxp));
        if (fabs(M(dtox)) > tol) tol = fabs(M(dtox));
        tol = tol * 1e-14;
        if (M(toxeGiven) && M(toxpGiven) && M(dtoxGiven) && (fabs(M(toxe) - (M(toxp) + M(dtox))) > tol)) { /* dtox ignored */ }
        else if (M(toxeGiven) && !M(toxpGiven)) M(toxp) = M(toxe) - M(dtox);
        else if (!M(toxeGiven) && M(toxpGiven)) {
            M(toxe) = M(toxp) + M(dtox);
            if (!M(toxmGiven)) M(toxm) = M(toxe);
        }
    } else if (M(mtrlCompatMod) != 0) {
        T0 = M(epsrox) / 3.9;
        if (M(eotGiven) && M(toxpGiven) && M(dtoxGiven) && (fabs(M(eot) * T0 - (M(toxp) + M(dtox))) > 1.0e-20)) { /* dtox ignored */ }
        else if (M(eotGiven) && !M(toxpGiven)) M(toxp) = T0 * M(eot) - M(dtox);
        else if (!M(eotGiven) && M(toxpGiven)) {
            M(eot) = (M(toxp) + M(dtox)) / T0;
            if (!M(toxmGiven)) M(toxm) = M(eot);
        }
    }
    if (M(mtrlMod)) { e->epsrox = 3.9; e->toxe = M(eot); e->epssub = T_EPS0 * M(epsrsub); }
    else { e->epsrox = M(epsrox); e->toxe = M(toxe); e->epssub = T_EPSSI; }
    {
        const double epsrox = e->epsrox, toxe = e->toxe;
        if (!M(cfGiven)) M(cf) = 2.0 * epsrox * T_EPS0 / T_PI * log(1.0 + 0.4e-6 / toxe);
        M(coxe) = epsrox * T_EPS0 / toxe;
        if (M(mtrlMod) == 0 || M(mtrlCompatMod) != 0) M(coxp) = M(epsrox) * T_EPS0 / M(toxp);
        if (!M(cgdoGiven)) {
            if (M(dlcGiven) && (M(dlc) > 0.0)) M(cgdo) = M(dlc) * M(coxe) - M(cgdl);
            else M(cgdo) = 0.6 * M(xj) * M(coxe);
        }
        if (!M(cgsoGiven)) {
            if (M(dlcGiven) && (M(dlc) > 0.0)) M(cgso) = M(dlc) * M(coxe) - M(cgsl);
            else M(cgso) = 0.6 * M(xj) * M(coxe);
        }
        if (!M(cgboGiven)) M(cgbo) = 2.0 * M(dwc) * M(coxe);

        e->Tnom = M(tnom);
        e->TRatio = Temp / e->Tnom;
        M(vcrit) = e->vt0 * log(e->vt0 / (T_SQRT2 * 1.0e-14));
        M(factor1) = sqrt(e->epssub / (epsrox * T_EPS0) * toxe);
    }
    {
        const double Tnom = e->Tnom;
        double Vtm0, Eg0, ni;
        Vtm0 = M(vtm0) = T_KBOQ * Tnom;
        if (M(mtrlMod) == 0) {
            Eg0 = 1.16 - 7.02e-4 * Tnom * Tnom / (Tnom + 1108.0);
            ni = 1.45e10 * (Tnom / 300.15) * sqrt(Tnom / 300.15) * exp(21.5565981 - Eg0 / (2.0 * Vtm0));
        } else {
            Eg0 = M(bg0sub) - M(tbgasub) * Tnom * Tnom / (Tnom + M(tbgbsub));
            T0 = M(bg0sub) - M(tbgasub) * 90090.0225 / (300.15 + M(tbgbsub));
            ni = M(ni0sub) * (Tnom / 300.15) * sqrt(Tnom / 300.15) * exp((T0 - Eg0) / (2.0 * Vtm0));
        }
        e->Vtm0 = Vtm0; e->Eg0 = Eg0; e->ni = ni;
        M(Eg0) = Eg0;
        M(vtm) = T_KBOQ * Temp;
        if (M(mtrlMod) == 0) Eg = 1.16 - 7.02e-4 * Temp * Temp / (Temp + 1108.0);
        else Eg = M(bg0sub) - M(tbgasub) * Temp * Temp / (Temp + M(tbgbsub));
        /* junction saturation current densities at temperature */
        if (Temp != Tnom) {
            T0 = Eg0 / Vtm0 - Eg / M(vtm);
            T1 = log(Temp / Tnom);
            T2 = T0 + M(SjctTempExponent) * T1;
            T3 = exp(T2 / M(SjctEmissionCoeff));
            M(SjctTempSatCurDensity) = M(SjctSatCurDensity) * T3;
            M(SjctSidewallTempSatCurDensity) = M(SjctSidewallSatCurDensity) * T3;
            M(SjctGateSidewallTempSatCurDensity) = M(SjctGateSidewallSatCurDensity) * T3;
            T2 = T0 + M(DjctTempExponent) * T1;
            T3 = exp(T2 / M(DjctEmissionCoeff));
            M(DjctTempSatCurDensity) = M(DjctSatCurDensity) * T3;
            M(DjctSidewallTempSatCurDensity) = M(DjctSidewallSatCurDensity) * T3;
            M(DjctGateSidewallTempSatCurDensity) = M(DjctGateSidewallSatCurDensity) * T3;
        } else {
            M(SjctTempSatCurDensity) = M(SjctSatCurDensity);
            M(SjctSidewallTempSatCurDensity) = M(SjctSidewallSatCurDensity);
            M(SjctGateSidewallTempSatCurDensity) = M(SjctGateSidewallSatCurDensity);
            M(DjctTempSatCurDensity) = M(DjctSatCurDensity);
            M(DjctSidewallTempSatCurDensity) = M(DjctSidewallSatCurDensity);
            M(DjctGateSidewallTempSatCurDensity) = M(DjctGateSidewallSatCurDensity);
        }
    }
    if (M(SjctTempSatCurDensity) < 0.0) M(SjctTempSatCurDensity) = 0.0;
    if (M(SjctSidewallTempSatCurDensity) < 0.0) M(SjctSidewallTempSatCurDensity) = 0.0;
    if (M(SjctGateSidewallTempSatCurDensity) < 0.0) M(SjctGateSidewallTempSatCurDensity) = 0.0;
    if (M(DjctTempSatCurDensity) < 0.0) M(DjctTempSatCurDensity) = 0.0;
    if (M(DjctSidewallTempSatCurDensity) < 0.0) M(DjctSidewallTempSatCurDensity) = 0.0;
    if (M(DjctGateSidewallTempSatCurDensity) < 0.0) M(DjctGateSidewallTempSatCurDensity) = 0.0;

    /* junction capacitances and built-in potentials at temperature */
    e->delTemp = Temp - M(tnom);
    T0 = M(tcj) * e->delTemp;
    if (T0 >= -1.0) {
        M(SunitAreaTempJctCap) = M(SunitAreaJctCap) * (1.0 + T0);
        M(DunitAreaTempJctCap) = M(DunitAreaJctCap) * (1.0 + T0);
    } else {
        if (M(SunitAreaJctCap) > 0.0) M(SunitAreaTempJctCap) = 0.0;
        if (M(DunitAreaJctCap) > 0.0) M(DunitAreaTempJctCap) = 0.0;
    }
    T0 = M(tcjsw) * e->delTemp;
    if (M(SunitLengthSidewallJctCap) < 0.0) M(SunitLengthSidewallJctCap) = 0.0;
    if (M(DunitLengthSidewallJctCap) < 0.0) M(DunitLengthSidewallJctCap) = 0.0;
    if (T0 >= -1.0) {
        M(SunitLengthSidewallTempJctCap) = M(SunitLengthSidewallJctCap) * (1.0 + T0);
        M(DunitLengthSidewallTempJctCap) = M(DunitLengthSidewallJctCap) * (1.0 + T0);
    } else {
        if (M(SunitLengthSidewallJctCap) > 0.0) M(SunitLengthSidewallTempJctCap) = 0.0;
        if (M(DunitLengthSidewallJctCap) > 0.0) M(DunitLengthSidewallTempJctCap) = 0.0;
    }
    T0 = M(tcjswg) * e->delTemp;
    if (T0 >= -1.0) {
        M(SunitLengthGateSidewallTempJctCap) = M(SunitLengthGateSidewallJctCap) * (1.0 + T0);
        M(DunitLengthGateSidewallTempJctCap) = M(DunitLengthGateSidewallJctCap) * (1.0 + T0);
    } else {
        if (M(SunitLengthGateSidewallJctCap) > 0.0) M(SunitLengthGateSidewallTempJctCap) = 0.0;
        if (M(DunitLengthGateSidewallJctCap) > 0.0) M(DunitLengthGateSidewallTempJctCap) = 0.0;
    }
    M(PhiBS) = M(SbulkJctPotential) - M(tpb) * e->delTemp;
    if (M(PhiBS) < 0.01) M(PhiBS) = 0.01;
    M(PhiBD) = M(DbulkJctPotential) - M(tpb) * e->delTemp;
    if (M(PhiBD) < 0.01) M(PhiBD) = 0.01;
    M(PhiBSWS) = M(SsidewallJctPotential) - M(tpbsw) * e->delTemp;
    if (M(PhiBSWS) < 0.01) M(PhiBSWS) = 0.01;
    M(PhiBSWD) = M(DsidewallJctPotential) - M(tpbsw) * e->delTemp;
    if (M(PhiBSWD) < 0.01) M(PhiBSWD) = 0.01;
    M(PhiBSWGS) = M(SGatesidewallJctPotential) - M(tpbswg) * e->delTemp;
    if (M(PhiBSWGS) < 0.01) M(PhiBSWGS) = 0.01;
    M(PhiBSWGD) = M(DGatesidewallJctPotential) - M(tpbswg) * e->delTemp;
    if (M(PhiBSWGD) < 0.01) M(PhiBSWGD) = 0.01;

    if (M(ijthdfwd) <= 0.0) M(ijthdfwd) = 0.0;
    if (M(ijthsfwd) <= 0.0) M(ijthsfwd) = 0.0;
    if (M(ijthdrev) <= 0.0) M(ijthdrev) = 0.0;
    if (M(ijthsrev) <= 0.0) M(ijthsrev) = 0.0;
    if ((M(xjbvd) <= 0.0) && (M(dioMod) == 2)) M(xjbvd) = 0.0;
    else if ((M(xjbvd) < 0.0) && (M(dioMod) == 0)) M(xjbvd) = 0.0;
    if (M(bvd) <= 0.0) M(bvd) = 0.0;
    if ((M(xjbvs) <= 0.0) && (M(dioMod) == 2)) M(xjbvs) = 0.0;
    else if ((M(xjbvs) < 0.0) && (M(dioMod) == 0)) M(xjbvs) = 0.0;
    if (M(bvs) <= 0.0) M(bvs) = 0.0;
}

/* ------------------------------------------------------------------ size level, b4temp.c:441-1670 */
/* Lnew / Wnew of the size are returned through lw[2] (the instance level needs them again) */
static int b4t_size(double *mdl, const B4TEnv *e, double *sz, double l, double w, double nf, double lw[2])
{
    const double epsrox = e->epsrox, toxe = e->toxe, epssub = e->epssub, Vtm0 = e->Vtm0, ni = e->ni, Eg0 = e->Eg0;
    const double TRatio = e->TRatio, delTemp = e->delTemp, Tnom = e->Tnom;
    double T0, T1, T2, T3, T4, T5, T8, T9, T10, tmp, tmp1, tmp2, tmp3, Inv_L, Inv_W, Inv_LW, PowWeffWr;
    const double Lnew = l + M(xl), Wnew = w / nf + M(xw);
    size_t k;
    lw[0] = Lnew; lw[1] = Wnew;
    P(Length) = l; P(Width) = w; P(NFinger) = nf;

    T0 = pow(Lnew, M(Lln));
    T1 = pow(Wnew, M(Lwn));
    tmp1 = M(Ll) / T0 + M(Lw) / T1 + M(Lwl) / (T0 * T1);
    P(dl) = M(Lint) + tmp1;
    tmp2 = M(Llc) / T0 + M(Lwc) / T1 + M(Lwlc) / (T0 * T1);
    P(dlc) = M(dlc) + tmp2;
    T2 = pow(Lnew, M(Wln));
    T3 = pow(Wnew, M(Wwn));
    tmp1 = M(Wl) / T2 + M(Ww) / T3 + M(Wwl) / (T2 * T3);
    P(dw) = M(Wint) + tmp1;
    tmp2 = M(Wlc) / T2 + M(Wwc) / T3 + M(Wwlc) / (T2 * T3);
    P(dwc) = M(dwc) + tmp2;
    P(dwj) = M(dwj) + tmp2;

    P(leff) = Lnew - 2.0 * P(dl);
    if (P(leff) <= 0.0) { ngb_set_error("BSIM4: effective channel length <= 0"); return NGB_E_PANIC; }
    P(weff) = Wnew - 2.0 * P(dw);
    if (P(weff) <= 0.0) { ngb_set_error("BSIM4: effective channel width <= 0"); return NGB_E_PANIC; }
    P(leffCV) = Lnew - 2.0 * P(dlc);
    if (P(leffCV) <= 0.0) { ngb_set_error("BSIM4: effective channel length for C-V <= 0"); return NGB_E_PANIC; }
    P(weffCV) = Wnew - 2.0 * P(dwc);
    if (P(weffCV) <= 0.0) { ngb_set_error("BSIM4: effective channel width for C-V <= 0"); return NGB_E_PANIC; }
    P(weffCJ) = Wnew - 2.0 * P(dwj);
    if (P(weffCJ) <= 0.0) { ngb_set_error("BSIM4: effective channel width for S/D junctions <= 0"); return NGB_E_PANIC; }

    if (M(binUnit) == 1) {
        Inv_L = 1.0e-6 / P(leff);
        Inv_W = 1.0e-6 / P(weff);
        Inv_LW = 1.0e-12 / (P(leff) * P(weff));
    } else {
        Inv_L = 1.0 / P(leff);
        Inv_W = 1.0 / P(weff);
        Inv_LW = 1.0 / (P(leff) * P(weff));
    }
    /* the 160 binned parameters: x + lx / L + wx / W + px / (L W) */
    for (k = 0; k < sizeof binned / sizeof binned[0]; k++)
        sz[binned[k].s] = mdl[binned[k].m0] + mdl[binned[k].ml] * Inv_L + mdl[binned[k].mw] * Inv_W + mdl[binned[k].mp] * Inv_LW;

    P(abulkCVfactor) = 1.0 + pow((P(clc) / P(leffCV)), P(cle));

    /* temperature dependence of mobility, saturation velocity and the series resistances */
    T0 = (TRatio - 1.0);
    PowWeffWr = pow(P(weffCJ) * 1.0e6, P(wr)) * nf;
    T1 = T2 = T3 = T4 = 0.0;
    P(ucs) = P(ucs) * pow(TRatio, P(ucste));
    if (M(tempMod) == 0) {
        P(ua) = P(ua) + P(ua1) * T0;
        P(ub) = P(ub) + P(ub1) * T0;
        P(uc) = P(uc) + P(uc1) * T0;
        P(ud) = P(ud) + P(ud1) * T0;
        P(vsattemp) = P(vsat) - P(at) * T0;
        T10 = P(prt) * T0;
        if (M(rdsMod)) {
            T1 = P(rdw) + T10; T2 = M(rdwmin) + T10;       /* external Rd(V) */
            T3 = P(rsw) + T10; T4 = M(rswmin) + T10;       /* external Rs(V) */
        }
        P(rds0) = (P(rdsw) + T10) * nf / PowWeffWr;        /* internal Rds(V) */
        P(rdswmin) = (M(rdswmin) + T10) * nf / PowWeffWr;
    } else {
        if (M(tempMod) == 3) {
            P(ua) = P(ua) * pow(TRatio, P(ua1));
            P(ub) = P(ub) * pow(TRatio, P(ub1));
            P(uc) = P(uc) * pow(TRatio, P(uc1));
            P(ud) = P(ud) * pow(TRatio, P(ud1));
        } else {
            P(ua) = P(ua) * (1.0 + P(ua1) * delTemp);
            P(ub) = P(ub) * (1.0 + P(ub1) * delTemp);
            P(uc) = P(uc) * (1.0 + P(uc1) * delTemp);
            P(ud) = P(ud) * (1.0 + P(ud1) * delTemp);
        }
        P(vsattemp) = P(vsat) * (1.0 - P(at) * delTemp);
        T10 = 1.0 + P(prt) * delTemp;
        if (M(rdsMod)) {
            T1 = P(rdw) * T10; T2 = M(rdwmin) * T10;
            T3 = P(rsw) * T10; T4 = M(rswmin) * T10;
        }
        P(rds0) = P(rdsw) * T10 * nf / PowWeffWr;
        P(rdswmin) = M(rdswmin) * T10 * nf / PowWeffWr;
    }
    if (T1 < 0.0) T1 = 0.0;
    if (T2 < 0.0) T2 = 0.0;
    P(rd0) = T1 / PowWeffWr;
    P(rdwmin) = T2 / PowWeffWr;
    if (T3 < 0.0) T3 = 0.0;
    if (T4 < 0.0) T4 = 0.0;
    P(rs0) = T3 / PowWeffWr;
    P(rswmin) = T4 / PowWeffWr;

    if (P(u0) > 1.0) P(u0) = P(u0) / 1.0e4;
    T5 = 1.0 - P(up) * exp(-P(leff) / P(lp));             /* mobility channel-length dependence */
    P(u0temp) = P(u0) * T5 * pow(TRatio, P(ute));
    if (P(eu) < 0.0) P(eu) = 0.0;
    if (P(ucs) < 0.0) P(ucs) = 0.0;

    P(vfbsdoff) = P(vfbsdoff) * (1.0 + P(tvfbsdoff) * delTemp);
    P(voff) = P(voff) * (1.0 + P(tvoff) * delTemp);
    P(nfactor) = P(nfactor) + P(tnfactor) * delTemp / Tnom;
    P(voffcv) = P(voffcv) * (1.0 + P(tvoffcv) * delTemp);
    P(eta0) = P(eta0) + P(teta0) * delTemp / Tnom;

    if (M(vtlGiven) && (M(vtl) > 0.0)) {                   /* source-end velocity limit */
        if (M(lc) < 0.0) P(lc) = 0.0;
        else P(lc) = M(lc);
        T0 = P(leff) / (P(xn) * P(leff) + P(lc));
        P(tfactor) = (1.0 - T0) / (1.0 + T0);
    }
    P(cgdo) = (M(cgdo) + P(cf)) * P(weffCV);
    P(cgso) = (M(cgso) + P(cf)) * P(weffCV);
    P(cgbo) = M(cgbo) * P(leffCV) * nf;

    if (!M(ndepGiven) && M(gamma1Given)) {
        T0 = P(gamma1) * M(coxe);
        P(ndep) = 3.01248e22 * T0 * T0;
    }
    P(phi) = Vtm0 * log(P(ndep) / ni) + P(phin) + 0.4;
    if (P(phi) <= 0.0) { ngb_set_error("BSIM4: Phi = %g is not positive (check Phin and Ndep)", P(phi)); return NGB_E_PANIC; }
    P(sqrtPhi) = sqrt(P(phi));
    P(phis3) = P(sqrtPhi) * P(phi);
    P(Xdep0) = sqrt(2.0 * epssub / (T_Q * P(ndep) * 1.0e6)) * P(sqrtPhi);
    P(sqrtXdep0) = sqrt(P(Xdep0));
    if (M(mtrlMod) == 0) P(litl) = sqrt(3.0 * 3.9 / epsrox * P(xj) * toxe);
    else P(litl) = sqrt(M(epsrsub) / epsrox * P(xj) * toxe);
    P(vbi) = Vtm0 * log(P(nsd) * P(ndep) / (ni * ni));
    if (M(mtrlMod) == 0) {
        if (P(ngate) > 0.0) P(vfbsd) = Vtm0 * log(P(ngate) / P(nsd));
        else P(vfbsd) = 0.0;
    } else {
        T0 = Vtm0 * log(P(nsd) / ni);
        T1 = 0.5 * Eg0;
        if (T0 > T1) T0 = T1;
        T2 = M(easub) + T1 - M(type) * T0;
        P(vfbsd) = M(phig) - T2;
    }
    P(cdep0) = sqrt(T_Q * epssub * P(ndep) * 1.0e6 / 2.0 / P(phi));

    /* gate tunnelling prefactors */
    P(ToxRatio) = exp(P(ntox) * log(M(toxref) / toxe)) / toxe / toxe;
    P(ToxRatioEdge) = exp(P(ntox) * log(M(toxref) / (toxe * P(poxedge)))) / toxe / toxe / P(poxedge) / P(poxedge);
    P(Aechvb) = (M(type) == 1) ? 4.97232e-7 : 3.42537e-7;
    P(Bechvb) = (M(type) == 1) ? 7.45669e11 : 1.16645e12;
    if (!(M(v48intVersion) <= 480)) {
        if (M(dlcig) < 0.0) M(dlcig) = 0.0;
        if (M(dlcigd) < 0.0) M(dlcigd) = 0.0;
    }
    P(AechvbEdgeS) = P(Aechvb) * P(weff) * M(dlcig) * P(ToxRatioEdge);
    P(AechvbEdgeD) = P(Aechvb) * P(weff) * M(dlcigd) * P(ToxRatioEdge);
    P(BechvbEdge) = -P(Bechvb) * toxe * P(poxedge);
    P(Aechvb) *= P(weff) * P(leff) * P(ToxRatio);
    P(Bechvb) *= -toxe;

    P(mstar) = 0.5 + atan(P(minv)) / T_PI;
    P(mstarcv) = 0.5 + atan(P(minvcv)) / T_PI;
    P(voffcbn) = P(voff) + M(voffl) / P(leff);
    P(voffcbncv) = P(voffcv) + M(voffcvl) / P(leff);
    P(ldeb) = sqrt(epssub * Vtm0 / (T_Q * P(ndep) * 1.0e6)) / 3.0;
    P(acde) *= pow((P(ndep) / 2.0e16), -0.25);

    /* body-effect coefficients */
    if (M(k1Given) || M(k2Given)) {
        if (!M(k1Given)) P(k1) = 0.53;
        if (!M(k2Given)) P(k2) = -0.0186;
    } else {
        if (!M(vbxGiven)) P(vbx) = P(phi) - 7.7348e-4 * P(ndep) * P(xt) * P(xt);
        if (P(vbx) > 0.0) P(vbx) = -P(vbx);
        if (P(vbm) > 0.0) P(vbm) = -P(vbm);
        if (!M(gamma1Given)) P(gamma1) = 5.753e-12 * sqrt(P(ndep)) / M(coxe);
        if (!M(gamma2Given)) P(gamma2) = 5.753e-12 * sqrt(P(nsub)) / M(coxe);
        T0 = P(gamma1) - P(gamma2);
        T1 = sqrt(P(phi) - P(vbx)) - P(sqrtPhi);
        T2 = sqrt(P(phi) * (P(phi) - P(vbm))) - P(phi);
        P(k2) = T0 * T1 / (2.0 * T2 + P(vbm));
        P(k1) = P(gamma2) - 2.0 * P(k2) * sqrt(P(phi) - P(vbm));
    }
    if (!M(vfbGiven)) {
        if (M(vth0Given)) P(vfb) = M(type) * P(vth0) - P(phi) - P(k1) * P(sqrtPhi);
        else if (M(mtrlMod) && M(phigGiven) && M(nsubGiven)) {
            T0 = Vtm0 * log(P(nsub) / ni);
            T1 = 0.5 * Eg0;
            if (T0 > T1) T0 = T1;
            T2 = M(easub) + T1 + M(type) * T0;
            P(vfb) = M(phig) - T2;
        } else P(vfb) = -1.0;
    }
    if (!M(vth0Given)) P(vth0) = M(type) * (P(vfb) + P(phi) + P(k1) * P(sqrtPhi));
    P(k1ox) = P(k1) * toxe / M(toxm);

    /* short-channel / narrow-width factors evaluated at zero bias */
    tmp = sqrt(epssub / (epsrox * T_EPS0) * toxe * P(Xdep0));
    P(theta0vb0) = b4t_theta(P(dsub) * P(leff) / tmp);
    T5 = b4t_theta(P(drout) * P(leff) / tmp);
    P(thetaRout) = P(pdibl1) * T5 + P(pdibl2);
    tmp = sqrt(P(Xdep0));
    tmp1 = P(vbi) - P(phi);
    tmp2 = M(factor1) * tmp;
    T8 = b4t_theta(P(dvt1w) * P(weff) * P(leff) / tmp2);
    T0 = P(dvt0w) * T8;
    T8 = T0 * tmp1;
    T9 = b4t_theta(P(dvt1) * P(leff) / tmp2);
    T9 = P(dvt0) * T9 * tmp1;
    T4 = toxe * P(phi) / (P(weff) + P(w0));
    T0 = sqrt(1.0 + P(lpe0) / P(leff));
    T3 = 0.0;
    if ((M(tempMod) == 1) || (M(tempMod) == 0)) T3 = (P(kt1) + P(kt1l) / P(leff)) * (TRatio - 1.0);
    if ((M(tempMod) == 2) || (M(tempMod) == 3)) T3 = -P(kt1) * (TRatio - 1.0);
    T5 = P(k1ox) * (T0 - 1.0) * P(sqrtPhi) + T3;
    P(vfbzbfactor) = -T8 - T9 + P(k3) * T4 + T5 - P(phi) - P(k1) * P(sqrtPhi);

    /* stress effect: reference values of the size */
    {
        double wlod = M(wlod), W_tmp, Inv_saref, Inv_sbref;
        if (M(wlod) < 0.0) wlod = 0.0;
        T0 = pow(Lnew, M(llodku0));
        W_tmp = Wnew + wlod;
        T1 = pow(W_tmp, M(wlodku0));
        tmp1 = M(lku0) / T0 + M(wku0) / T1 + M(pku0) / (T0 * T1);
        P(ku0) = 1.0 + tmp1;
        T0 = pow(Lnew, M(llodvth));
        T1 = pow(W_tmp, M(wlodvth));
        tmp1 = M(lkvth0) / T0 + M(wkvth0) / T1 + M(pkvth0) / (T0 * T1);
        P(kvth0) = 1.0 + tmp1;
        P(kvth0) = sqrt(P(kvth0) * P(kvth0) + T_DELTA);
        T0 = (TRatio - 1.0);
        P(ku0temp) = P(ku0) * (1.0 + M(tku0) * T0) + T_DELTA;
        Inv_saref = 1.0 / (M(saref) + 0.5 * l);
        Inv_sbref = 1.0 / (M(sbref) + 0.5 * l);
        P(inv_od_ref) = Inv_saref + Inv_sbref;
        P(rho_ref) = M(ku0) / P(ku0temp) * P(inv_od_ref);
    }
    if (M(mobMod) == 3) {           /* VgsteffVth of the high-k mobility model: n at zero bias */
        double lt1, Theta0, n0;
        lt1 = M(factor1) * P(sqrtXdep0);
        Theta0 = b4t_theta(P(dvt1) * P(leff) / lt1);
        tmp1 = epssub / P(Xdep0);
        tmp2 = P(nfactor) * tmp1;
        tmp3 = (tmp2 + P(cdsc) * Theta0 + P(cit)) / M(coxe);
        if (tmp3 >= -0.5) n0 = 1.0 + tmp3;
        else { T0 = 1.0 / (3.0 + 8.0 * tmp3); n0 = (1.0 + 3.0 * tmp3) * T0; }
        T0 = n0 * M(vtm);
        T1 = P(voffcbn);
        T2 = T1 / T0;
        if (T2 < -T_EXP_THRESHOLD) { T3 = M(coxe) * T_MIN_EXP / P(cdep0); T4 = P(mstar) + T3 * n0; }
        else if (T2 > T_EXP_THRESHOLD) { T3 = M(coxe) * T_MAX_EXP / P(cdep0); T4 = P(mstar) + T3 * n0; }
        else { T3 = exp(T2) * M(coxe) / P(cdep0); T4 = P(mstar) + T3 * n0; }
        P(VgsteffVth) = T0 * log(2.0) / T4;
    }
    T0 = -P(dvtp3) * log(P(leff));                          /* DITS term of 4.7 */
    T1 = b4t_dexp(T0);
    P(dvtp2factor) = P(dvtp5) + P(dvtp2) * T1;
    return NGB_OK;
}

/* the value clamps of BSIM4checkModel (b4check.c): applied after an instance has used the size's parameters */
static void b4t_check_clamps(double *mdl, double *sz)
{
    if (P(ckappas) < 0.02) P(ckappas) = 0.02;
    if (P(ckappad) < 0.02) P(ckappad) = 0.02;
    if (P(a2) < 0.01) P(a2) = 0.01;
    else if (P(a2) > 1.0) { P(a2) = 1.0; P(a1) = 0.0; }
    if (P(prwg) < 0.0) P(prwg) = 0.0;
    if (P(rdsw) < 0.0) { P(rdsw) = 0.0; P(rds0) = 0.0; }
    if (P(rds0) < 0.0) P(rds0) = 0.0;
    if (P(rdswmin) < 0.0) P(rdswmin) = 0.0;
    if (M(vtlGiven) && (P(vtl) > 0.0)) {
        if (P(xn) < 3.0) P(xn) = 3.0;
        if (M(lc) < 0.0) P(lc) = 0.0;
    }
    if (M(cgdo) < 0.0) M(cgdo) = 0.0;
    if (M(cgso) < 0.0) M(cgso) = 0.0;
    if (M(cgbo) < 0.0) M(cgbo) = 0.0;
}

/* ------------------------------------------------------------------ geometry helpers, b4geo.c */
typedef struct { double intD, endD, intS, endS; } B4TFingers;
static B4TFingers b4t_fingers(double nf, int minSD)          /* how many shared / end diffusions the layout has */
{
    B4TFingers f;
    const int NF = (int)nf;
    if ((NF % 2) != 0) {
        f.endD = f.endS = 1.0;
        f.intD = f.intS = 2.0 * (((nf - 1.0) / 2.0 > 0.0) ? (nf - 1.0) / 2.0 : 0.0);
    } else if (minSD == 1) {
        f.endD = 2.0; f.intD = 2.0 * ((nf / 2.0 - 1.0 > 0.0) ? (nf / 2.0 - 1.0) : 0.0);
        f.endS = 0.0; f.intS = nf;
    } else {
        f.endD = 0.0; f.intD = nf;
        f.endS = 2.0; f.intS = 2.0 * ((nf / 2.0 - 1.0 > 0.0) ? (nf / 2.0 - 1.0) : 0.0);
    }
    return f;
}
/* effective perimeters and areas: out = { Ps, Pd, As, Ad } */
static void b4t_pa_eff(double nf, int geo, int minSD, double Weffcj, double DMCG, double DMCI, double DMDG, double out[4])
{
    B4TFingers f = { 0.0, 0.0, 0.0, 0.0 };
    const double T0 = DMCG + DMCI, T1 = DMCG + DMCG, T2 = DMDG + DMDG;
    const double Piso = T0 + T0 + Weffcj, Psha = T1, Pmer = T2;
    const double Aiso = T0 * Weffcj, Asha = DMCG * Weffcj, Amer = DMDG * Weffcj;
    /* kind of the END diffusion on each side per geoMod 0..8: 0 isolated, 1 shared, 2 merged */
    static const int endS_kind[9] = { 0, 0, 1, 1, 0, 1, 2, 2, 2 }, endD_kind[9] = { 0, 1, 0, 1, 2, 2, 0, 1, 2 };
    if (geo < 9) f = b4t_fingers(nf, minSD);
    if (geo >= 0 && geo <= 8) {
        const int ks = endS_kind[geo], kd = endD_kind[geo];
        const double PsE = ks == 0 ? Piso : (ks == 1 ? Psha : Pmer), AsE = ks == 0 ? Aiso : (ks == 1 ? Asha : Amer);
        const double PdE = kd == 0 ? Piso : (kd == 1 ? Psha : Pmer), AdE = kd == 0 ? Aiso : (kd == 1 ? Asha : Amer);
        if (ks == 1) { out[0] = (f.endS + f.intS) * Psha; out[2] = (f.endS + f.intS) * Asha; }
        else { out[0] = f.endS * PsE + f.intS * Psha; out[2] = f.endS * AsE + f.intS * Asha; }
        if (kd == 1) { out[1] = (f.endD + f.intD) * Psha; out[3] = (f.endD + f.intD) * Asha; }
        else { out[1] = f.endD * PdE + f.intD * Psha; out[3] = f.endD * AdE + f.intD * Asha; }
    } else if (geo == 9) {
        out[0] = Piso + (nf - 1.0) * Psha; out[1] = nf * Psha;
        out[2] = Aiso + (nf - 1.0) * Asha; out[3] = nf * Asha;
    } else if (geo == 10) {
        out[0] = nf * Psha; out[1] = Piso + (nf - 1.0) * Psha;
        out[2] = nf * Asha; out[3] = Aiso + (nf - 1.0) * Asha;
    }
}
/* end resistance of an isolated (iso = 1) or shared end diffusion; the contact style decides between the two formulas */
static double b4t_rds_end(int iso, double Weffcj, double Rsh, double DMCG, double DMCI, double nuEnd, int rgeo, int is_source)
{
    /* rgeoMod values whose END contact is wide (formula A) / point (formula B), source and drain side */
    const int wide = is_source ? (rgeo == 1 || rgeo == 2 || rgeo == 5) : (rgeo == 1 || rgeo == 3 || rgeo == 7);
    const int point = is_source ? (rgeo == 3 || rgeo == 4 || rgeo == 6) : (rgeo == 2 || rgeo == 4 || rgeo == 8);
    if (wide) return (nuEnd == 0.0) ? 0.0 : Rsh * DMCG / (Weffcj * nuEnd);
    if (point) {
        if (iso) {
            if ((nuEnd == 0.0) || ((DMCG + DMCI) == 0.0)) return 0.0;
            return Rsh * Weffcj / (3.0 * nuEnd * (DMCG + DMCI));
        }
        return (nuEnd == 0.0) ? 0.0 : Rsh * Weffcj / (6.0 * nuEnd * DMCG);
    }
    return 0.0;
}
static double b4t_rds_eff(double nf, int geo, int rgeo, int minSD, double Weffcj, double Rsh, double DMCG, double DMCI, double DMDG, int is_source)
{
    double Rint = 0.0, Rend = 0.0;
    B4TFingers f = { 0.0, 0.0, 0.0, 0.0 };
    /* kind of the end diffusion per geoMod 0..8, as in b4t_pa_eff; merged ends are a plain sheet resistance */
    static const int endS_kind[9] = { 0, 0, 1, 1, 0, 1, 2, 2, 2 }, endD_kind[9] = { 0, 1, 0, 1, 2, 2, 0, 1, 2 };
    if (geo < 9) {
        f = b4t_fingers(nf, minSD);
        if (is_source) Rint = (f.intS == 0.0) ? 0.0 : Rsh * DMCG / (Weffcj * f.intS);
        else Rint = (f.intD == 0.0) ? 0.0 : Rsh * DMCG / (Weffcj * f.intD);
    }
    if (geo >= 0 && geo <= 8) {
        const int kind = is_source ? endS_kind[geo] : endD_kind[geo];
        const double nuEnd = is_source ? f.endS : f.endD;
        if (kind == 2) {
            /* geoMod 4, 6, 8 divide by the width alone, 5 and 7 by width x number of ends (b4geo.c:212-236) */
            const int by_count = (geo == 5 && !is_source) || (geo == 7 && is_source);
            Rend = by_count ? Rsh * DMDG / (Weffcj * nuEnd) : Rsh * DMDG / Weffcj;
        } else Rend = b4t_rds_end(kind == 0, Weffcj, Rsh, DMCG, DMCI, nuEnd, rgeo, is_source);
    } else if (geo == 9 || geo == 10) {
        const int wide_side = (geo == 9) ? is_source : !is_source;
        if (wide_side) {
            Rend = 0.5 * Rsh * DMCG / Weffcj;
            Rint = (nf == 2.0) ? 0.0 : Rsh * DMCG / (Weffcj * (nf - 2.0));
        } else {
            Rend = 0.0;
            Rint = Rsh * DMCG / (Weffcj * nf);
        }
    }
    if (Rint <= 0.0) return Rend;
    if (Rend <= 0.0) return Rint;
    return Rint * Rend / (Rint + Rend);
}

/* forward-bias knee of a junction diode, BSIM4DioIjthVjmEval (b4temp.c:52-66) */
static double b4t_vjm(double Nvtm, double Ijth, double Isb, double XExpBV)
{
    const double Tc = XExpBV, Tb = 1.0 + Ijth / Isb - Tc;
    const double EVjmovNv = 0.5 * (Tb + sqrt(Tb * Tb + 4.0 * Tc));
    return Nvtm * log(EVjmovNv);
}
/* one junction (source or drain): limiting voltages / currents / slopes by dioMod, b4temp.c:2107-2232.
 * q = { XExpBV, vjmFwd, vjmRev, IVjmFwd, IVjmRev, slpFwd, slpRev } in / out */
static void b4t_junction(int dioMod, double Nvtm, double Isat, double bv, double xjbv, double ijthfwd, double ijthrev, double q[7])
{
    double T0, T1, T2;
    if (!(Isat > 0.0)) return;
    switch (dioMod) {
    case 0:
        if ((bv / Nvtm) > T_EXP_THRESHOLD) q[0] = xjbv * T_MIN_EXP;
        else q[0] = xjbv * exp(-bv / Nvtm);
        break;
    case 1:
        q[1] = b4t_vjm(Nvtm, ijthfwd, Isat, 0.0);
        q[3] = Isat * exp(q[1] / Nvtm);
        break;
    case 2:
        if ((bv / Nvtm) > T_EXP_THRESHOLD) q[0] = xjbv * T_MIN_EXP;
        else { q[0] = exp(-bv / Nvtm); q[0] *= xjbv; }
        q[1] = b4t_vjm(Nvtm, ijthfwd, Isat, q[0]);
        T0 = exp(q[1] / Nvtm);
        q[3] = Isat * (T0 - q[0] / T0 + q[0] - 1.0);
        q[5] = Isat * (T0 + q[0] / T0) / Nvtm;
        T2 = ijthrev / Isat;
        if (T2 < 1.0) T2 = 10.0;
        q[2] = -bv - Nvtm * log((T2 - 1.0) / xjbv);
        T1 = xjbv * exp(-(bv + q[2]) / Nvtm);
        q[4] = Isat * (1.0 + T1);
        q[6] = -Isat * T1 / Nvtm;
        break;
    default: break;
    }
}

/* ------------------------------------------------------------------ instance level, b4temp.c:1672-2398 */
static int b4t_instance(double *mdl, const B4TEnv *e, double *sz, double *in, double Lnew)
{
    const double epsrox = e->epsrox, toxe = e->toxe, epssub = e->epssub, ni = e->ni, Eg0 = e->Eg0, TRatio = e->TRatio;
    const double nf = H(nf), Ldrn = H(l), Wdrn = H(w) / H(nf);
    double T0, T1, T2, T3, T4, T5, T6, T7, T10, T11, tmp1, tmp2, tmp3;
    int i;

    /* stress effect */
    if ((H(sa) > 0.0) && (H(sb) > 0.0) && ((nf == 1.0) || ((nf > 1.0) && (H(sd) > 0.0)))) {
        double Inv_sa = 0, Inv_sb = 0, kvsat = M(kvsat), Inv_ODeff, rho, OD_offset, dvth0_lod, dk2_lod, deta0_lod;
        if (M(kvsat) < -1.0) kvsat = -1.0;
        if (M(kvsat) > 1.0) kvsat = 1.0;
        for (i = 0; i < nf; i++) {
            T0 = 1.0 / nf / (H(sa) + 0.5 * Ldrn + i * (H(sd) + Ldrn));
            T1 = 1.0 / nf / (H(sb) + 0.5 * Ldrn + i * (H(sd) + Ldrn));
            Inv_sa += T0;
            Inv_sb += T1;
        }
        Inv_ODeff = Inv_sa + Inv_sb;
        rho = M(ku0) / P(ku0temp) * Inv_ODeff;
        T0 = (1.0 + rho) / (1.0 + P(rho_ref));
        H(u0temp) = P(u0temp) * T0;
        T1 = (1.0 + kvsat * rho) / (1.0 + kvsat * P(rho_ref));
        H(vsattemp) = P(vsattemp) * T1;
        OD_offset = Inv_ODeff - P(inv_od_ref);
        dvth0_lod = M(kvth0) / P(kvth0) * OD_offset;
        dk2_lod = M(stk2) / pow(P(kvth0), M(lodk2)) * OD_offset;
        deta0_lod = M(steta0) / pow(P(kvth0), M(lodeta0)) * OD_offset;
        H(vth0) = P(vth0) + dvth0_lod;
        H(eta0) = P(eta0) + deta0_lod;
        H(k2) = P(k2) + dk2_lod;
    } else {
        H(u0temp) = P(u0temp);
        H(vth0) = P(vth0);
        H(vsattemp) = P(vsattemp);
        H(eta0) = P(eta0);
        H(k2) = P(k2);
    }
    /* well proximity effect */
    if (M(wpemod)) {
        double sceff;
        if (!H(scaGiven) && !H(scbGiven) && !H(sccGiven)) {
            if (H(scGiven) && (H(sc) > 0.0)) {
                T1 = H(sc) + Wdrn;
                T2 = 1.0 / M(scref);
                H(sca) = M(scref) * M(scref) / (H(sc) * T1);
                H(scb) = ((0.1 * H(sc) + 0.01 * M(scref)) * exp(-10.0 * H(sc) * T2) - (0.1 * T1 + 0.01 * M(scref)) * exp(-10.0 * T1 * T2)) / Wdrn;
                H(scc) = ((0.05 * H(sc) + 0.0025 * M(scref)) * exp(-20.0 * H(sc) * T2) - (0.05 * T1 + 0.0025 * M(scref)) * exp(-20.0 * T1 * T2)) / Wdrn;
            }
        }
        if (H(sca) < 0.0) H(sca) = 0.0;
        if (H(scb) < 0.0) H(scb) = 0.0;
        if (H(scc) < 0.0) H(scc) = 0.0;
        if (H(sc) < 0.0) H(sc) = 0.0;
        sceff = H(sca) + M(web) * H(scb) + M(wec) * H(scc);
        H(vth0) += P(kvth0we) * sceff;
        H(k2) += P(k2we) * sceff;
        T3 = 1.0 + P(ku0we) * sceff;
        if (T3 <= 0.0) T3 = 0.0;
        H(u0temp) *= T3;
    }
    /* per-instance threshold shift and mobility multiplier (the latter replaces the stress / WPE mobility: reference behaviour) */
    H(vth0) += H(delvto);
    H(vfb) = P(vfb) + M(type) * H(delvto);
    H(u0temp) = P(u0temp) * H(mulu0);

    T3 = M(type) * H(vth0) - H(vfb) - P(phi);
    T4 = T3 + T3;
    T5 = 2.5 * T3;
    H(vtfbphi1) = (M(type) == 1) ? T4 : T5;
    if (H(vtfbphi1) < 0.0) H(vtfbphi1) = 0.0;
    H(vtfbphi2) = 4.0 * T3;
    if (H(vtfbphi2) < 0.0) H(vtfbphi2) = 0.0;
    if (H(k2) < 0.0) {
        T0 = 0.5 * P(k1) / H(k2);
        H(vbsc) = 0.9 * (P(phi) - T0 * T0);
        if (H(vbsc) > -3.0) H(vbsc) = -3.0;
        else if (H(vbsc) < -30.0) H(vbsc) = -30.0;
    } else H(vbsc) = -30.0;
    if (H(vbsc) > P(vbm)) H(vbsc) = P(vbm);
    H(k2ox) = H(k2) * toxe / M(toxm);
    H(vfbzb) = P(vfbzbfactor) + M(type) * H(vth0);
    H(cgso) = P(cgso);
    H(cgdo) = P(cgdo);

    /* substrate resistance network */
    {
        const double lnl = log(P(leff) * 1.0e6), lnw = log(P(weff) * 1.0e6), lnnf = log(nf);
        const double gbmin = M(gbmin);
        int bodymode = 5;
        if (!M(rbps0Given) || !M(rbpd0Given)) bodymode = 1;
        else if ((!M(rbsbx0Given) && !M(rbsby0Given)) || (!M(rbdbx0Given) && !M(rbdby0Given))) bodymode = 3;
#define B4T_G(r) (((r) < 1.0e-3) ? 1.0e3 : gbmin + 1.0 / (r))
        if (H(rbodyMod) == 2) {
            double rx, ry;
            if (bodymode == 5) {
                rx = M(rbsbx0) * exp(M(rbsdbxl) * lnl + M(rbsdbxw) * lnw + M(rbsdbxnf) * lnnf);
                ry = M(rbsby0) * exp(M(rbsdbyl) * lnl + M(rbsdbyw) * lnw + M(rbsdbynf) * lnnf);
                H(rbsb) = rx * ry / (rx + ry);
                rx = M(rbdbx0) * exp(M(rbsdbxl) * lnl + M(rbsdbxw) * lnw + M(rbsdbxnf) * lnnf);
                ry = M(rbdby0) * exp(M(rbsdbyl) * lnl + M(rbsdbyw) * lnw + M(rbsdbynf) * lnnf);
                H(rbdb) = rx * ry / (rx + ry);
            }
            if ((bodymode == 3) || (bodymode == 5)) {
                H(rbps) = M(rbps0) * exp(M(rbpsl) * lnl + M(rbpsw) * lnw + M(rbpsnf) * lnnf);
                H(rbpd) = M(rbpd0) * exp(M(rbpdl) * lnl + M(rbpdw) * lnw + M(rbpdnf) * lnnf);
            }
            rx = M(rbpbx0) * exp(M(rbpbxl) * lnl + M(rbpbxw) * lnw + M(rbpbxnf) * lnnf);
            ry = M(rbpby0) * exp(M(rbpbyl) * lnl + M(rbpbyw) * lnw + M(rbpbynf) * lnnf);
            H(rbpb) = rx * ry / (rx + ry);
        }
        if ((H(rbodyMod) == 1) || ((H(rbodyMod) == 2) && (bodymode == 5))) {
            H(grbdb) = B4T_G(H(rbdb)); H(grbpb) = B4T_G(H(rbpb)); H(grbps) = B4T_G(H(rbps));
            H(grbsb) = B4T_G(H(rbsb)); H(grbpd) = B4T_G(H(rbpd));
        }
        if ((H(rbodyMod) == 2) && (bodymode == 3)) {
            H(grbdb) = H(grbsb) = gbmin;
            H(grbpb) = B4T_G(H(rbpb)); H(grbps) = B4T_G(H(rbps)); H(grbpd) = B4T_G(H(rbpd));
        }
        if ((H(rbodyMod) == 2) && (bodymode == 1)) {
            H(grbdb) = H(grbsb) = gbmin;
            H(grbps) = H(grbpd) = 1.0e3;
            H(grbpb) = B4T_G(H(rbpb));
        }
#undef B4T_G
    }
    /* geometry-dependent parasitics */
    H(grgeltd) = M(rshg) * (H(xgw) + P(weffCJ) / 3.0 / H(ngcon)) / (H(ngcon) * nf * (Lnew - M(xgl)));
    if (H(grgeltd) > 0.0) H(grgeltd) = 1.0 / H(grgeltd);
    else H(grgeltd) = 1.0e3;
    {
        const double DMCGeff = M(dmcg) - M(dmcgt), DMCIeff = M(dmci), DMDGeff = M(dmdg) - M(dmcgt);
        double pa[4];
        if (H(sourcePerimeterGiven)) {
            if (H(sourcePerimeter) == 0.0) H(Pseff) = 0.0;
            else if (H(sourcePerimeter) < 0.0) H(Pseff) = 0.0;
            else if (M(perMod) == 0) H(Pseff) = H(sourcePerimeter);
            else H(Pseff) = H(sourcePerimeter) - P(weffCJ) * nf;
        } else { b4t_pa_eff(nf, (int)H(geoMod), (int)H(min), P(weffCJ), DMCGeff, DMCIeff, DMDGeff, pa); H(Pseff) = pa[0]; }
        if (H(Pseff) < 0.0) H(Pseff) = 0.0;
        if (H(drainPerimeterGiven)) {
            if (H(drainPerimeter) == 0.0) H(Pdeff) = 0.0;
            else if (H(drainPerimeter) < 0.0) H(Pdeff) = 0.0;
            else if (M(perMod) == 0) H(Pdeff) = H(drainPerimeter);
            else H(Pdeff) = H(drainPerimeter) - P(weffCJ) * nf;
        } else { b4t_pa_eff(nf, (int)H(geoMod), (int)H(min), P(weffCJ), DMCGeff, DMCIeff, DMDGeff, pa); H(Pdeff) = pa[1]; }
        if (H(Pdeff) < 0.0) H(Pdeff) = 0.0;
        if (H(sourceAreaGiven)) H(Aseff) = H(sourceArea);
        else { b4t_pa_eff(nf, (int)H(geoMod), (int)H(min), P(weffCJ), DMCGeff, DMCIeff, DMDGeff, pa); H(Aseff) = pa[2]; }
        if (H(Aseff) < 0.0) H(Aseff) = 0.0;
        if (H(drainAreaGiven)) H(Adeff) = H(drainArea);
        else { b4t_pa_eff(nf, (int)H(geoMod), (int)H(min), P(weffCJ), DMCGeff, DMCIeff, DMDGeff, pa); H(Adeff) = pa[3]; }
        if (H(Adeff) < 0.0) H(Adeff) = 0.0;

        /* source / drain series conductances */
        if (H(sNodePrime) != H(sNode)) {
            H(sourceConductance) = 0.0;
            if (H(sourceSquaresGiven)) H(sourceConductance) = M(sheetResistance) * H(sourceSquares);
            else if (H(rgeoMod) > 0) H(sourceConductance) = b4t_rds_eff(nf, (int)H(geoMod), (int)H(rgeoMod), (int)H(min), P(weffCJ), M(sheetResistance), DMCGeff, DMCIeff, DMDGeff, 1);
            else H(sourceConductance) = 0.0;
            if (H(sourceConductance) > 0.0) H(sourceConductance) = 1.0 / H(sourceConductance);
            else H(sourceConductance) = 1.0e3;
        } else H(sourceConductance) = 0.0;
        if (H(dNodePrime) != H(dNode)) {
            H(drainConductance) = 0.0;
            if (H(drainSquaresGiven)) H(drainConductance) = M(sheetResistance) * H(drainSquares);
            else if (H(rgeoMod) > 0) H(drainConductance) = b4t_rds_eff(nf, (int)H(geoMod), (int)H(rgeoMod), (int)H(min), P(weffCJ), M(sheetResistance), DMCGeff, DMCIeff, DMDGeff, 0);
            else H(drainConductance) = 0.0;
            if (H(drainConductance) > 0.0) H(drainConductance) = 1.0 / H(drainConductance);
            else H(drainConductance) = 1.0e3;
        } else H(drainConductance) = 0.0;
    }
    /* junction diodes: saturation currents and the limiting quantities of the selected dioMod */
    {
        const double Nvtms = M(vtm) * M(SjctEmissionCoeff), Nvtmd = M(vtm) * M(DjctEmissionCoeff);
        double Isat, q[7];
        if ((H(Aseff) <= 0.0) && (H(Pseff) <= 0.0)) Isat = 0.0;
        else Isat = H(Aseff) * M(SjctTempSatCurDensity) + H(Pseff) * M(SjctSidewallTempSatCurDensity) + P(weffCJ) * nf * M(SjctGateSidewallTempSatCurDensity);
        q[0] = H(XExpBVS); q[1] = H(vjsmFwd); q[2] = H(vjsmRev); q[3] = H(IVjsmFwd); q[4] = H(IVjsmRev); q[5] = H(SslpFwd); q[6] = H(SslpRev);
        b4t_junction((int)M(dioMod), Nvtms, Isat, M(bvs), M(xjbvs), M(ijthsfwd), M(ijthsrev), q);
        H(XExpBVS) = q[0]; H(vjsmFwd) = q[1]; H(vjsmRev) = q[2]; H(IVjsmFwd) = q[3]; H(IVjsmRev) = q[4]; H(SslpFwd) = q[5]; H(SslpRev) = q[6];
        if ((H(Adeff) <= 0.0) && (H(Pdeff) <= 0.0)) Isat = 0.0;
        else Isat = H(Adeff) * M(DjctTempSatCurDensity) + H(Pdeff) * M(DjctSidewallTempSatCurDensity) + P(weffCJ) * nf * M(DjctGateSidewallTempSatCurDensity);
        q[0] = H(XExpBVD); q[1] = H(vjdmFwd); q[2] = H(vjdmRev); q[3] = H(IVjdmFwd); q[4] = H(IVjdmRev); q[5] = H(DslpFwd); q[6] = H(DslpRev);
        b4t_junction((int)M(dioMod), Nvtmd, Isat, M(bvd), M(xjbvd), M(ijthdfwd), M(ijthdrev), q);
        H(XExpBVD) = q[0]; H(vjdmFwd) = q[1]; H(vjdmRev) = q[2]; H(IVjdmFwd) = q[3]; H(IVjdmRev) = q[4]; H(DslpFwd) = q[5]; H(DslpRev) = q[6];
    }
    /* trap-assisted tunnelling (reverse bias) */
    T0 = (TRatio - 1.0);
    M(njtsstemp) = M(njts) * (1.0 + M(tnjts) * T0);
    M(njtsswstemp) = M(njtssw) * (1.0 + M(tnjtssw) * T0);
    M(njtsswgstemp) = M(njtsswg) * (1.0 + M(tnjtsswg) * T0);
    M(njtsdtemp) = M(njtsd) * (1.0 + M(tnjtsd) * T0);
    M(njtsswdtemp) = M(njtsswd) * (1.0 + M(tnjtsswd) * T0);
    M(njtsswgdtemp) = M(njtsswgd) * (1.0 + M(tnjtsswgd) * T0);
    T7 = Eg0 / M(vtm) * T0;
    T1 = b4t_dexp(M(xtss) * T7);
    T2 = b4t_dexp(M(xtsd) * T7);
    T3 = b4t_dexp(M(xtssws) * T7);
    T4 = b4t_dexp(M(xtsswd) * T7);
    T5 = b4t_dexp(M(xtsswgs) * T7);
    T6 = b4t_dexp(M(xtsswgd) * T7);
    if (M(jtweff) < 0.0) M(jtweff) = 0.0;
    T11 = sqrt(M(jtweff) / P(weffCJ)) + 1.0;
    T10 = P(weffCJ) * nf;
    H(SjctTempRevSatCur) = T1 * H(Aseff) * M(jtss);
    H(DjctTempRevSatCur) = T2 * H(Adeff) * M(jtsd);
    H(SswTempRevSatCur) = T3 * H(Pseff) * M(jtssws);
    H(DswTempRevSatCur) = T4 * H(Pdeff) * M(jtsswd);
    H(SswgTempRevSatCur) = T5 * T10 * T11 * M(jtsswgs);
    H(DswgTempRevSatCur) = T6 * T10 * T11 * M(jtsswgd);

    /* physical oxide thickness from EOT (new material model) */
    if (M(mtrlMod) != 0 && M(mtrlCompatMod) == 0) {
        const double Vtm0eot = T_KBOQ * M(tempeot), Vtmeot = Vtm0eot;
        double vbieot, phieot, vddeot, Vgs_eff, V0, lt1, ltw, Theta0, Delt_vth, TempRatioeot, Vth_NarrowW, Lpe_Vb, Vth, n, Vgsteff;
        double vtfbphi2eot, toxpf, toxpi, Tcen, T8;
        int niter;
        vbieot = Vtm0eot * log(P(nsd) * P(ndep) / (ni * ni));
        phieot = Vtm0eot * log(P(ndep) / ni) + P(phin) + 0.4;
        if (phieot <= 0.0) { ngb_set_error("BSIM4: phieot = %g is not positive (check Phin and Ndep)", phieot); return NGB_E_PANIC; }
        tmp2 = H(vfb) + phieot;
        vddeot = M(type) * M(vddeot);
        T0 = M(epsrgate) * T_EPS0;
        if ((P(ngate) > 1.0e18) && (P(ngate) < 1.0e25) && (vddeot > tmp2) && (T0 != 0)) {
            T1 = 1.0e6 * T_CHARGE * T0 * P(ngate) / (M(coxe) * M(coxe));
            T8 = vddeot - tmp2;
            T4 = sqrt(1.0 + 2.0 * T8 / T1);
            T2 = 2.0 * T8 / (T4 + 1.0);
            T3 = 0.5 * T2 * T2 / T1;
            T7 = 1.12 - T3 - 0.05;
            T6 = sqrt(T7 * T7 + 0.224);
            T5 = 1.12 - 0.5 * (T7 + T6);
            Vgs_eff = vddeot - T5;
        } else Vgs_eff = vddeot;
        V0 = vbieot - phieot;
        lt1 = M(factor1) * P(sqrtXdep0);
        ltw = lt1;
        Theta0 = b4t_theta(P(dvt1) * M(leffeot) / lt1);
        Delt_vth = P(dvt0) * Theta0 * V0;
        T5 = b4t_theta(P(dvt1w) * M(weffeot) * M(leffeot) / ltw);
        T2 = P(dvt0w) * T5 * V0;
        TempRatioeot = M(tempeot) / M(tnom) - 1.0;
        T0 = sqrt(1.0 + P(lpe0) / M(leffeot));
        T1 = P(k1ox) * (T0 - 1.0) * sqrt(phieot) + (P(kt1) + P(kt1l) / M(leffeot)) * TempRatioeot;
        Vth_NarrowW = toxe * phieot / (M(weffeot) + P(w0));
        Lpe_Vb = sqrt(1.0 + P(lpeb) / M(leffeot));
        Vth = M(type) * H(vth0) + (P(k1ox) - P(k1)) * sqrt(phieot) * Lpe_Vb - Delt_vth - T2 + P(k3) * Vth_NarrowW + T1;
        tmp1 = epssub / P(Xdep0);
        tmp2 = P(nfactor) * tmp1;
        tmp3 = (tmp2 + P(cdsc) * Theta0 + P(cit)) / M(coxe);
        if (tmp3 >= -0.5) n = 1.0 + tmp3;
        else { T0 = 1.0 / (3.0 + 8.0 * tmp3); n = (1.0 + 3.0 * tmp3) * T0; }
        if (P(dvtp0) > 0.0) {
            T3 = M(leffeot) + P(dvtp0) * 2.0;
            if (M(tempMod) < 2) T4 = Vtmeot * log(M(leffeot) / T3);
            else T4 = Vtm0eot * log(M(leffeot) / T3);
            Vth -= n * T4;
        }
        Vgsteff = Vgs_eff - Vth;
        T3 = M(type) * H(vth0) - H(vfb) - phieot;
        vtfbphi2eot = 4.0 * T3;
        if (vtfbphi2eot < 0.0) vtfbphi2eot = 0.0;
        niter = 0;
        toxpf = toxe;
        do {
            toxpi = toxpf;
            tmp2 = 2.0e8 * toxpf;
            T0 = (Vgsteff + vtfbphi2eot) / tmp2;
            T1 = 1.0 + exp(M(bdos) * 0.7 * log(T0));
            Tcen = M(ados) * 1.9e-9 / T1;
            toxpf = toxe - epsrox / M(epsrsub) * Tcen;
            niter++;
        } while ((niter <= 4) && (fabs(toxpf - toxpi) > 1e-12));
        H(toxp) = toxpf;
        H(coxp) = epsrox * T_EPS0 / H(toxp);
    } else {
        H(toxp) = M(toxp);
        H(coxp) = M(coxp);
    }
    return NGB_OK;
}

/* ------------------------------------------------------------------ C ABI */
/* BSIM4temp for a set of model cards and instances at circuit temperature `temp` (K):
 *   model [nmodel][B4TM]  in: the cards after BSIM4setup; out: with the model-level derived quantities
 *   inst  [ninst][B4TI]   in: geometry, options, `...Given` flags, node numbers; out: with the instance-level quantities
 *   inst_model [ninst]    which card an instance belongs to
 * and the tables of the BSIM4 load for them (the layout ngbCircuitAddBsim4 / ngbBatchSetBsim4Rows take):
 *   prow [ninst], *nrows, mtab [nrows][B4M_COUNT], ptab [nrows][B4P_COUNT] (room for ninst rows), itab [B4I_COUNT][ninst].
 * One row per distinct (model, l, w, nf) in order of first appearance, like the reference's size-parameter list. */
int ngbBsim4Temp(double temp, double vt0, int nmodel, double *model, int ninst, const int *inst_model, double *inst,
                 int *prow, int *nrows, double *mtab, double *ptab, double *itab)
{
    B4TEnv *env = (B4TEnv *)calloc((size_t)(nmodel > 0 ? nmodel : 1), sizeof(B4TEnv));
    double *sizes = (double *)calloc((size_t)(ninst > 0 ? ninst : 1) * B4TS_COUNT, sizeof(double));
    double (*lwnew)[2] = (double (*)[2])calloc((size_t)(ninst > 0 ? ninst : 1), sizeof(double[2]));
    int *size_model = (int *)calloc((size_t)(ninst > 0 ? ninst : 1), sizeof(int));
    int m, i, r, ns = 0, rc = NGB_OK;
    if (!env || !sizes || !lwnew || !size_model) { rc = NGB_E_PANIC; goto out; }
    for (m = 0; m < nmodel; m++) {
        env[m].Temp = temp; env[m].vt0 = vt0;
        b4t_model(model + (size_t)m * B4TM_COUNT, &env[m]);
    }
    /* instances in model order, then list order (BSIM4temp's loops); the caller's order within a model is kept */
    for (m = 0; m < nmodel && rc == NGB_OK; m++)
        for (i = 0; i < ninst && rc == NGB_OK; i++) {
            double *mdl = model + (size_t)m * B4TM_COUNT, *in = inst + (size_t)i * B4TI_COUNT, *sz;
            if (inst_model[i] != m) continue;
            for (r = 0; r < ns; r++) {
                sz = sizes + (size_t)r * B4TS_COUNT;
                if (size_model[r] == m && H(l) == P(Length) && H(w) == P(Width) && H(nf) == P(NFinger)) break;
            }
            sz = sizes + (size_t)r * B4TS_COUNT;
            if (r == ns) {
                size_model[ns] = m;
                if ((rc = b4t_size(mdl, &env[m], sz, H(l), H(w), H(nf), lwnew[ns]))) break;
                ns++;
            }
            prow[i] = r;
            if ((rc = b4t_instance(mdl, &env[m], sz, in, lwnew[r][0]))) break;
            b4t_check_clamps(mdl, sz);
        }
    if (rc == NGB_OK) {
        /* rows in order of first appearance over the CALLER's instance order (what the flattening of the circuit uses) */
        int *remap = (int *)malloc(sizeof(int) * (size_t)(ns > 0 ? ns : 1)), nr = 0;
        if (!remap) { rc = NGB_E_PANIC; goto out; }
        for (r = 0; r < ns; r++) remap[r] = -1;
        for (i = 0; i < ninst; i++) { if (remap[prow[i]] < 0) remap[prow[i]] = nr++; }
        for (r = 0; r < ns; r++) {
            const double *mdl = model + (size_t)size_model[r] * B4TM_COUNT, *sz = sizes + (size_t)r * B4TS_COUNT;
            int k = 0;
            if (remap[r] < 0) continue;
#define X(nm) mtab[(size_t)remap[r] * B4M_COUNT + (k++)] = M(nm);
            NGB_B4_MODEL_FIELDS(X)
#undef X
            k = 0;
#define X(nm) ptab[(size_t)remap[r] * B4P_COUNT + (k++)] = P(nm);
            NGB_B4_BIN_FIELDS(X)
#undef X
        }
        for (i = 0; i < ninst; i++) {
            const double *in = inst + (size_t)i * B4TI_COUNT;
            int k = 0;
            prow[i] = remap[prow[i]];
#define X(nm) itab[(size_t)(k++) * ninst + i] = H(nm);
            NGB_B4_INST_FIELDS(X)
#undef X
        }
        *nrows = nr;
        free(remap);
    }
out:
    free(env); free(sizes); free(lwnew); free(size_model);
    return rc;
}
