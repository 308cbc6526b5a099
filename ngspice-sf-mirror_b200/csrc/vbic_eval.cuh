/* vbic_eval.cuh -- VBIC load, one thread per (instance, sample).  C++ only (device code and the host
 * build of the kernels); the C host code sees vbic_types.h.
 *
 * VBICload (src/spicelib/devices/vbic/vbicload.c:42-1478) is restated here: branch-voltage selection
 * :174-467, pnjlim on the six junctions :655-672, charge integration :840-961, state stores
 * :963-1034 and the stamps :1076-1267 (vbic_types.h lists them in statement order).
 *
 * The model core -- the reference's vbic_4T_et_cf_fj (:1479-4118) is 2 600 lines of machine-generated
 * code in which every statement is followed by its partial derivatives -- is NOT transcribed.  The
 * VALUE equations are written once, compactly (the five junction charges are one function), over a
 * small forward-mode dual number (VD): the partial derivatives with respect to the nine non-linear
 * branch voltages come out of operator overloading.  Values follow the reference's expressions
 * operation for operation (and exp/log/pow are the glibc-compatible ones), so currents and charges
 * are bit-identical; derivatives agree to rounding.  Self-heating is one more dual variable (the
 * temperature rise Vrth): the temperature mapping of every parameter (:1591-2230) is written over the
 * same dual number, so all the `_Vrth` partials of the generated code come out of the same expressions;
 * excess phase adds the two filter nodes of vbicload.c:703-725.
 */
#ifndef NGB_VBIC_EVAL_CUH
#define NGB_VBIC_EVAL_CUH
#include "vbic_types.h"
#include "devsup.cuh"

/* ---- forward-mode dual number over the nine non-linear branch voltages ---- */
enum { VD_bei, VD_bex, VD_bci, VD_bep, VD_bcp, VD_bcx, VD_rci, VD_rbi, VD_rbp, VD_rth, VD_N };
struct VD { double v; double d[VD_N]; };

NGB_HD VD vd_c(double c) { VD r; r.v = c; for (int i = 0; i < VD_N; i++) r.d[i] = 0.0; return r; }
NGB_HD VD vd_var(double x, int k) { VD r = vd_c(x); r.d[k] = 1.0; return r; }
NGB_HD VD operator+(const VD &a, const VD &b) { VD r; r.v = a.v + b.v; for (int i = 0; i < VD_N; i++) r.d[i] = a.d[i] + b.d[i]; return r; }
NGB_HD VD operator-(const VD &a, const VD &b) { VD r; r.v = a.v - b.v; for (int i = 0; i < VD_N; i++) r.d[i] = a.d[i] - b.d[i]; return r; }
NGB_HD VD operator-(const VD &a) { VD r; r.v = -a.v; for (int i = 0; i < VD_N; i++) r.d[i] = -a.d[i]; return r; }
NGB_HD VD operator*(const VD &a, const VD &b) { VD r; r.v = a.v * b.v; for (int i = 0; i < VD_N; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
NGB_HD VD operator/(const VD &a, const VD &b) { VD r; r.v = a.v / b.v; for (int i = 0; i < VD_N; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) / b.v; return r; }
NGB_HD VD operator+(const VD &a, double b) { VD r = a; r.v = a.v + b; return r; }
NGB_HD VD operator+(double a, const VD &b) { VD r = b; r.v = a + b.v; return r; }
NGB_HD VD operator-(const VD &a, double b) { VD r = a; r.v = a.v - b; return r; }
NGB_HD VD operator-(double a, const VD &b) { VD r; r.v = a - b.v; for (int i = 0; i < VD_N; i++) r.d[i] = -b.d[i]; return r; }
NGB_HD VD operator*(const VD &a, double b) { VD r; r.v = a.v * b; for (int i = 0; i < VD_N; i++) r.d[i] = a.d[i] * b; return r; }
NGB_HD VD operator*(double a, const VD &b) { VD r; r.v = a * b.v; for (int i = 0; i < VD_N; i++) r.d[i] = a * b.d[i]; return r; }
NGB_HD VD operator/(const VD &a, double b) { VD r; r.v = a.v / b; for (int i = 0; i < VD_N; i++) r.d[i] = a.d[i] / b; return r; }
NGB_HD VD operator/(double a, const VD &b) { VD r; r.v = a / b.v; for (int i = 0; i < VD_N; i++) r.d[i] = -r.v * b.d[i] / b.v; return r; }
NGB_HD VD vd_sqrt(const VD &a) { VD r; r.v = sqrt(a.v); for (int i = 0; i < VD_N; i++) r.d[i] = 0.5 * a.d[i] / r.v; return r; }
NGB_HD VD vd_exp(const VD &a) { VD r; r.v = ngb_exp(a.v); for (int i = 0; i < VD_N; i++) r.d[i] = r.v * a.d[i]; return r; }
NGB_HD VD vd_log(const VD &a) { VD r; r.v = ngb_log(a.v); for (int i = 0; i < VD_N; i++) r.d[i] = a.d[i] / a.v; return r; }
NGB_HD VD vd_pow(const VD &a, double e) { VD r; r.v = ngb_pow(a.v, e); const double f = r.v * e / a.v; for (int i = 0; i < VD_N; i++) r.d[i] = f * a.d[i]; return r; }

/* temperature mapping of a saturation current: IS * (rT^x * exp(-ea*(1-rT)/Vtv))^(1/n), :1655-1663 */
NGB_HD VD vb_isat(double is, const VD &rT, const VD &Vtv, double xi, double ea, double n)
{
    const VD x2 = vd_pow(rT, xi);
    const VD x3 = -ea * (1.0 - rT) / Vtv;
    const VD x4 = vd_exp(x3);
    const VD x1 = (x2 * x4);
    const double x5 = (1.0 / n);
    return is * vd_pow(x1, x5);
}
/* built-in potential at temperature, :1919-1977 */
NGB_HD VD vb_psi(double p0, double ea, const VD &rT, const VD &Vtv)
{
    const VD x2 = 0.5 * p0 * rT / Vtv, x3 = vd_exp(x2);
    const VD x4 = -0.5 * p0 * rT / Vtv, x5 = vd_exp(x4);
    const VD x1 = x3 - x5, x6 = vd_log(x1);
    const VD psiio = 2.0 * (Vtv / rT) * x6;
    const VD lg = vd_log(rT);
    const VD psiin = psiio * rT - 3.0 * Vtv * lg - ea * (rT - 1.0);
    const VD y2 = -psiin / Vtv, y3 = vd_exp(y2);
    const VD y1 = 0.5 * (1.0 + vd_sqrt(1.0 + 4.0 * y3)), y4 = vd_log(y1);
    return psiin + 2.0 * Vtv * y4;
}
/* depletion charge function qj(V) with optional smoothing (A > 0) and reach-through (VRT, ART),
 * :2235-2722 (BE, BEX, BC, BEP, BCP instances of the same generated block) */
NGB_HD VD vb_qj(const VD &V, const VD &P, double M, double FC, double A, double VRT, double ART)
{
    const VD dv0 = -P * FC;
    if (A <= 0.0) {
        const VD dvh = V + dv0;
        VD qlo, qhi;
        if (dvh.v > 0.0) {
            const double pwq = ngb_pow((1.0 - FC), (-1.0 - M));
            qlo = P * (1.0 - pwq * (1.0 - FC) * (1.0 - FC)) / (1.0 - M);
            qhi = dvh * (1.0 - FC + 0.5 * M * dvh / P) * pwq;
        } else {
            if ((VRT > 0.0) && (V.v < -VRT)) {
                const VD x3 = vd_pow((1.0 + VRT / P), (1.0 - M));
                qlo = P * (1.0 - x3 * (1.0 - ((1.0 - M) * (V + VRT)) / (P + VRT))) / (1.0 - M);
            } else {
                const VD x3 = vd_pow((1.0 - V / P), (1.0 - M));
                qlo = P * (1.0 - x3) / (1.0 - M);
            }
            qhi = vd_c(0.0);
        }
        return qlo + qhi;
    }
    if ((VRT > 0.0) && (ART > 0.0)) {
        const VD vn0 = (VRT + dv0) / (VRT - dv0);
        const VD vnl0 = 2.0 * vn0 / (vd_sqrt((vn0 - 1.0) * (vn0 - 1.0) + 4.0 * A * A) + vd_sqrt((vn0 + 1.0) * (vn0 + 1.0) + 4.0 * ART * ART));
        const VD vl0 = 0.5 * (vnl0 * (VRT - dv0) - VRT - dv0);
        const VD qlo0 = P * (1.0 - vd_pow((1.0 - vl0 / P), (1.0 - M))) / (1.0 - M);
        const VD vn = (2.0 * V + VRT + dv0) / (VRT - dv0);
        const VD vnl = 2.0 * vn / (vd_sqrt((vn - 1.0) * (vn - 1.0) + 4.0 * A * A) + vd_sqrt((vn + 1.0) * (vn + 1.0) + 4.0 * ART * ART));
        const VD vl = 0.5 * (vnl * (VRT - dv0) - VRT - dv0);
        const VD qlo = P * (1.0 - vd_pow((1.0 - vl / P), (1.0 - M))) / (1.0 - M);
        const VD sel = 0.5 * (vnl + 1.0);
        const VD crt = vd_pow((1.0 + VRT / P), (-M));
        const VD cmx = vd_pow((1.0 + dv0 / P), (-M));
        const VD cl = (1.0 - sel) * crt + sel * cmx;
        const VD ql = (V - vl + vl0) * cl;
        return ql + qlo - qlo0;
    }
    {
        const VD mv0 = vd_sqrt(dv0 * dv0 + 4.0 * A * A);
        const VD vl0 = -0.5 * (dv0 + mv0);
        const VD q0 = -P * vd_pow((1.0 - vl0 / P), (1.0 - M)) / (1.0 - M);
        const VD dv = V + dv0;
        const VD mv = vd_sqrt(dv * dv + 4.0 * A * A);
        const VD vl = 0.5 * (dv - mv) - dv0;
        const VD qlo = -P * vd_pow((1.0 - vl / P), (1.0 - M)) / (1.0 - M);
        const double x3 = ngb_pow((1.0 - FC), (-M));
        return qlo + x3 * (V - vl + vl0) - q0;
    }
}

/* everything VBICload needs back from the model core */
struct VbOut {
    VD Ibe, Ibex, Itzf, Itzr, Ibc, Ibep, Irci, Irbi, Irbp, Qbe, Qbex, Qbc, Qbcx, Qbep, Ibcp, Iccp, Qbcp;
    /* the linear resistors: current (with its d/dVrth) and conductance */
    VD Ircx, Irbx, Ire, Irs;
    double Ircx_Vrcx, Irbx_Vrbx, Ire_Vre, Irs_Vrs, Qbeo, Qbeo_Vbe, Qbco, Qbco_Vbc;
    /* thermal network: dissipated power with its partials (the dual part covers the ten variables; the branch voltages that
     * are not dual variables have their own), thermal resistance current, thermal charge */
    VD Ith;
    double Ith_Vcei, Ith_Vcep, Ith_Vrcx, Ith_Vrbx, Ith_Vre, Ith_Vrs, Irth, Irth_Vrth, Qcth, Qcth_Vrth;
};

/* the VBIC equations at device temperature p[0] + Vrth (Vrth: temperature rise of the thermal node, 0 without
 * self-heating), values as in vbic_4T_et_cf_fj :1591-4118 */
NGB_HD void vbic_core(const double *p, double Vrth_, double Vbei_, double Vbex_, double Vbci_, double Vbep_, double Vbcp_, double Vrcx,
                      double Vbcx_, double Vrci_, double Vrbx, double Vrbi_, double Vre, double Vrbp_, double Vrs,
                      double Vbe, double Vbc, double Vcei, double Vcep, double SCALE, VbOut *o)
{
    const VD Vbei = vd_var(Vbei_, VD_bei), Vbex = vd_var(Vbex_, VD_bex), Vbci = vd_var(Vbci_, VD_bci), Vbep = vd_var(Vbep_, VD_bep),
             Vbcp = vd_var(Vbcp_, VD_bcp), Vbcx = vd_var(Vbcx_, VD_bcx), Vrci = vd_var(Vrci_, VD_rci), Vrbi = vd_var(Vrbi_, VD_rbi),
             Vrbp = vd_var(Vrbp_, VD_rbp);
    const double Tini = 2.731500e+02 + p[0];
    const VD Tdev = (2.731500e+02 + p[0]) + vd_var(Vrth_, VD_rth);
    VD Vtv = 1.380662e-23 * Tdev / 1.602189e-19;
    Vtv.d[VD_rth] = 8.617347e-5;        /* the generated code's dVtv/dT is this rounded literal (:1595), not k/q to full precision */
    const VD rT = Tdev / Tini;
    const VD dT = Tdev - Tini;
    const VD IKFatT = p[53] * vd_pow(rT, p[90]);
    const VD RCXatT = p[1] * vd_pow(rT, p[91]);
    const VD RCIatT = p[2] * vd_pow(rT, p[68]);
    const VD RBXatT = p[6] * vd_pow(rT, p[92]);
    const VD RBIatT = p[7] * vd_pow(rT, p[67]);
    const VD REatT = p[8] * vd_pow(rT, p[66]);
    const VD RSatT = p[9] * vd_pow(rT, p[69]);
    const VD RBPatT = p[10] * vd_pow(rT, p[93]);
    const VD ISatT = vb_isat(p[11], rT, Vtv, p[78], p[71], p[12]);
    const VD ISRRatT = vb_isat(p[94], rT, Vtv, p[95], p[96], p[13]);
    const VD ISPatT = vb_isat(p[42], rT, Vtv, p[78], p[97], p[44]);
    const VD IBEIatT = vb_isat(p[31], rT, Vtv, p[79], p[72], p[33]);
    const VD IBENatT = vb_isat(p[34], rT, Vtv, p[80], p[75], p[35]);
    const VD IBCIatT = vb_isat(p[36], rT, Vtv, p[79], p[73], p[37]);
    const VD IBCNatT = vb_isat(p[38], rT, Vtv, p[80], p[76], p[39]);
    const VD IBEIPatT = vb_isat(p[45], rT, Vtv, p[79], p[73], p[37]);
    const VD IBENPatT = vb_isat(p[46], rT, Vtv, p[80], p[76], p[39]);
    const VD IBCIPatT = vb_isat(p[47], rT, Vtv, p[79], p[74], p[48]);
    const VD IBCNPatT = vb_isat(p[49], rT, Vtv, p[80], p[77], p[50]);
    const VD NFatT = p[12] * (1.0 + dT * p[81]);
    const VD NRatT = p[13] * (1.0 + dT * p[81]);
    const VD AVC2atT = p[41] * (1.0 + dT * p[82]);
    const VD VBBEatT = p[98] * (1.0 + dT * (p[101] + dT * p[102]));
    const VD NBBEatT = p[99] * (1.0 + dT * p[103]);
    const VD PEatT = vb_psi(p[17], p[72], rT, Vtv);
    const VD PCatT = vb_psi(p[24], p[73], rT, Vtv);
    const VD PSatT = vb_psi(p[28], p[74], rT, Vtv);
    const VD CJEatT = p[16] * vd_pow(p[17] / PEatT, p[18]);
    const VD CJCatT = p[21] * vd_pow(p[24] / PCatT, p[25]);
    const VD CJEPatT = p[23] * vd_pow(p[24] / PCatT, p[25]);
    const VD CJCPatT = p[27] * vd_pow(p[28] / PSatT, p[29]);
    const VD GAMMatT = p[4] * vd_pow(rT, p[78]) * vd_exp(-p[71] * (1.0 - rT) / Vtv);
    const VD VOatT = p[3] * vd_pow(rT, p[70]);
    const VD EBBEatT = vd_exp(-VBBEatT / (NBBEatT * Vtv));
    const double IVEF = (p[51] > 0.0) ? 1.0 / p[51] : 0.0;
    const double IVER = (p[52] > 0.0) ? 1.0 / p[52] : 0.0;
    const VD IIKF = (p[53] > 0.0) ? 1.0 / IKFatT : vd_c(0.0);
    const double IIKR = (p[54] > 0.0) ? 1.0 / p[54] : 0.0;
    const double IIKP = (p[55] > 0.0) ? 1.0 / p[55] : 0.0;
    const VD IVO = (p[3] > 0.0) ? 1.0 / VOatT : vd_c(0.0);
    const double IHRCF = (p[5] > 0.0) ? 1.0 / p[5] : 0.0;
    const double IVTF = (p[59] > 0.0) ? 1.0 / p[59] : 0.0;
    const double IITF = (p[60] > 0.0) ? 1.0 / p[60] : 0.0;
    const double slTF = (p[60] > 0.0) ? 0.0 : 1.0;

    /* junction charges */
    const VD qdbe = vb_qj(Vbei, PEatT, p[18], p[14], p[19], 0.0, 0.0);
    const VD qdbex = vb_qj(Vbex, PEatT, p[18], p[14], p[19], 0.0, 0.0);
    const VD qdbc = vb_qj(Vbci, PCatT, p[25], p[14], p[26], p[85], p[86]);
    const VD qdbep = vb_qj(Vbep, PCatT, p[25], p[14], p[26], p[85], p[86]);
    const VD qdbcp = (p[27] > 0.0) ? vb_qj(Vbcp, PSatT, p[29], p[14], p[30], 0.0, 0.0) : vd_c(0.0);

    /* transport current and base charge */
    const VD Ifi = ISatT * (vd_exp(Vbei / (NFatT * Vtv)) - 1.0);
    const VD Iri = ISatT * ISRRatT * (vd_exp(Vbci / (NRatT * Vtv)) - 1.0);
    const VD q1z = 1.0 + qdbe * IVER + qdbc * IVEF;
    const VD q1 = 0.5 * (vd_sqrt((q1z - 1.0e-4) * (q1z - 1.0e-4) + 1.0e-8) + q1z - 1.0e-4) + 1.0e-4;
    const VD q2 = Ifi * IIKF + Iri * IIKR;
    VD qb;
    if (p[88] < 0.5) {
        const VD x3 = vd_pow(q1, 1.0 / p[89]);
        const VD x1 = (x3 + 4.0 * q2);
        qb = 0.5 * (q1 + vd_pow(x1, p[89]));
    } else {
        const VD x1 = (1.0 + 4.0 * q2);
        qb = 0.5 * q1 * (1.0 + vd_pow(x1, p[89]));
    }
    o->Itzr = Iri / qb;
    o->Itzf = Ifi / qb;

    /* parasitic transistor */
    VD Ifp, qbp;
    if (p[42] > 0.0) {
        const VD expi = vd_exp(Vbep / (p[44] * Vtv));
        const VD expx = vd_exp(Vbci / (p[44] * Vtv));
        Ifp = ISPatT * (p[43] * expi + (1.0 - p[43]) * expx - 1.0);
        const VD q2p = Ifp * IIKP;
        qbp = 0.5 * (1.0 + vd_sqrt(1.0 + 4.0 * q2p));
        const VD Irp = ISPatT * (vd_exp(Vbcp / (p[44] * Vtv)) - 1.0);
        o->Iccp = (Ifp - Irp) / qbp;
    } else {
        Ifp = vd_c(0.0);
        qbp = vd_c(1.0);
        o->Iccp = vd_c(0.0);
    }

    /* base-emitter currents, split between the intrinsic (WBE) and the extrinsic junction */
    if (p[32] == 1.0) {
        const VD expi = vd_exp(Vbei / (p[33] * Vtv)), expn = vd_exp(Vbei / (p[35] * Vtv));
        if (p[98] > 0.0) {
            const VD expx = vd_exp((-VBBEatT - Vbei) / (NBBEatT * Vtv));
            o->Ibe = IBEIatT * (expi - 1.0) + IBENatT * (expn - 1.0) - p[100] * (expx - EBBEatT);
        } else {
            o->Ibe = IBEIatT * (expi - 1.0) + IBENatT * (expn - 1.0);
        }
        o->Ibex = vd_c(0.0);
    } else if (p[32] == 0.0) {
        o->Ibe = vd_c(0.0);
        const VD expi = vd_exp(Vbex / (p[33] * Vtv)), expn = vd_exp(Vbex / (p[35] * Vtv));
        if (p[98] > 0.0) {
            const VD expx = vd_exp((-VBBEatT - Vbex) / (NBBEatT * Vtv));
            o->Ibex = IBEIatT * (expi - 1.0) + IBENatT * (expn - 1.0) - p[100] * (expx - EBBEatT);
        } else {
            o->Ibex = IBEIatT * (expi - 1.0) + IBENatT * (expn - 1.0);
        }
    } else {
        {
            const VD expi = vd_exp(Vbei / (p[33] * Vtv)), expn = vd_exp(Vbei / (p[35] * Vtv));
            if (p[98] > 0.0) {
                const VD expx = vd_exp((-VBBEatT - Vbei) / (NBBEatT * Vtv));
                o->Ibe = p[32] * (IBEIatT * (expi - 1.0) + IBENatT * (expn - 1.0) - p[100] * (expx - EBBEatT));
            } else {
                o->Ibe = p[32] * (IBEIatT * (expi - 1.0) + IBENatT * (expn - 1.0));
            }
        }
        {
            const VD expi = vd_exp(Vbex / (p[33] * Vtv)), expn = vd_exp(Vbex / (p[35] * Vtv));
            if (p[98] > 0.0) {
                const VD expx = vd_exp((-VBBEatT - Vbex) / (NBBEatT * Vtv));
                o->Ibex = (1.0 - p[32]) * (IBEIatT * (expi - 1.0) + IBENatT * (expn - 1.0) - p[100] * (expx - EBBEatT));
            } else {
                o->Ibex = (1.0 - p[32]) * (IBEIatT * (expi - 1.0) + IBENatT * (expn - 1.0));
            }
        }
    }

    /* base-collector current with weak avalanche */
    const VD Ibcj = IBCIatT * (vd_exp(Vbci / (p[37] * Vtv)) - 1.0) + IBCNatT * (vd_exp(Vbci / (p[39] * Vtv)) - 1.0);
    if ((p[45] > 0.0) || (p[46] > 0.0))
        o->Ibep = IBEIPatT * (vd_exp(Vbep / (p[37] * Vtv)) - 1.0) + IBENPatT * (vd_exp(Vbep / (p[39] * Vtv)) - 1.0);
    else
        o->Ibep = vd_c(0.0);
    VD Igc;
    if (p[40] > 0.0) {
        const VD vl = 0.5 * (vd_sqrt((PCatT - Vbci) * (PCatT - Vbci) + 0.01) + (PCatT - Vbci));
        const VD x3 = vd_pow(vl, (p[25] - 1.0));
        const VD x4 = vd_exp(-AVC2atT * x3);
        const VD avalf = p[40] * vl * x4;
        Igc = (o->Itzf - o->Itzr - Ibcj) * avalf;
    } else {
        Igc = vd_c(0.0);
    }
    o->Ibc = Ibcj - Igc;

    /* resistors; the intrinsic collector resistance is the quasi-saturation model */
    if (p[1] > 0.0) { o->Ircx = Vrcx / RCXatT; o->Ircx_Vrcx = (1.0 / RCXatT).v; } else { o->Ircx = vd_c(0.0); o->Ircx_Vrcx = 0.0; }
    const VD Kbci = vd_sqrt(1.0 + GAMMatT * vd_exp(Vbci / Vtv));
    const VD Kbcx = vd_sqrt(1.0 + GAMMatT * vd_exp(Vbcx / Vtv));
    if (p[2] > 0.0) {
        const VD rKp1 = (Kbci + 1.0) / (Kbcx + 1.0);
        const VD Iohm = (Vrci + Vtv * (Kbci - Kbcx - vd_log(rKp1))) / RCIatT;
        const VD derf = IVO * RCIatT * Iohm / (1.0 + 0.5 * IVO * IHRCF * vd_sqrt(Vrci * Vrci + 0.01));
        o->Irci = Iohm / vd_sqrt(1.0 + derf * derf);
    } else {
        o->Irci = vd_c(0.0);
    }
    if (p[6] > 0.0) { o->Irbx = Vrbx / RBXatT; o->Irbx_Vrbx = (1.0 / RBXatT).v; } else { o->Irbx = vd_c(0.0); o->Irbx_Vrbx = 0.0; }
    o->Irbi = (p[7] > 0.0) ? Vrbi * qb / RBIatT : vd_c(0.0);
    if (p[8] > 0.0) { o->Ire = Vre / REatT; o->Ire_Vre = (1.0 / REatT).v; } else { o->Ire = vd_c(0.0); o->Ire_Vre = 0.0; }
    o->Irbp = (p[10] > 0.0) ? Vrbp * qbp / RBPatT : vd_c(0.0);
    if ((p[47] > 0.0) || (p[49] > 0.0))
        o->Ibcp = IBCIPatT * (vd_exp(Vbcp / (p[48] * Vtv)) - 1.0) + IBCNPatT * (vd_exp(Vbcp / (p[50] * Vtv)) - 1.0);
    else
        o->Ibcp = vd_c(0.0);
    if (p[9] > 0.0) { o->Irs = Vrs / RSatT; o->Irs_Vrs = (1.0 / RSatT).v; } else { o->Irs = vd_c(0.0); o->Irs_Vrs = 0.0; }

    /* transit time and charges */
    const double sgIf = (Ifi.v > 0.0) ? 1.0 : 0.0;
    const VD rIf = Ifi * sgIf * IITF;
    const VD mIf = rIf / (rIf + 1.0);
    const VD x2 = vd_exp(Vbci * IVTF / 1.44);
    const VD tff = p[56] * (1.0 + p[57] * q1) * (1.0 + p[58] * x2 * (slTF + mIf * mIf) * sgIf);
    o->Qbe = CJEatT * qdbe * p[32] + tff * Ifi / qb;
    o->Qbex = CJEatT * qdbex * (1.0 - p[32]);
    o->Qbc = CJCatT * qdbc + p[61] * Iri + p[22] * Kbci;
    o->Qbcx = p[22] * Kbcx;
    o->Qbep = CJEPatT * qdbep + p[61] * Ifp;
    o->Qbcp = CJCPatT * qdbcp + p[87] * Vbcp;
    o->Qbeo = Vbe * p[15]; o->Qbeo_Vbe = p[15];
    o->Qbco = Vbc * p[20]; o->Qbco_Vbc = p[20];

    /* thermal network (:3929-4010): dissipated power -- every branch current times its branch voltage --, the thermal
     * resistance and the thermal charge.  Vcei, Vcep and the four linear-resistor voltages are not dual variables: the
     * partials with respect to them are written out */
    o->Ith = -(o->Ibe * Vbei + o->Ibc * Vbci + (o->Itzf - o->Itzr) * Vcei + o->Ibex * Vbex + o->Ibep * Vbep + o->Irs * Vrs + o->Ibcp * Vbcp
               + o->Iccp * Vcep + o->Ircx * Vrcx + o->Irci * Vrci + o->Irbx * Vrbx + o->Irbi * Vrbi + o->Ire * Vre + o->Irbp * Vrbp);
    o->Ith_Vcei = o->Itzr.v - o->Itzf.v;
    o->Ith_Vcep = -o->Iccp.v;
    o->Ith_Vrcx = -o->Ircx.v + -Vrcx * o->Ircx_Vrcx;
    o->Ith_Vrbx = -o->Irbx.v + -Vrbx * o->Irbx_Vrbx;
    o->Ith_Vre = -o->Ire.v + -Vre * o->Ire_Vre;
    o->Ith_Vrs = -o->Irs.v + -Vrs * o->Irs_Vrs;
    if (p[83] > 0.0) { o->Irth = Vrth_ / p[83]; o->Irth_Vrth = 1.0 / p[83]; } else { o->Irth = 0.0; o->Irth_Vrth = 0.0; }
    o->Qcth = Vrth_ * p[84]; o->Qcth_Vrth = p[84];

    if (SCALE != 1.0) {
        o->Ibe = SCALE * o->Ibe; o->Ibex = SCALE * o->Ibex; o->Itzf = SCALE * o->Itzf; o->Itzr = SCALE * o->Itzr;
        o->Ibc = SCALE * o->Ibc; o->Ibep = SCALE * o->Ibep; o->Irci = SCALE * o->Irci; o->Irbi = SCALE * o->Irbi;
        o->Irbp = SCALE * o->Irbp; o->Qbe = SCALE * o->Qbe; o->Qbex = SCALE * o->Qbex; o->Qbc = SCALE * o->Qbc;
        o->Qbcx = SCALE * o->Qbcx; o->Qbep = SCALE * o->Qbep; o->Ibcp = SCALE * o->Ibcp; o->Iccp = SCALE * o->Iccp;
        o->Qbcp = SCALE * o->Qbcp;
        o->Ircx = SCALE * o->Ircx; o->Ircx_Vrcx = SCALE * o->Ircx_Vrcx; o->Irbx = SCALE * o->Irbx; o->Irbx_Vrbx = SCALE * o->Irbx_Vrbx;
        o->Ire = SCALE * o->Ire; o->Ire_Vre = SCALE * o->Ire_Vre; o->Irs = SCALE * o->Irs; o->Irs_Vrs = SCALE * o->Irs_Vrs;
        o->Qbeo = SCALE * o->Qbeo; o->Qbeo_Vbe = SCALE * o->Qbeo_Vbe; o->Qbco = SCALE * o->Qbco; o->Qbco_Vbc = SCALE * o->Qbco_Vbc;
        o->Ith = SCALE * o->Ith; o->Ith_Vcei = SCALE * o->Ith_Vcei; o->Ith_Vcep = SCALE * o->Ith_Vcep; o->Ith_Vrcx = SCALE * o->Ith_Vrcx;
        o->Ith_Vrbx = SCALE * o->Ith_Vrbx; o->Ith_Vre = SCALE * o->Ith_Vre; o->Ith_Vrs = SCALE * o->Ith_Vrs;
        o->Irth = SCALE * o->Irth; o->Irth_Vrth = SCALE * o->Irth_Vrth; o->Qcth = SCALE * o->Qcth; o->Qcth_Vrth = SCALE * o->Qcth_Vrth;
    }
}

NGB_HD int vbic_load_thread(const NgbVbicCtx *c, size_t t)
{
    const int S = c->S;
    const int inst = (int)(t / (size_t)S);
    const int s = (int)(t - (size_t)inst * S);
    if (!NGB_LDG(&c->ctl.active[s])) return NGB_OK;
    const int mode = NGB_LDG(&c->ctl.mode[s]);
    const int head = NGB_LDG(&c->ctl.head[s]);
    const int nh = c->ctl.nhist;
    const int off = NGB_LDG(&c->flags[inst]) & VBF_OFF;
    const int selfheat = (NGB_LDG(&c->flags[inst]) & VBF_SELFHEAT) != 0, excess = (NGB_LDG(&c->flags[inst]) & VBF_EXCESS) != 0;
#define VST(h, k) c->state[((size_t)(((head) + (h)) % nh) * VBS_COUNT + (k)) * c->T + t]
#define VAUX(k) NGB_LDG(&c->aux[(size_t)(k) * c->T + t])
    {   /* deferred whole-vector state copies of DCtran (dctran.c:319-322, 711-716) */
        const int sop = NGB_LDG(&c->ctl.stateop[s]);
        if (sop) {
            for (int k = 0; k < VBS_COUNT; k++) {
                if (sop & NGB_OP_COPY01) VST(1, k) = VST(0, k);
                if (sop & NGB_OP_COPY1_23) { const double v = VST(1, k); VST(2, k) = v; if (nh > 3) VST(3, k) = v; }
                if (sop & NGB_OP_COPY23) { const double v = VST(2, k); VST(0, k) = v; if (nh > 3) VST(3, k) = v; }
            }
        }
    }
    double p[VBIC_NP];
    for (int k = 0; k < VBIC_NP; k++) p[k] = NGB_LDG(&c->par[(size_t)k * c->T + t]);
    const double type = VAUX(VBA_type), tVcrit = VAUX(VBA_tVcrit), SCALE = VAUX(VBA_scale);
    const double vt = VAUX(VBA_temp) * (1.38064852e-23 / 1.6021766208e-19);        /* CONSTKoverQ */
    const double gmin = NGB_LDG(&c->ctl.gmin[s]);
    const double *xo = c->x + (size_t)NGB_LDG(&c->ctl.xsel[s]) * c->neq1 * S;
#define XN(role) NGB_LDG(&xo[(size_t)NGB_LDG(&c->nodes[(role) * c->ninst + inst]) * S + s])
    double Vbei, Vbex, Vbci, Vbcx, Vbep, Vrci, Vrbi, Vrbp, Vbcp, Vbe, Vbc, Vrcx, Vrbx, Vre, Vrs;
    double Vrth = 0.0, Vxf1 = 0.0, Vxf2 = 0.0;       /* thermal node and excess-phase filter nodes (ground = 0 when absent) */
    double gbcx = 0.0, cbcx = 0.0, gqbeo = 0.0, gqbco = 0.0, Icth = 0.0, Icth_Vrth = 0.0;
    int icheck = 1;

    if (mode & (NGB_MODEINITSMSIG | NGB_MODEINITTRAN)) {
        const int h = (mode & NGB_MODEINITSMSIG) ? 0 : 1;
        Vbe = type * (XN(VBN_base) - XN(VBN_emit));
        Vbc = type * (XN(VBN_base) - XN(VBN_coll));
        Vbei = VST(h, VBS_vbei); Vbex = VST(h, VBS_vbex); Vbci = VST(h, VBS_vbci); Vbcx = VST(h, VBS_vbcx);
        Vbep = VST(h, VBS_vbep); Vrci = VST(h, VBS_vrci); Vrbi = VST(h, VBS_vrbi); Vrbp = VST(h, VBS_vrbp);
        Vrcx = type * (XN(VBN_coll) - XN(VBN_cx));
        Vrbx = type * (XN(VBN_base) - XN(VBN_bx));
        Vre = type * (XN(VBN_emit) - XN(VBN_ei));
        Vbcp = VST(h, VBS_vbcp);
        Vrs = type * (XN(VBN_subs) - XN(VBN_si));
        if (selfheat) Vrth = VST(h, VBS_vrth);
        Vxf1 = XN(VBN_xf1); Vxf2 = XN(VBN_xf2);
    } else if ((mode & NGB_MODEINITJCT) && (mode & NGB_MODETRANOP) && (mode & NGB_MODEUIC)) {
        Vbe = type * VAUX(VBA_icVBE);
        Vbei = Vbex = Vbe;
        { const double Vce = type * VAUX(VBA_icVCE); Vbc = Vbe - Vce; }
        Vbci = Vbcx = Vbc;
        Vbep = Vbcp = 0.0;
        Vrci = Vrbi = Vrbp = 0.0;
        Vrcx = Vrbx = Vre = Vrs = 0.0;
    } else if ((mode & NGB_MODEINITJCT) && !off) {
        Vbe = Vbei = Vbex = type * tVcrit;
        Vbc = Vbcx = Vbep = 0.0;
        Vbci = -type * tVcrit;
        Vbcp = Vbc - Vbe;
        Vrci = Vrbi = Vrbp = 0.0;
        Vrcx = Vrbx = Vre = Vrs = 0.0;
    } else if ((mode & NGB_MODEINITJCT) || ((mode & NGB_MODEINITFIX) && off)) {
        Vbe = 0.0; Vbei = Vbex = Vbe;
        Vbc = 0.0; Vbci = Vbcx = Vbc;
        Vbep = Vbcp = 0.0;
        Vrci = Vrbi = Vrbp = 0.0;
        Vrcx = Vrbx = Vre = Vrs = 0.0;
    } else {
        if (mode & NGB_MODEINITPRED) {
            const double xfact = NGB_LDG(&c->ctl.delta[s]) / NGB_LDG(&c->ctl.delta_old[(size_t)1 * S + s]);
#define PRED(k) ((1 + xfact) * VST(1, k) - xfact * VST(2, k))
            Vbei = PRED(VBS_vbei); Vbex = PRED(VBS_vbex); Vbci = PRED(VBS_vbci); Vbcx = PRED(VBS_vbcx); Vbep = PRED(VBS_vbep);
            Vrci = PRED(VBS_vrci); Vrbi = PRED(VBS_vrbi); Vrbp = PRED(VBS_vrbp); Vbcp = PRED(VBS_vbcp);
#undef PRED
            {   /* the state0 <- state1 copies of vbicload.c:303-395 (everything the hat currents read) */
                static const unsigned char cp[] = { VBS_vbei, VBS_vbex, VBS_vbci, VBS_vbcx, VBS_vbep, VBS_vrci, VBS_vrbi, VBS_vrbp, VBS_vbcp,
                    VBS_ibe, VBS_ibe_Vbei, VBS_ibex, VBS_ibex_Vbex, VBS_ibc, VBS_ibc_Vbci, VBS_ibc_Vbei, VBS_ibep, VBS_ibep_Vbep,
                    VBS_irci, VBS_irci_Vrci, VBS_irci_Vbci, VBS_irci_Vbcx, VBS_irbi, VBS_irbi_Vrbi, VBS_irbi_Vbei, VBS_irbi_Vbci,
                    VBS_irbp, VBS_irbp_Vrbp, VBS_irbp_Vbep, VBS_irbp_Vbci, VBS_ibcp, VBS_ibcp_Vbcp, VBS_iccp, VBS_iccp_Vbep,
                    VBS_iccp_Vbci, VBS_iccp_Vbcp, VBS_gqbeo, VBS_gqbco, VBS_ircx_Vrcx, VBS_irbx_Vrbx, VBS_irs_Vrs, VBS_ire_Vre,
                    VBS_iciei, VBS_iciei_Vbei, VBS_iciei_Vbci, VBS_iciei_Vxf2 };
                for (unsigned k = 0; k < sizeof cp; k++) VST(0, cp[k]) = VST(1, cp[k]);
                if (selfheat) { VST(0, VBS_vrth) = VST(1, VBS_vrth); VST(0, VBS_qcth) = VST(1, VBS_qcth); }
                if (excess) {
                    static const unsigned char cx2[] = { VBS_vxf1, VBS_qxf1, VBS_cqxf1, VBS_gqxf1, VBS_ixf1_Vbei, VBS_ixf1_Vbci, VBS_ixf1_Vxf2,
                        VBS_ixf1_Vxf1, VBS_vxf2, VBS_qxf2, VBS_cqxf2, VBS_gqxf2, VBS_ixf2_Vxf1, VBS_ixf2_Vxf2 };
                    for (unsigned k = 0; k < sizeof cx2; k++) VST(0, cx2[k]) = VST(1, cx2[k]);
                    if (selfheat) VST(0, VBS_ixf1_Vrth) = VST(1, VBS_ixf1_Vrth);
                }
            }
        } else {
            Vbei = type * (XN(VBN_bi) - XN(VBN_ei));
            Vbex = type * (XN(VBN_bx) - XN(VBN_ei));
            Vbci = type * (XN(VBN_bi) - XN(VBN_ci));
            Vbcx = type * (XN(VBN_bi) - XN(VBN_cx));
            Vbep = type * (XN(VBN_bx) - XN(VBN_bp));
            Vrci = type * (XN(VBN_cx) - XN(VBN_ci));
            Vrbi = type * (XN(VBN_bx) - XN(VBN_bi));
            Vrbp = type * (XN(VBN_bp) - XN(VBN_cx));
            Vbcp = type * (XN(VBN_si) - XN(VBN_bp));
        }
        Vbe = type * (XN(VBN_base) - XN(VBN_emit));
        Vbc = type * (XN(VBN_base) - XN(VBN_coll));
        Vrcx = type * (XN(VBN_coll) - XN(VBN_cx));
        Vrbx = type * (XN(VBN_base) - XN(VBN_bx));
        Vre = type * (XN(VBN_emit) - XN(VBN_ei));
        Vrs = type * (XN(VBN_subs) - XN(VBN_si));
        /* the thermal and filter node voltages are read from the last solution whether or not the junction voltages were
         * predicted (vbicload.c:483-486 follows the MODEINITPRED branch) */
        if (selfheat) Vrth = XN(VBN_temp);
        Vxf1 = XN(VBN_xf1); Vxf2 = XN(VBN_xf2);
        {   /* limit the junction voltages */
            int i1 = 1, i2 = 1, i3 = 1, i4 = 1, i5 = 1, i6 = 0;
            Vbei = ngb_pnjlim(Vbei, VST(0, VBS_vbei), vt, tVcrit, &icheck);
            Vbex = ngb_pnjlim(Vbex, VST(0, VBS_vbex), vt, tVcrit, &i1);
            Vbci = ngb_pnjlim(Vbci, VST(0, VBS_vbci), vt, tVcrit, &i2);
            Vbcx = ngb_pnjlim(Vbcx, VST(0, VBS_vbcx), vt, tVcrit, &i3);
            Vbep = ngb_pnjlim(Vbep, VST(0, VBS_vbep), vt, tVcrit, &i4);
            Vbcp = ngb_pnjlim(Vbcp, VST(0, VBS_vbcp), vt, tVcrit, &i5);
            if (selfheat) {          /* DEVlimitlog (devsup.c:157-184): logarithmic damping of a temperature step beyond 100 K */
                const double told = VST(0, VBS_vrth);
                i6 = 0;
                if (Vrth != Vrth || told != told) { Vrth = 0.0; i6 = 1; }
                if (Vrth > told + 100.0) { Vrth = told + 100.0 + log10((Vrth - told) / 100.0); i6 = 1; }
                else if (Vrth < told - 100.0) { Vrth = told - 100.0 - log10((told - Vrth) / 100.0); i6 = 1; }
            }
            if ((i1 == 1) || (i2 == 1) || (i3 == 1) || (i4 == 1) || (i5 == 1) || (i6 == 1)) icheck = 1;
        }
    }
#undef XN

    const double Vcei = Vbei - Vbci, Vcep = Vbep - Vbcp;
    VbOut o;
    vbic_core(p, Vrth, Vbei, Vbex, Vbci, Vbep, Vbcp, Vrcx, Vbcx, Vrci, Vrbx, Vrbi, Vre, Vrbp, Vrs, Vbe, Vbc, Vcei, Vcep, SCALE, &o);

    /* excess phase (vbicload.c:703-748): the forward transport current passes a second-order filter built from the two
     * internal nodes -- xf1 carries Itzf, the collector current source reads xf2; without it Itzf is used directly */
    double Ixf1, Ixf1_Vxf1, Ixf1_Vxf2, Ixf1_Vbei, Ixf1_Vbci, Ixf1_Vrth, Ixf2, Ixf2_Vxf1, Ixf2_Vxf2, Qxf1 = 0.0, Qxf1_Vxf1 = 0.0, Qxf2 = 0.0, Qxf2_Vxf2 = 0.0;
    double Itxf, Itxf_Vbei, Itxf_Vbci, Itxf_Vrth, Itxf_Vxf2;
    if (excess) {
        Ixf1 = Vxf2 - o.Itzf.v; Ixf1_Vxf1 = 0.0; Ixf1_Vxf2 = 1.0;
        Ixf1_Vbei = -1.0 * o.Itzf.d[VD_bei]; Ixf1_Vbci = -1.0 * o.Itzf.d[VD_bci]; Ixf1_Vrth = -1.0 * o.Itzf.d[VD_rth];
        Ixf2 = Vxf2 - Vxf1; Ixf2_Vxf2 = 1.0; Ixf2_Vxf1 = -1.0;
        Qxf1 = p[62] * Vxf1; Qxf1_Vxf1 = p[62];
        Qxf2 = p[62] * Vxf2 / 3; Qxf2_Vxf2 = p[62] / 3;
        Itxf = Vxf2; Itxf_Vbei = 0.0; Itxf_Vbci = 0.0; Itxf_Vrth = 0.0; Itxf_Vxf2 = 1.0;
    } else {
        Ixf1 = Vxf1; Ixf1_Vxf1 = 1.0; Ixf1_Vxf2 = 0.0; Ixf1_Vbei = 0.0; Ixf1_Vbci = 0.0; Ixf1_Vrth = 0.0;
        Ixf2 = Vxf2; Ixf2_Vxf2 = 1.0; Ixf2_Vxf1 = 0.0;
        Itxf = o.Itzf.v; Itxf_Vbei = o.Itzf.d[VD_bei]; Itxf_Vbci = o.Itzf.d[VD_bci]; Itxf_Vrth = o.Itzf.d[VD_rth]; Itxf_Vxf2 = 0.0;
    }
    double Iciei = Itxf - o.Itzr.v;
    double Iciei_Vbei = Itxf_Vbei - o.Itzr.d[VD_bei];
    double Iciei_Vbci = Itxf_Vbci - o.Itzr.d[VD_bci];
    const double Iciei_Vrth = Itxf_Vrth - o.Itzr.d[VD_rth], Iciei_Vxf2 = Itxf_Vxf2;
    /* d/dVrth of the branch currents (stamped only with self-heating) */
    const double Ibe_Vrth = o.Ibe.d[VD_rth], Ibex_Vrth = o.Ibex.d[VD_rth], Ibc_Vrth = o.Ibc.d[VD_rth], Ibep_Vrth = o.Ibep.d[VD_rth];
    const double Ircx_Vrth = o.Ircx.d[VD_rth], Irci_Vrth = o.Irci.d[VD_rth], Irbx_Vrth = o.Irbx.d[VD_rth], Irbi_Vrth = o.Irbi.d[VD_rth];
    const double Ire_Vrth = o.Ire.d[VD_rth], Irbp_Vrth = o.Irbp.d[VD_rth], Ibcp_Vrth = o.Ibcp.d[VD_rth], Iccp_Vrth = o.Iccp.d[VD_rth];
    const double Irs_Vrth = o.Irs.d[VD_rth];
    double Ibe = o.Ibe.v, Ibe_Vbei = o.Ibe.d[VD_bei];
    double Ibex = o.Ibex.v, Ibex_Vbex = o.Ibex.d[VD_bex];
    double Ibc = o.Ibc.v, Ibc_Vbci = o.Ibc.d[VD_bci];
    const double Ibc_Vbei = o.Ibc.d[VD_bei];
    double Ibep = o.Ibep.v, Ibep_Vbep = o.Ibep.d[VD_bep];
    double Irci = o.Irci.v, Irci_Vrci = o.Irci.d[VD_rci], Irci_Vbci = o.Irci.d[VD_bci], Irci_Vbcx = o.Irci.d[VD_bcx];
    const double Irbi = o.Irbi.v, Irbi_Vrbi = o.Irbi.d[VD_rbi], Irbi_Vbei = o.Irbi.d[VD_bei], Irbi_Vbci = o.Irbi.d[VD_bci];
    const double Irbp = o.Irbp.v, Irbp_Vrbp = o.Irbp.d[VD_rbp], Irbp_Vbep = o.Irbp.d[VD_bep], Irbp_Vbci = o.Irbp.d[VD_bci];
    double Ibcp = o.Ibcp.v, Ibcp_Vbcp = o.Ibcp.d[VD_bcp];
    const double Iccp = o.Iccp.v, Iccp_Vbep = o.Iccp.d[VD_bep], Iccp_Vbci = o.Iccp.d[VD_bci], Iccp_Vbcp = o.Iccp.d[VD_bcp];
    const double Ircx_Vrcx = o.Ircx_Vrcx, Irbx_Vrbx = o.Irbx_Vrbx, Ire_Vre = o.Ire_Vre, Irs_Vrs = o.Irs_Vrs;
    (void)Ixf1_Vrth;

    Ibe += gmin * Vbei;   Ibe_Vbei += gmin;
    Ibex += gmin * Vbex;  Ibex_Vbex += gmin;
    Ibc += gmin * Vbci;   Ibc_Vbci += gmin;
    Ibep += gmin * Vbep;  Ibep_Vbep += gmin;
    Irci += gmin * Vrci;  Irci_Vrci += gmin;
    Irci += gmin * Vbci;  Irci_Vbci += gmin;
    Irci += gmin * Vbcx;  Irci_Vbcx += gmin;
    Ibcp += gmin * Vbcp;  Ibcp_Vbcp += gmin;

    const int order = NGB_LDG(&c->ctl.order[s]);
    const double ag0 = NGB_LDG(&c->ctl.ag0[s]), ag1 = NGB_LDG(&c->ctl.ag1[s]);
    const int gear = c->ctl.gear;
    const double ag2 = gear ? NGB_LDG(&c->ctl.ag2[s]) : 0.0;
#define INTEGRATE(q) ngb_integrate(gear, order, ag0, ag1, ag2, VST(0, q), VST(1, q), (gear && order == 2) ? VST(2, q) : 0.0, (order == 2) ? VST(1, (q) + 1) : 0.0)
    if ((mode & (NGB_MODEDCTRANCURVE | NGB_MODETRAN | NGB_MODEAC)) || ((mode & NGB_MODETRANOP) && (mode & NGB_MODEUIC)) ||
        (mode & NGB_MODEINITSMSIG)) {
        VST(0, VBS_qbe) = o.Qbe.v;   VST(0, VBS_qbex) = o.Qbex.v; VST(0, VBS_qbc) = o.Qbc.v; VST(0, VBS_qbcx) = o.Qbcx.v;
        VST(0, VBS_qbep) = o.Qbep.v; VST(0, VBS_qbeo) = o.Qbeo;   VST(0, VBS_qbco) = o.Qbco; VST(0, VBS_qbcp) = o.Qbcp.v;
        if (selfheat) VST(0, VBS_qcth) = o.Qcth;
        if (excess) { VST(0, VBS_qxf1) = Qxf1; VST(0, VBS_cqxf1) = Qxf1_Vxf1; VST(0, VBS_qxf2) = Qxf2; VST(0, VBS_cqxf2) = Qxf2_Vxf2; }
        if (!(mode & NGB_MODETRANOP) || !(mode & NGB_MODEUIC)) {
            if (mode & NGB_MODEINITSMSIG) {
                VST(0, VBS_cqbe) = o.Qbe.d[VD_bei];   VST(0, VBS_cqbeci) = o.Qbe.d[VD_bci]; VST(0, VBS_cqbex) = o.Qbex.d[VD_bex];
                VST(0, VBS_cqbc) = o.Qbc.d[VD_bci];   VST(0, VBS_cqbcx) = o.Qbcx.d[VD_bcx]; VST(0, VBS_cqbep) = o.Qbep.d[VD_bep];
                VST(0, VBS_cqbepci) = o.Qbep.d[VD_bci]; VST(0, VBS_cqbeo) = o.Qbeo_Vbe;     VST(0, VBS_cqbco) = o.Qbco_Vbc;
                VST(0, VBS_cqbcp) = o.Qbcp.d[VD_bcp]; VST(0, VBS_cqxf1) = Qxf1_Vxf1;       VST(0, VBS_cqxf2) = Qxf2_Vxf2;
                return NGB_OK;
            }
            if (order != 1 && order != 2) return NGB_E_ORDER;
            if (mode & NGB_MODEINITTRAN) {
                static const unsigned char q8[] = { VBS_qbe, VBS_qbex, VBS_qbc, VBS_qbcx, VBS_qbep, VBS_qbeo, VBS_qbco, VBS_qbcp };
                for (unsigned k = 0; k < sizeof q8; k++) VST(1, q8[k]) = VST(0, q8[k]);
                if (selfheat) VST(1, VBS_qcth) = VST(0, VBS_qcth);
                if (excess) { VST(1, VBS_qxf1) = VST(0, VBS_qxf1); VST(1, VBS_qxf2) = VST(0, VBS_qxf2); }
            }
            { const double cq = INTEGRATE(VBS_qbe);  VST(0, VBS_cqbe) = cq;  Ibe_Vbei = Ibe_Vbei + ag0 * o.Qbe.d[VD_bei];   Ibe = Ibe + cq; }
            { const double cq = INTEGRATE(VBS_qbex); VST(0, VBS_cqbex) = cq; Ibex_Vbex = Ibex_Vbex + ag0 * o.Qbex.d[VD_bex]; Ibex = Ibex + cq; }
            { const double cq = INTEGRATE(VBS_qbc);  VST(0, VBS_cqbc) = cq;  Ibc_Vbci = Ibc_Vbci + ag0 * o.Qbc.d[VD_bci];   Ibc = Ibc + cq; }
            { const double cq = INTEGRATE(VBS_qbcx); VST(0, VBS_cqbcx) = cq; gbcx = ag0 * o.Qbcx.d[VD_bcx]; cbcx = cq; }
            { const double cq = INTEGRATE(VBS_qbep); VST(0, VBS_cqbep) = cq; Ibep_Vbep = Ibep_Vbep + ag0 * o.Qbep.d[VD_bep]; Ibep = Ibep + cq; }
            { const double cq = INTEGRATE(VBS_qbcp); VST(0, VBS_cqbcp) = cq; Ibcp_Vbcp = Ibcp_Vbcp + ag0 * o.Qbcp.d[VD_bcp]; Ibcp = Ibcp + cq; }
            if (selfheat) { const double cq = INTEGRATE(VBS_qcth); VST(0, VBS_cqcth) = cq; Icth_Vrth = ag0 * o.Qcth_Vrth; Icth = cq; }
            if (excess) {
                { const double cq = INTEGRATE(VBS_qxf1); VST(0, VBS_cqxf1) = cq; Ixf1_Vxf1 += ag0 * Qxf1_Vxf1; Ixf1 += cq; }
                { const double cq = INTEGRATE(VBS_qxf2); VST(0, VBS_cqxf2) = cq; Ixf2_Vxf2 += ag0 * Qxf2_Vxf2; Ixf2 += cq; }
            }
            if (mode & NGB_MODEINITTRAN) {
                static const unsigned char c6[] = { VBS_cqbe, VBS_cqbex, VBS_cqbc, VBS_cqbcx, VBS_cqbep, VBS_cqbcp };
                for (unsigned k = 0; k < sizeof c6; k++) VST(1, c6[k]) = VST(0, c6[k]);
                if (selfheat) VST(1, VBS_cqcth) = VST(0, VBS_cqcth);
                if (excess) { VST(1, VBS_cqxf1) = VST(0, VBS_cqxf1); VST(1, VBS_cqxf2) = VST(0, VBS_cqxf2); }
            }
        }
    }

    /* convergence flag */
    if (!(mode & NGB_MODEINITFIX) || !off) {
        if (icheck == 1) {
#ifdef __CUDA_ARCH__
            atomicAdd(&c->ctl.noncon[s], 1);
#else
            c->ctl.noncon[s] += 1;
#endif
        }
    }
    /* outer base-emitter and base-collector overlap charges */
    if (mode & (NGB_MODETRAN | NGB_MODEAC)) {
        if (order != 1 && order != 2) return NGB_E_ORDER;
        { const double cq = INTEGRATE(VBS_qbeo); VST(0, VBS_cqbeo) = cq; gqbeo = ag0 * o.Qbeo_Vbe; }
        { const double cq = INTEGRATE(VBS_qbco); VST(0, VBS_cqbco) = cq; gqbco = ag0 * o.Qbco_Vbc; }
        if (mode & NGB_MODEINITTRAN) { VST(1, VBS_cqbeo) = VST(0, VBS_cqbeo); VST(1, VBS_cqbco) = VST(0, VBS_cqbco); }
        if (c->ctl.lte) {                   /* VBICtrunc */
            static const unsigned char q8[] = { VBS_qbe, VBS_qbex, VBS_qbc, VBS_qbcx, VBS_qbep, VBS_qbeo, VBS_qbco, VBS_qbcp };
            for (unsigned k = 0; k < sizeof q8; k++)
                ngb_lte_state(&c->ctl, s, c->state, VBS_COUNT, (size_t)c->T, t, head, q8[k], order);
        }
    }
#undef INTEGRATE
    VST(0, VBS_vrth) = Vrth;
    if (excess) { VST(0, VBS_vxf1) = Vxf1; VST(0, VBS_vxf2) = Vxf2; }
    VST(0, VBS_vbei) = Vbei; VST(0, VBS_vbex) = Vbex; VST(0, VBS_vbci) = Vbci; VST(0, VBS_vbcx) = Vbcx; VST(0, VBS_vbep) = Vbep;
    VST(0, VBS_vrci) = Vrci; VST(0, VBS_vrbi) = Vrbi; VST(0, VBS_vrbp) = Vrbp; VST(0, VBS_vbcp) = Vbcp;
    VST(0, VBS_ibe) = Ibe; VST(0, VBS_ibe_Vbei) = Ibe_Vbei; VST(0, VBS_ibex) = Ibex; VST(0, VBS_ibex_Vbex) = Ibex_Vbex;
    VST(0, VBS_iciei) = Iciei; VST(0, VBS_iciei_Vbei) = Iciei_Vbei; VST(0, VBS_iciei_Vbci) = Iciei_Vbci;
    VST(0, VBS_iciei_Vrth) = Iciei_Vrth; VST(0, VBS_iciei_Vxf2) = Iciei_Vxf2;
    VST(0, VBS_ibc) = Ibc; VST(0, VBS_ibc_Vbci) = Ibc_Vbci; VST(0, VBS_ibc_Vbei) = Ibc_Vbei;
    VST(0, VBS_ibep) = Ibep; VST(0, VBS_ibep_Vbep) = Ibep_Vbep;
    VST(0, VBS_irci) = Irci; VST(0, VBS_irci_Vrci) = Irci_Vrci; VST(0, VBS_irci_Vbci) = Irci_Vbci; VST(0, VBS_irci_Vbcx) = Irci_Vbcx;
    VST(0, VBS_irbi) = Irbi; VST(0, VBS_irbi_Vrbi) = Irbi_Vrbi; VST(0, VBS_irbi_Vbei) = Irbi_Vbei; VST(0, VBS_irbi_Vbci) = Irbi_Vbci;
    VST(0, VBS_irbp) = Irbp; VST(0, VBS_irbp_Vrbp) = Irbp_Vrbp; VST(0, VBS_irbp_Vbep) = Irbp_Vbep; VST(0, VBS_irbp_Vbci) = Irbp_Vbci;
    VST(0, VBS_ibcp) = Ibcp; VST(0, VBS_ibcp_Vbcp) = Ibcp_Vbcp;
    VST(0, VBS_iccp) = Iccp; VST(0, VBS_iccp_Vbep) = Iccp_Vbep; VST(0, VBS_iccp_Vbci) = Iccp_Vbci; VST(0, VBS_iccp_Vbcp) = Iccp_Vbcp;
    VST(0, VBS_gqbeo) = gqbeo; VST(0, VBS_gqbco) = gqbco;
    VST(0, VBS_ircx_Vrcx) = Ircx_Vrcx; VST(0, VBS_irbx_Vrbx) = Irbx_Vrbx; VST(0, VBS_irs_Vrs) = Irs_Vrs; VST(0, VBS_ire_Vre) = Ire_Vre;
    VST(0, VBS_ixf1) = Ixf1; VST(0, VBS_ixf1_Vbei) = Ixf1_Vbei; VST(0, VBS_ixf1_Vbci) = Ixf1_Vbci; VST(0, VBS_ixf1_Vxf2) = Ixf1_Vxf2;
    VST(0, VBS_ixf1_Vxf1) = Ixf1_Vxf1; VST(0, VBS_ixf1_Vrth) = Ixf1_Vrth; VST(0, VBS_ixf2) = Ixf2; VST(0, VBS_ixf2_Vxf1) = Ixf2_Vxf1; VST(0, VBS_ixf2_Vxf2) = Ixf2_Vxf2;
    if (selfheat) { VST(0, VBS_cqcth) = Icth; VST(0, VBS_icth_Vrth) = Icth_Vrth; }

    /* stamps, statement order of vbicload.c:1076-1267 */
    {
        const double rc_beo = type * (VST(0, VBS_cqbeo) - Vbe * gqbeo);
        const double rc_bco = type * (VST(0, VBS_cqbco) - Vbc * gqbco);
        const double rc_be = type * (Ibe - Ibe_Vbei * Vbei);
        const double rc_bex = type * (Ibex - Ibex_Vbex * Vbex);
        const double rc_ciei = type * (Iciei - Iciei_Vbei * Vbei - Iciei_Vbci * Vbci);
        const double rc_bc = type * (Ibc - Ibc_Vbci * Vbci - Ibc_Vbei * Vbei);
        const double rc_bep = type * (Ibep - Ibep_Vbep * Vbep);
        const double rc_rci = type * (Irci - Irci_Vrci * Vrci - Irci_Vbci * Vbci - Irci_Vbcx * Vbcx);
        const double rc_rbi = type * (Irbi - Irbi_Vrbi * Vrbi - Irbi_Vbei * Vbei - Irbi_Vbci * Vbci);
        const double rc_rbp = type * (Irbp - Irbp_Vrbp * Vrbp - Irbp_Vbep * Vbep - Irbp_Vbci * Vbci);
        const double rc_bcp = type * (Ibcp - Ibcp_Vbcp * Vbcp);
        const double rc_ccp = type * (Iccp - Iccp_Vbep * Vbep - Iccp_Vbci * Vbci - Iccp_Vbcp * Vbcp);
        /* excess phase */
        const double rc_cixf = type * (-Iciei_Vxf2 * Vxf2);
        const double rc_xf1 = Ixf1 - Ixf1_Vbci * Vbci - Ixf1_Vbei * Vbei - Ixf1_Vxf1 * Vxf1 - Ixf1_Vxf2 * Vxf2;
        const double rc_xf2 = Ixf2 - Ixf2_Vxf1 * Vxf1 - Ixf2_Vxf2 * Vxf2;
        /* self-heating: the d/dVrth terms of every element, the thermal capacitance and the dissipated power */
        const double rt_be = -Ibe_Vrth * Vrth, rt_bex = -Ibex_Vrth * Vrth, rt_ciei = -Iciei_Vrth * Vrth, rt_bc = -Ibc_Vrth * Vrth;
        const double rt_bep = -Ibep_Vrth * Vrth, rt_rcx = -Ircx_Vrth * Vrth, rt_rci = -Irci_Vrth * Vrth, rt_rbx = -Irbx_Vrth * Vrth;
        const double rt_rbi = -Irbi_Vrth * Vrth, rt_re = -Ire_Vrth * Vrth, rt_rbp = -Irbp_Vrth * Vrth, rt_bcp = -Ibcp_Vrth * Vrth;
        const double rt_ccp = -Iccp_Vrth * Vrth, rt_rs = -Irs_Vrth * Vrth, rt_xf1 = -Ixf1_Vrth * Vrth;
        const double Irth_Vrth = o.Irth_Vrth, Ith_Vrth = o.Ith.d[VD_rth];
        const double Ith_Vbei = o.Ith.d[VD_bei], Ith_Vbci = o.Ith.d[VD_bci], Ith_Vcei = o.Ith_Vcei, Ith_Vbex = o.Ith.d[VD_bex];
        const double Ith_Vbep = o.Ith.d[VD_bep], Ith_Vbcp = o.Ith.d[VD_bcp], Ith_Vcep = o.Ith_Vcep, Ith_Vrci = o.Ith.d[VD_rci];
        const double Ith_Vbcx = o.Ith.d[VD_bcx], Ith_Vrbi = o.Ith.d[VD_rbi], Ith_Vrbp = o.Ith.d[VD_rbp];
        const double Ith_Vrcx = o.Ith_Vrcx, Ith_Vrbx = o.Ith_Vrbx, Ith_Vre = o.Ith_Vre, Ith_Vrs = o.Ith_Vrs;
        const double rc_cth = Icth - Icth_Vrth * Vrth;
        const double rc_ith = -o.Ith.v - Ith_Vrth * Vrth
                              - Ith_Vbei * Vbei - Ith_Vbci * Vbci - Ith_Vcei * Vcei
                              - Ith_Vbex * Vbex - Ith_Vbep * Vbep - Ith_Vbcp * Vbcp
                              - Ith_Vcep * Vcep - Ith_Vrci * Vrci - Ith_Vbcx * Vbcx
                              - Ith_Vrbi * Vrbi - Ith_Vrbp * Vrbp
                              - Ith_Vrcx * Vrcx - Ith_Vrbx * Vrbx - Ith_Vre * Vre - Ith_Vrs * Vrs;
        int k = 0;
#define VB_PUT(v) do { const int r_ = NGB_LDG(&c->spos[(k) * c->ninst + inst]); if (r_ >= 0) c->stamp[(size_t)r_ * S + s] = (v); k++; } while (0)
#define VB_R(n, v) VB_PUT(v);
#define VB_M(r, cc, v) VB_PUT(v);
        NGB_VBIC_STAMPS(VB_R, VB_M, VB_R, VB_M, VB_R, VB_M, VB_R, VB_M)
#undef VB_R
#undef VB_M
#undef VB_PUT
    }
#undef VST
#undef VAUX
    return NGB_OK;
}
#endif
