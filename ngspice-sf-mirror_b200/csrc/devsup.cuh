/* devsup.cuh -- Newton step limiters and the companion-model integrator as inlines.
 *
 * Branch-for-branch restatements of the reference's scalar helpers so that a device
 * thread takes exactly the decisions the CPU path takes:
 *   limvds  : DEVlimvds   src/spicelib/devices/devsup.c:21-44
 *   pnjlim  : DEVpnjlim   src/spicelib/devices/devsup.c:50-85
 *   fetlim  : DEVfetlim   src/spicelib/devices/devsup.c:93-151
 *   ngb_integrate : NIintegrate src/maths/ni/niinteg.c:17-80 (TRAPEZOIDAL orders 1,2)
 */
#ifndef NGB_DEVSUP_CUH
#define NGB_DEVSUP_CUH
#include "ngb_common.h"
#include "ngb_types.h"
#include "ngb_math.cuh"

NGB_HD double ngb_limvds(double vnew, double vold)
{
    if (vold >= 3.5) {
        if (vnew > vold) {
            double cap = 3.0 * vold + 2.0;
            if (cap < vnew) vnew = cap;
        } else if (vnew < 3.5) {
            if (vnew < 2.0) vnew = 2.0;
        }
    } else {
        if (vnew > vold) { if (vnew > 4.0) vnew = 4.0; }
        else             { if (vnew < -0.5) vnew = -0.5; }
    }
    return vnew;
}

NGB_HD double ngb_pnjlim(double vnew, double vold, double vt, double vcrit, int *icheck)
{
    if ((vnew > vcrit) && (fabs(vnew - vold) > (vt + vt))) {
        if (vold > 0) {
            double arg = (vnew - vold) / vt;
            if (arg > 0) vnew = vold + vt * (2 + ngb_log(arg - 2));
            else         vnew = vold - vt * (2 + ngb_log(2 - arg));
        } else {
            vnew = vt * ngb_log(vnew / vt);
        }
        *icheck = 1;
    } else if (vnew < 0) {
        double arg = (vold > 0) ? (-1 * vold - 1) : (2 * vold - 1);
        if (vnew < arg) { vnew = arg; *icheck = 1; }
        else *icheck = 0;
    } else {
        *icheck = 0;
    }
    return vnew;
}

NGB_HD double ngb_fetlim(double vnew, double vold, double vto)
{
    double vtsthi = fabs(2 * (vold - vto)) + 2;
    double vtstlo = fabs(vold - vto) + 1;
    double vtox = vto + 3.5;
    double delv = vnew - vold;

    if (vold >= vto) {
        if (vold >= vtox) {
            if (delv <= 0) {                       /* going off */
                if (vnew >= vtox) {
                    if (-delv > vtstlo) vnew = vold - vtstlo;
                } else {
                    double lo = vto + 2;
                    if (vnew < lo) vnew = lo;
                }
            } else if (delv >= vtsthi) {           /* staying on */
                vnew = vold + vtsthi;
            }
        } else {                                   /* middle region */
            if (delv <= 0) { double lo = vto - .5; if (vnew < lo) vnew = lo; }
            else           { double hi = vto + 4;  if (vnew > hi) vnew = hi; }
        }
    } else {                                       /* off */
        if (delv <= 0) {
            if (-delv > vtsthi) vnew = vold - vtsthi;
        } else {
            double vtemp = vto + .5;
            if (vnew <= vtemp) { if (delv > vtstlo) vnew = vold + vtstlo; }
            else vnew = vtemp;
        }
    }
    return vnew;
}

/* companion current of a charge state: returns ccap for state0, given q0, q1 and the
 * previous companion current c1 (NIintegrate, TRAPEZOIDAL). */
NGB_HD double ngb_integrate_trap(int order, double ag0, double ag1, double q0, double q1, double c1)
{
    if (order == 1) return ag0 * q0 + ag1 * q1;
    return -c1 * ag1 + ag0 * (q0 - q1);
}
/* NIintegrate for both methods (niinteg.c:17-80): GEAR accumulates ag[order]*q[order] ... ag[0]*q[0] from the oldest
 * charge down, starting from zero; DCtran never raises the order above 2 (dctran.c:794-826), so q2 is the oldest */
NGB_HD double ngb_integrate(int gear, int order, double ag0, double ag1, double ag2, double q0, double q1, double q2, double c1)
{
    if (gear) {
        double cc = 0;
        if (order == 2) cc += ag2 * q2;
        cc += ag1 * q1;
        cc += ag0 * q0;
        return cc;
    }
    return ngb_integrate_trap(order, ag0, ag1, q0, q1, c1);
}

/* CKTterr (src/spicelib/analysis/cktterr.c:10-79), TRAPEZOIDAL: the step size the local
 * truncation error of one charge state allows.  q[0..order+1] = CKTstates[i][qcap],
 * cc0/cc1 = CKTstate0/1[ccap], dold = CKTdeltaOld. */
NGB_HD double ngb_terr(int gear, int order, const double *q, double cc0, double cc1, const double *dold,
                       double delta, double reltol, double abstol, double chgtol, double trtol)
{
    double diff[4], deltmp[3], volttol, chargetol, tol, factor, del, a, b;
    int i, j;
    a = fabs(cc0); b = fabs(cc1);
    volttol = abstol + reltol * ((a > b) ? a : b);
    a = fabs(q[0]); b = fabs(q[1]);
    chargetol = (a > b) ? a : b;
    chargetol = reltol * ((chargetol > chgtol) ? chargetol : chgtol) / delta;
    tol = (volttol > chargetol) ? volttol : chargetol;
    for (i = order + 1; i >= 0; i--) diff[i] = q[i];
    for (i = 0; i <= order; i++) deltmp[i] = dold[i];
    j = order;
    for (;;) {
        for (i = 0; i <= j; i++) diff[i] = (diff[i] - diff[i + 1]) / deltmp[i];
        if (--j < 0) break;
        for (i = 0; i <= j; i++) deltmp[i] = deltmp[i + 1] + dold[i];
    }
    factor = (order == 1) ? .5 : (gear ? .2222222222 : .08333333333);       /* gearCoeff / trapCoeff, cktterr.c:23-35 */
    a = factor * fabs(diff[0]);
    del = trtol * tol / ((abstol > a) ? abstol : a);
    if (order == 2) del = sqrt(del);
    return del;
}

/* atomic min on non-negative doubles (bit patterns of non-negative doubles order like integers) */
NGB_HD void ngb_atomic_min_pos(double *addr, double v)
{
#ifdef __CUDA_ARCH__
    atomicMin((unsigned long long *)addr, (unsigned long long)__double_as_longlong(v));
#else
    if (v < *addr) *addr = v;
#endif
}

/* LTE contribution of one charge state of thread t: state array [hist][nstate][T] */
/* the two bounds of one charge state (order as is, and order 2 while the order is still 1), not yet reduced */
NGB_HD void ngb_lte_values(const NgbCtl *k, int s, const double *state, int nstate, size_t T, size_t t,
                           int head, int kq, int order, double *d1, double *d2)
{
    const int nh = k->nhist, S = k->S;
    double q[4], dold[3];
    const double delta = NGB_LDG(&k->delta[s]);
    int i;
#define LST(h, kk) state[((size_t)(((head) + (h)) % nh) * nstate + (kk)) * T + t]
    const double cc0 = LST(0, kq + 1), cc1 = LST(1, kq + 1);
    for (i = 0; i < 3; i++) dold[i] = NGB_LDG(&k->delta_old[(size_t)i * S + s]);
    for (i = 0; i < 4; i++) q[i] = (i <= order + 1 || (order == 1 && nh >= 4)) && i < nh ? LST(i, kq) : 0.0;
    /* literal orders: the difference tables unroll into registers */
    *d1 = (order == 1) ? ngb_terr(k->gear, 1, q, cc0, cc1, dold, delta, k->reltol, k->abstol, k->chgtol, k->trtol)
                       : ngb_terr(k->gear, 2, q, cc0, cc1, dold, delta, k->reltol, k->abstol, k->chgtol, k->trtol);
    *d2 = (order == 1 && nh >= 4) ? ngb_terr(k->gear, 2, q, cc0, cc1, dold, delta, k->reltol, k->abstol, k->chgtol, k->trtol) : 1e300;
#undef LST
}

/* one out-of-line copy per kernel: the BSIM4 load carried ten inlined CKTterr bodies (4 k instructions, a sixth of
 * its code) that the transient driver no longer executes there */
#ifdef __CUDACC__
__device__ __noinline__
#else
static inline
#endif
void ngb_lte_state(const NgbCtl *k, int s, const double *state, int nstate, size_t T, size_t t,
                          int head, int kq, int order)
{
    const int nh = k->nhist, S = k->S;
    double q[4], dold[3];
    const double delta = NGB_LDG(&k->delta[s]);
    int i;
#define LST(h, kk) state[((size_t)(((head) + (h)) % nh) * nstate + (kk)) * T + t]
    const double cc0 = LST(0, kq + 1), cc1 = LST(1, kq + 1);
    for (i = 0; i < 3; i++) dold[i] = NGB_LDG(&k->delta_old[(size_t)i * S + s]);
    for (i = 0; i <= order + 1 && i < nh; i++) q[i] = LST(i, kq);
    ngb_atomic_min_pos(&k->lte[s], ngb_terr(k->gear, order, q, cc0, cc1, dold, delta, k->reltol, k->abstol, k->chgtol, k->trtol));
    if (order == 1 && nh >= 4) {
        q[3] = LST(3, kq);
        ngb_atomic_min_pos(&k->lte2[s], ngb_terr(k->gear, 2, q, cc0, cc1, dold, delta, k->reltol, k->abstol, k->chgtol, k->trtol));
    }
#undef LST
}
#endif
