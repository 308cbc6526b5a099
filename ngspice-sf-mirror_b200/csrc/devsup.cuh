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
            if (arg > 0) vnew = vold + vt * (2 + log(arg - 2));
            else         vnew = vold - vt * (2 + log(2 - arg));
        } else {
            vnew = vt * log(vnew / vt);
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
#endif
