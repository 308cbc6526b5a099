/* ngb_common.h -- shared definitions for the B200 Newton hot path.
 *
 * Every kernel body in csrc/ is written as an NGB_HD function so that the SAME source
 * compiles (a) with nvcc for sm_100a -- the product -- and (b) with plain g++ into the
 * test-only "hostsim" library under tests/hostsim, which lets the CPU CI single-step
 * kernel logic against the oracle without a GPU.  hostsim is never linked into the
 * product library and is not reachable from the C-ABI.
 *
 * Mode bits mirror the reference's CKTmode encoding (src/include/ngspice/cktdefs.h:171-199)
 * because the drop-in boundary passes ckt->CKTmode through unchanged.
 */
#ifndef NGB_COMMON_H
#define NGB_COMMON_H

#include <math.h>
#include <stdint.h>
#include <stddef.h>

#ifdef __CUDACC__
#define NGB_HD __device__ __forceinline__
#define NGB_D __device__ __forceinline__
/* helpers that the BSIM4 evaluation calls twice or more (source and drain junctions, the two gate-edge
 * tunnelling currents, GIDL / GISL): one out-of-line copy each when NGB_OUTLINE_HELPERS is set -- the
 * evaluation is bound by instruction fetch, see DESIGN.md */
#ifdef NGB_OUTLINE_HELPERS
#define NGB_HD_SHARED __device__ __noinline__
#else
#define NGB_HD_SHARED __device__ __forceinline__
#endif
#define NGB_LDG(p) __ldg(p)
#else
#define NGB_HD static inline
#define NGB_D static inline
#define NGB_HD_SHARED static inline
#define NGB_LDG(p) (*(p))
#endif

/* CKTmode bits (cktdefs.h:171-199) */
#define NGB_MODETRAN         0x1
#define NGB_MODEAC           0x2
#define NGB_MODEDC           0x70
#define NGB_MODEDCOP         0x10
#define NGB_MODETRANOP       0x20
#define NGB_MODEDCTRANCURVE  0x40
#define NGB_INITF            0x3f00
#define NGB_MODEINITFLOAT    0x100
#define NGB_MODEINITJCT      0x200
#define NGB_MODEINITFIX      0x400
#define NGB_MODEINITSMSIG    0x800
#define NGB_MODEINITTRAN     0x1000
#define NGB_MODEINITPRED     0x2000
#define NGB_MODEUIC          0x10000

/* integration methods (cktdefs.h TRAPEZOIDAL/GEAR) */
#define NGB_TRAPEZOIDAL 1
#define NGB_GEAR        2

/* error codes mirrored from src/include/ngspice/sperror.h / iferrmsg.h */
#define NGB_OK          0
#define NGB_E_PANIC     1
#define NGB_E_SINGULAR  102
#define NGB_E_ITERLIM   103
#define NGB_E_ORDER     104
#define NGB_E_METHOD    105
#define NGB_E_TIMESTEP  106
#define NGB_E_UNSUPP    10     /* a model option this path does not implement: fail loudly */

/* number of state-history vectors kept per device (CKTstates[0..maxOrder+1], cktsetup.c:192) */
#define NGB_NHIST 4            /* TRAP maxorder 2 -> states 0..3 */

#define NGB_MAX(a,b) ((a) > (b) ? (a) : (b))
#define NGB_MIN(a,b) ((a) < (b) ? (a) : (b))

#endif
