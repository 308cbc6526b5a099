/* c_compat.h -- lets the plain-C host files include the .cuh headers (which only contain
 * C-compatible declarations outside NGB_HD bodies). */
#ifndef NGB_C_COMPAT_H
#define NGB_C_COMPAT_H
#endif
