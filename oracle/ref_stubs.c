/* oracle/ref_stubs.c -- TEST INFRASTRUCTURE ONLY.
 * Link-time stand-ins for reference pieces that are deliberately left out of the
 * oracle build (the C++ HICUM model; no other device is affected). */
#include "ngspice/ngspice.h"
#include "ngspice/devdefs.h"
SPICEdev *get_hicum_info(void);
SPICEdev *get_hicum_info(void) { return NULL; }
