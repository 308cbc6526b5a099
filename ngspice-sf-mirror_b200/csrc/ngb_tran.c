/* placeholder until the transient driver lands */
#include "ngb_host.h"
#include "../../include/ngb200.h"
void ngb_tran_free(struct ngb_batch *b) { (void)b; }
int ngbTranRun(ngb_batch *b, int max_points, const int *save_eq, int nsave) { (void)b; (void)max_points; (void)save_eq; (void)nsave; return NGB_E_UNSUPP; }
int ngbTranStats(ngb_batch *b, int *a, int *r, int *n, int *p) { (void)b; (void)a; (void)r; (void)n; (void)p; return NGB_E_UNSUPP; }
long ngbTranWaveBytes(ngb_batch *b) { (void)b; return 0; }
int ngbTranWaves(ngb_batch *b, double *t, double *v) { (void)b; (void)t; (void)v; return NGB_E_UNSUPP; }
int ngbCircuitAnalyze(ngb_circuit *c, const double *Ax) { (void)c; (void)Ax; return NGB_E_UNSUPP; }
