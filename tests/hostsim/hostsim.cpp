/* tests/hostsim/hostsim.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU stand-in for the device runtime of ngspice-sf-mirror_b200/csrc/ngb_dev.h: the same
 * kernel bodies (NGB_HD functions) executed by plain loops, one "thread" at a time, so that
 * the CPU-only CI can check kernel logic against the oracle.  This object is linked only into
 * tests/hostsim/libngb200_hostsim.so; the product library (libngb200.so) contains the CUDA
 * implementation and nothing else. */
#include <cstdlib>
#include <cstring>
#include <vector>
#include "ngb_dev.h"
#include "ngb_kernels.cuh"
#include "vbic_eval.cuh"

static long g_launches = 0;
extern "C" {
const char *ngb_dev_backend(void) { return "hostsim"; }
int ngb_dev_init(int) { return 0; }
void *ngb_dev_malloc(size_t bytes) { return calloc(bytes ? bytes : 1, 1); }
void ngb_dev_free(void *p) { free(p); }
void ngb_dev_l2_persist(const void *, size_t) {}
int ngb_dev_h2d(void *d, const void *s, size_t n) { memcpy(d, s, n); return 0; }
int ngb_dev_d2h(void *d, const void *s, size_t n) { memcpy(d, s, n); return 0; }
int ngb_dev_memset(void *d, int v, size_t n) { memset(d, v, n); return 0; }
int ngb_dev_sync(void) { return 0; }
long ngb_dev_launch_count(void) { return g_launches; }
void *ngb_dev_stream(void) { return nullptr; }
int ngb_dev_set_stream(void *) { return 0; }
void ngb_dev_profile(int, int) {}
int ngb_dev_profile_read(double *ms, long *n) { if (ms) *ms = 0; if (n) *n = 0; return 0; }
int ngb_dev_profile_due(void) { return 0; }
void ngb_dev_stage_begin(void) {}
void ngb_dev_stage_mark(int) {}
int ngb_dev_stage_read(double *ms) { for (int k = 0; k < 8; k++) ms[k] = 0; return 0; }
int ngb_dev_fp64_peak(double out[3]) { out[0] = out[1] = out[2] = 0; return NGB_E_PANIC; }
int ngb_dev_branch_begin(void) { return 0; }
void ngb_dev_branch(int) {}
int ngb_dev_branch_end(void) { return 0; }
int ngb_dev_graph_begin(void) { return -1; }           /* no graphs on the host build: every step is launched directly */
int ngb_dev_graph_end(void **, int *) { return -1; }
int ngb_dev_graph_launch(void *, int) { return -1; }
void ngb_dev_graph_destroy(void *) {}

int ngb_launch_bsim4_load(const B4Ctx *c, int *errflag)
{
    g_launches++;
    for (size_t t = 0; t < (size_t)c->T; t++) { int e = b4_load_thread_variant(c, t); if (e && errflag && !errflag[0]) errflag[0] = e; }
    return 0;
}
int ngb_launch_bsim4_lte(const B4Ctx *c)
{
    g_launches++;
    for (int s = 0; s < c->S; s++)
        if (b4_lte_wanted(c, s)) {
            double m1 = 1e300, m2 = 1e300;
            for (int inst = 0; inst < c->ninst; inst++) b4_lte_inst(c, inst, s, &m1, &m2);
            ngb_atomic_min_pos(&c->ctl.lte[s], m1); ngb_atomic_min_pos(&c->ctl.lte2[s], m2);
        }
    return 0;
}
int ngb_launch_cap_load(const NgbCapCtx *c, int *errflag)
{
    g_launches++;
    for (size_t t = 0; t < (size_t)c->T; t++) { int e = ngb_cap_thread(c, t); if (e && errflag && !errflag[0]) errflag[0] = e; }
    return 0;
}
int ngb_launch_bsim3_load(const B3Ctx *c, int *errflag)
{
    g_launches++;
    for (size_t t = 0; t < (size_t)c->T; t++) { int e = b3_load_thread(c, t); if (e && errflag && !errflag[0]) errflag[0] = e; }
    return 0;
}
int ngb_launch_vbic_load(const NgbVbicCtx *c, int *errflag)
{
    g_launches++;
    for (size_t t = 0; t < (size_t)c->T; t++) { int e = vbic_load_thread(c, t); if (e && errflag && !errflag[0]) errflag[0] = e; }
    return 0;
}
int ngb_launch_dio_load(const NgbDioCtx *c, int *errflag)
{
    g_launches++;
    for (size_t t = 0; t < (size_t)c->T; t++) { int e = ngb_dio_thread(c, t); if (e && errflag && !errflag[0]) errflag[0] = e; }
    return 0;
}
int ngb_launch_src_load(const NgbSrcCtx *c)
{
    g_launches++;
    for (size_t t = 0; t < (size_t)c->T; t++) ngb_src_thread(c, t);
    return 0;
}
int ngb_launch_assemble(const NgbAsmCtx *c)
{
    g_launches++;
    const size_t n = (size_t)(c->nnz + c->neq1) * c->S;
    for (size_t u = 0; u < n; u++) ngb_asm_thread(c, u);
    if (c->nov > 0) { g_launches++; for (int s = 0; s < c->S; s++) ngb_override_thread(c, s); }
    return 0;
}
int ngb_launch_lu(const NgbLuCtx *c)
{
    g_launches++;
    std::vector<double> V(c->sch.nV + 1), Rs(c->sch.n + 1), Z(c->sch.ntask + 1), As(c->sch.nnz + 1), P(c->pk.maxlp + 1);
    for (int s = 0; s < c->S; s++) {
        if (c->pk.ok2 && !getenv("NGB_LU_V1")) ngb_lu_sample_pk2(c, c->pk.blob2, s, 0, 1, V.data(), Rs.data(), Z.data(), P.data(), 0xffffffffu);
        else if (c->pk.ok) ngb_lu_sample_packed(c, c->pk.blob, s, 0, 1, V.data(), Rs.data(), Z.data(), As.data(), P.data(), 0xffffffffu);
        else ngb_lu_sample(c, s, 0, 1, V.data(), Rs.data(), Z.data());
    }
    return 0;
}
int ngb_launch_clear_i32(int *p, int value, int n) { for (int i = 0; i < n; i++) p[i] = value; return 0; }
int ngb_launch_fill_f64(double *p, double value, int n) { for (int i = 0; i < n; i++) p[i] = value; return 0; }
int ngb_launch_tran_control(const NgbTranCtx *c)
{
    g_launches++;
    for (int s = 0; s < c->S; s++) ngb_tran_control(c, s);
    return 0;
}
}

/* test entry points for the libm-compatible exp/log (tests/test_math_replica.py) */
extern "C" void hostsim_exp(const double *x, double *y, int n) { for (int i = 0; i < n; i++) y[i] = ngb_exp(x[i]); }
extern "C" void hostsim_pow(const double *x, const double *y, double *z, int n) { for (int i = 0; i < n; i++) z[i] = ngb_pow(x[i], y[i]); }
extern "C" void hostsim_log(const double *x, double *y, int n) { for (int i = 0; i < n; i++) y[i] = ngb_log(x[i]); }
