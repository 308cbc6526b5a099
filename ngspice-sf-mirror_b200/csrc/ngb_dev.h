/* ngb_dev.h -- the device runtime the host code is written against: memory, copies and
 * one launcher per kernel.  The product library implements it in ngb_cuda.cu (CUDA, sm_100a,
 * no fallback).  tests/hostsim implements the same interface on the CPU so the CPU-only CI
 * can single-step the kernel bodies; that object is never part of the product library. */
#ifndef NGB_DEV_H
#define NGB_DEV_H
#include <stddef.h>
#include "ngb_types.h"
#include "bsim4_eval.cuh"
#include "dio_eval.cuh"
#include "bsim3_eval.cuh"
#include "vbic_types.h"
#include "ngb_tran.cuh"

#ifdef __cplusplus
extern "C" {
#endif

const char *ngb_dev_backend(void);              /* "cuda-sm_100a" or "hostsim" */
int   ngb_dev_init(int device);
void *ngb_dev_malloc(size_t bytes);             /* zero-filled */
void  ngb_dev_free(void *p);
void  ngb_dev_l2_persist(const void *p, size_t bytes);   /* access-policy window (L2 persisting) of the launch streams; NULL: none */
int   ngb_dev_h2d(void *dst, const void *src, size_t bytes);
int   ngb_dev_d2h(void *dst, const void *src, size_t bytes);
int   ngb_dev_memset(void *dst, int byte, size_t bytes);
int   ngb_dev_sync(void);
long  ngb_dev_launch_count(void);               /* kernels launched so far (bench "gpu_launches") */
void *ngb_dev_stream(void);
int   ngb_dev_set_stream(void *stream);         /* adopt a caller-owned cudaStream_t */
void  ngb_dev_profile(int enable, int every);   /* CUDA-event timing of sampled bsim4_load launches */
int   ngb_dev_profile_read(double *ms_sum, long *count);
/* returns 1 when the sampled-launch timing wants THIS Newton step (the caller then launches it kernel by
 * kernel instead of replaying a graph) */
int   ngb_dev_profile_due(void);
void  ngb_dev_stage_begin(void);                /* per-stage timing of the sampled steps: see ngb_cuda.cu */
void  ngb_dev_stage_mark(int slot);
int   ngb_dev_stage_read(double ms[8]);
int   ngb_dev_fp64_peak(double out[3]);     /* measured DFMA and DADD/DMUL flop/s of this GPU */
/* CUDA graph of one Newton step: begin capture on the launch stream, end + instantiate, replay */
int   ngb_dev_graph_begin(void);
int   ngb_dev_graph_end(void **exec, int *nodes);
int   ngb_dev_graph_launch(void *exec, int nodes);
void  ngb_dev_graph_destroy(void *exec);                     /* cudaStream_t the kernels are launched on */

/* the load kernels of the device types are independent: between begin and end, ngb_dev_branch(i) sends the
 * following launches to side stream i (i < 0: the main stream); end joins them all into the main stream */
int   ngb_dev_branch_begin(void);
void  ngb_dev_branch(int i);
int   ngb_dev_branch_end(void);

int ngb_launch_bsim4_load(const B4Ctx *c, int *errflag);
int ngb_launch_bsim4_lte(const B4Ctx *c);          /* BSIM4trunc for the samples that can have converged */
int ngb_launch_cap_load(const NgbCapCtx *c, int *errflag);
int ngb_launch_src_load(const NgbSrcCtx *c);
int ngb_launch_assemble(const NgbAsmCtx *c);
int ngb_launch_lu(const NgbLuCtx *c);
int ngb_launch_clear_i32(int *p, int value, int n);
int ngb_launch_tran_control(const NgbTranCtx *c);
int ngb_launch_fill_f64(double *p, double value, int n);
int ngb_launch_dio_load(const NgbDioCtx *c, int *errflag);
int ngb_launch_bsim3_load(const B3Ctx *c, int *errflag);
int ngb_launch_vbic_load(const NgbVbicCtx *c, int *errflag);

#ifdef __cplusplus
}
#endif
#endif
