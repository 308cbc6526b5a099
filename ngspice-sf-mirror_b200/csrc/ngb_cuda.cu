/* ngb_cuda.cu -- the CUDA (sm_100a) implementation of the device runtime in ngb_dev.h:
 * __global__ wrappers around the kernel bodies, launch geometry, memory and stream handling.
 *
 * Launch geometry (B200: 148 SMs, 64 K registers/SM, up to 227 KB shared memory per CTA):
 *   bsim4_load   one thread per (instance, sample), 128-thread CTAs; FP64-pipe/HBM bound
 *   cap/src/asm  one thread per element, 256-thread CTAs; HBM bound, fully coalesced
 *   lu           one warp per sample (LU values, scale factors and solve vector in shared
 *                memory, __syncwarp between dependency levels), or one CTA per sample when a
 *                sample's LU does not fit a warp's share of shared memory
 * All kernels are launched on one non-default stream; nothing here falls back to the host.
 */
#include <cuda_runtime.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>

#include <cooperative_groups.h>
/* a group is a warp, a CTA, or (nl > 1024: ngb_k_lu_grid, cooperative launch) the whole grid */
#define NGB_GROUP_SYNC() do { if (nl <= 32) __syncwarp(ngb_gsync_mask); else if (nl <= 1024) __syncthreads(); else cooperative_groups::this_grid().sync(); } while (0)
/* programmatic dependent launch (sm_90+): a kernel launched with the attribute may be scheduled while its predecessor in
 * the stream is still draining; it must pass NGB_PDL_WAIT() before it touches anything the predecessor wrote (the wait
 * returns when the predecessor grid has completed and its writes are visible).  NGB_PDL_TRIGGER() lets the successor be
 * scheduled from that point on.  Both are no-ops in a launch without the attribute. */
#define NGB_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define NGB_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#ifndef NGB_B4_CTA
#define NGB_B4_CTA 256
#endif
#ifndef NGB_B4_MINBLOCKS
#define NGB_B4_MINBLOCKS 2      /* 128 registers/thread, 16 warps/SM: measured 1.5x faster than 255 registers */
#endif
#ifndef NGB_LU_CTA_DEFAULT
#define NGB_LU_CTA_DEFAULT 256
#endif
#include "ngb_dev.h"
#include "ngb_kernels.cuh"
#include "vbic_eval.cuh"

static thread_local cudaStream_t g_stream = nullptr;   /* per host thread: batches driven from different threads run concurrently */
/* device-type load kernels of one CKTload are independent of each other: they run on side streams between a
 * fork and a join event (parallel branches of the captured graph); g_cur is the stream a launch goes to */
#define NGB_SIDE 4
static thread_local cudaStream_t g_cur = nullptr, g_side[NGB_SIDE];
static thread_local cudaEvent_t g_fork, g_join[NGB_SIDE];
static thread_local int g_side_ready = 0, g_side_used[NGB_SIDE], g_branching = 0;
static int g_branch_off = -1;
static int g_device = -1;
static std::atomic<long> g_launches{0};
static int g_smem_optin = 0, g_sm_count = 148;

/* NGB_PDL=1: the LU, BSIM4trunc and controller launches of a Newton step carry the programmatic-stream-serialization
 * attribute (their kernels wait with griddepcontrol.wait); 2: the batch assembly as well.  0 / unset: plain launches */
static int g_pdl = -1;
template <typename... KArgs, typename... Args>
static void ngb_launch_dep(int level, void (*kern)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args... args)
{
    if (g_pdl < 0) { const char *e = getenv("NGB_PDL"); g_pdl = e ? atoi(e) : 0; }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    if (g_pdl >= level) { cfg.attrs = at; cfg.numAttrs = 1; }
    cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

extern "C" void ngb_set_error(const char *fmt, ...);

#define CUDA_OK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    ngb_set_error("%s failed: %s", #call, cudaGetErrorString(e_)); return NGB_E_PANIC; } } while (0)

/* ------------------------------------------------------------------ kernels */
/* one instantiation per variant key of bsim4_variants.h plus the generic one */
template <unsigned VK>
__global__ void __launch_bounds__(NGB_B4_CTA, NGB_B4_MINBLOCKS)
ngb_k_bsim4_load(const __grid_constant__ B4Ctx c, int *errflag)
{
    NGB_PDL_TRIGGER();
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)c.T) return;
    const int e = b4_load_thread<VK>(&c, t);
    if (e) atomicMax(errflag, e);
}

__global__ void __launch_bounds__(128)
ngb_k_bsim4_lte(const __grid_constant__ B4Ctx c)
{
    const int s = (int)(((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (s >= c.S || !b4_lte_wanted(&c, s)) return;
    double m1 = 1e300, m2 = 1e300;
    for (int inst = lane; inst < c.ninst; inst += 32) b4_lte_inst(&c, inst, s, &m1, &m2);
    for (int o = 16; o > 0; o >>= 1) {          /* min is exact in any order */
        m1 = fmin(m1, __shfl_xor_sync(0xffffffffu, m1, o));
        m2 = fmin(m2, __shfl_xor_sync(0xffffffffu, m2, o));
    }
    if (lane == 0) { ngb_atomic_min_pos(&c.ctl.lte[s], m1); ngb_atomic_min_pos(&c.ctl.lte2[s], m2); }
}

/* the same bounds with the load's own thread layout (thread = instance x sample, lanes on consecutive samples): every state
 * read is coalesced over the converged samples of a warp (a warp per sample reads 8 useful bytes per 32-byte sector); one
 * atomic minimum per instance and sample.  Default for batches (NGB_LTE_FLAT=0: the warp-per-sample kernel) */
__global__ void __launch_bounds__(256)
ngb_k_bsim4_lte_flat(const __grid_constant__ B4Ctx c)
{
    NGB_PDL_WAIT();
    NGB_PDL_TRIGGER();
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)c.T) return;
    const int inst = (int)(t / (size_t)c.S), s = (int)(t - (size_t)inst * c.S);
    if (!b4_lte_wanted(&c, s)) return;
    double m1 = 1e300, m2 = 1e300;
    b4_lte_inst(&c, inst, s, &m1, &m2);
    if (m1 < 1e300) ngb_atomic_min_pos(&c.ctl.lte[s], m1);
    if (m2 < 1e300) ngb_atomic_min_pos(&c.ctl.lte2[s], m2);
}

__global__ void __launch_bounds__(256)
ngb_k_cap_load(const __grid_constant__ NgbCapCtx c, int *errflag)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)c.T) return;
    const int e = ngb_cap_thread(&c, t);
    if (e) atomicMax(errflag, e);
}

__global__ void __launch_bounds__(256, 2)
ngb_k_bsim3_load(const __grid_constant__ B3Ctx c, int *errflag)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)c.T) return;
    const int e = b3_load_thread(&c, t);
    if (e) atomicMax(errflag, e);
}

__global__ void __launch_bounds__(128)
ngb_k_vbic_load(const __grid_constant__ NgbVbicCtx c, int *errflag)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)c.T) return;
    const int e = vbic_load_thread(&c, t);
    if (e) atomicMax(errflag, e);
}

__global__ void __launch_bounds__(256)
ngb_k_dio_load(const __grid_constant__ NgbDioCtx c, int *errflag)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)c.T) return;
    const int e = ngb_dio_thread(&c, t);
    if (e) atomicMax(errflag, e);
}

__global__ void __launch_bounds__(256)
ngb_k_src_load(const __grid_constant__ NgbSrcCtx c)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (size_t)c.T) return;
    ngb_src_thread(&c, t);
}

__global__ void __launch_bounds__(256)
ngb_k_assemble(const NgbAsmCtx c, size_t total)
{
    const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= total) return;
    ngb_asm_thread(&c, u);
}

/* batches: a CTA takes a tile of 32 targets x 32 samples.  The gather reads the stamp rows with the lanes on the samples
 * (the rows are sample-fastest), the matrix values leave through shared memory with the lanes on the targets: Ax is
 * sample-major for the LU kernel, and written straight from the gather every lane of a store hits another 32-byte sector.
 * Right-hand-side targets are sample-fastest like the stamp rows and are stored from the gather. */
__global__ void __launch_bounds__(256)
ngb_k_assemble_tiled(const NgbAsmCtx c, int ntile_s)
{
    NGB_PDL_WAIT();
    NGB_PDL_TRIGGER();
    __shared__ double tile[32][33];
    const int ts = (int)(blockIdx.x % (unsigned)ntile_s), tt = (int)(blockIdx.x / (unsigned)ntile_s);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int S = c.S, ntg = c.nnz + c.neq1;
    {
        const int s = ts * 32 + lane;
        const bool act = s < S && c.ctl.active[s];
        /* most targets collect one or two stamp rows: the first row of each of the thread's four targets is fetched
         * before any sum starts (four independent DRAM loads in flight), the rest of a list follows in list order */
        int lo[4], hi[4];
        double acc[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int tg = tt * 32 + w * 4 + j;
            lo[j] = hi[j] = 0;
            if (act && tg < ntg) {
                lo[j] = c.tgt_ptr[tg]; hi[j] = c.tgt_ptr[tg + 1];
                if (hi[j] - lo[j] > NGB_ASM_LONG && c.nlong) hi[j] = lo[j] = -1;       /* ngb_k_assemble_long*'s job */
            } else lo[j] = hi[j] = -1;
        }
        int r[4];
#pragma unroll
        for (int j = 0; j < 4; j++) r[j] = (lo[j] < hi[j]) ? c.tgt_rows[lo[j]] : -1;
        double a[4];
#pragma unroll
        for (int j = 0; j < 4; j++) a[j] = (r[j] >= 0) ? c.stamp[(size_t)r[j] * S + s] : 0.0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int tg = tt * 32 + w * 4 + j;
            if (lo[j] < 0) continue;
            acc[j] = 0.0;
            if (r[j] >= 0) acc[j] += a[j];
            for (int p = lo[j] + 1; p < hi[j]; p++) acc[j] += c.stamp[(size_t)c.tgt_rows[p] * S + s];
            if (tg >= c.nnz) { ngb_asm_store(&c, tg, s, acc[j]); continue; }
            if (c.add_diag_gmin && c.slot_diag[tg]) {              /* LoadGmin_CSC, as in ngb_asm_store */
                const double dg = c.ctl.diag_gmin[s];
                if (dg != 0.0) acc[j] += dg;
            }
            tile[w * 4 + j][lane] = acc[j];
        }
    }
    __syncthreads();
    {
        const int tg = tt * 32 + lane;
        if (tg >= c.nnz) return;
        if (c.nlong && c.tgt_ptr[tg + 1] - c.tgt_ptr[tg] > NGB_ASM_LONG) return;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int sl = w * 4 + j, s = ts * 32 + sl;
            if (s < S && c.ctl.active[s]) c.Ax[(size_t)s * c.nnz + tg] = tile[lane][sl];
        }
    }
}

/* long contribution lists (ngb_types.h): level 1 sums chunks of stamp rows, the next levels chunks of chunk totals */
__global__ void __launch_bounds__(256)
ngb_k_assemble_long1(const NgbAsmCtx c, int li, double *part, int nchunk)
{
    const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= (size_t)nchunk * c.S) return;
    const int ch = (int)(u / (size_t)c.S), s = (int)(u - (size_t)ch * c.S);
    if (!c.ctl.active[s]) return;
    const int tg = c.long_tgt[li];
    const int lo = c.tgt_ptr[tg] + ch * NGB_ASM_CHUNK, hi = min(lo + NGB_ASM_CHUNK, c.tgt_ptr[tg + 1]);
    double acc = 0.0;
    int p = lo;
    for (; p + 4 <= hi; p += 4) {
        const double a0 = c.stamp[(size_t)c.tgt_rows[p] * c.S + s], a1 = c.stamp[(size_t)c.tgt_rows[p + 1] * c.S + s];
        const double a2 = c.stamp[(size_t)c.tgt_rows[p + 2] * c.S + s], a3 = c.stamp[(size_t)c.tgt_rows[p + 3] * c.S + s];
        acc += a0; acc += a1; acc += a2; acc += a3;
    }
    for (; p < hi; p++) acc += c.stamp[(size_t)c.tgt_rows[p] * c.S + s];
    part[(size_t)ch * c.S + s] = acc;
}
__global__ void __launch_bounds__(256)
ngb_k_assemble_long2(const NgbAsmCtx c, int li, const double *in, double *out, int n_in, int n_out)
{
    const size_t u = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= (size_t)n_out * c.S) return;
    const int g = (int)(u / (size_t)c.S), s = (int)(u - (size_t)g * c.S);
    if (!c.ctl.active[s]) return;
    const int lo = g * NGB_ASM_CHUNK, hi = min(lo + NGB_ASM_CHUNK, n_in);
    double acc = 0.0;
    int i = lo;
    for (; i + 4 <= hi; i += 4) {
        const double a0 = in[(size_t)i * c.S + s], a1 = in[(size_t)(i + 1) * c.S + s], a2 = in[(size_t)(i + 2) * c.S + s], a3 = in[(size_t)(i + 3) * c.S + s];
        acc += a0; acc += a1; acc += a2; acc += a3;
    }
    for (; i < hi; i++) acc += in[(size_t)i * c.S + s];
    if (n_out == 1) ngb_asm_store(&c, c.long_tgt[li], s, acc);
    else out[(size_t)g * c.S + s] = acc;
}

/* one warp per sample: warp w of the CTA owns sample blockIdx.x * warps + w */
__global__ void ngb_k_lu_warp(const NgbLuCtx c, int per_sample_doubles)
{
    extern __shared__ double smem[];
    const int warps = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int s = blockIdx.x * warps + w;
    if (s >= c.S) return;
    double *V = smem + (size_t)w * per_sample_doubles;
    double *Rs = V + c.sch.nV;
    double *Z = Rs + c.sch.n;
    ngb_lu_sample(&c, s, lane, 32, V, Rs, Z);
}

/* one CTA per sample (larger matrices) */
/* one circuit too large for a CTA's shared memory (config 4's arrays): the whole grid works on one sample at a time, values,
 * scale factors and solve tasks in global memory, a grid-wide barrier between the levels of the schedule.  Same body,
 * same order of operations as every other LU kernel here */
__global__ void __launch_bounds__(256)
ngb_k_lu_grid(const NgbLuCtx c)
{
    const int nl = (int)(gridDim.x * blockDim.x), lane = (int)(blockIdx.x * blockDim.x + threadIdx.x);
    for (int s = 0; s < c.S; s++)
        ngb_lu_sample(&c, s, lane, nl, c.gV + (size_t)s * c.sch.nV, c.gRs + (size_t)s * c.sch.n, c.gZ + (size_t)s * c.sch.ntask);
}

__global__ void ngb_k_lu_block(const NgbLuCtx c)
{
    extern __shared__ double smem[];
    const int s = blockIdx.x;
    if (s >= c.S) return;
    double *V = smem;
    double *Rs = V + c.sch.nV;
    double *Z = Rs + c.sch.n;
    ngb_lu_sample(&c, s, threadIdx.x, blockDim.x, V, Rs, Z);
}

/* packed schedule in shared memory: `groups` samples per CTA, `tpg` threads per sample;
 * v2 selects the record packing (ngb_lu_sample_pk2) */
template <int V2>
__global__ void ngb_k_lu_packed(const NgbLuCtx c, int groups, int tpg, int per_sample_doubles)
{
    extern __shared__ double smem[];
    const NgbLuPacked *h = &c.pk;
    unsigned short *sb = reinterpret_cast<unsigned short *>(smem);
    const int blob_u16 = V2 ? h->blob2_u16 : h->blob_u16;
    const int blob_doubles = (blob_u16 + 3) / 4;
    {   /* one coalesced copy of the schedule per CTA */
        const unsigned *src = reinterpret_cast<const unsigned *>(V2 ? h->blob2 : h->blob);
        unsigned *dst = reinterpret_cast<unsigned *>(smem);
        for (int i = threadIdx.x; i < blob_u16 / 2; i += blockDim.x) dst[i] = __ldg(&src[i]);
    }
    /* the schedule is the circuit's, not the step's: under programmatic dependent launch its copy overlaps the tail of
     * the assembly; everything below reads what the assembly wrote */
    NGB_PDL_WAIT();
    NGB_PDL_TRIGGER();
    __syncthreads();
    const int g = threadIdx.x / tpg, lane = threadIdx.x - g * tpg;
    const int s = blockIdx.x * groups + g;
    /* groups narrower than a warp share it: the lanes of samples that sit this launch out (finished, or on
     * another pattern set) leave, the others synchronise among themselves */
    const bool live = g < groups && s < c.S && c.ctl.active[s] && !(c.ctl.lusel && c.ctl.lusel[s] != c.which);
    const unsigned mask = (tpg <= 32) ? __ballot_sync(0xffffffffu, live) : 0xffffffffu;
    if (!live) return;
    double *V = smem + blob_doubles + (size_t)g * per_sample_doubles;
    double *Rs = V + h->nV;
    double *Z = Rs + h->n;
    double *P = Z + h->ntask;
    if (V2) {
        if (tpg == 32) ngb_lu_sample_pk2(&c, sb, s, lane, 32, V, Rs, Z, P, mask);     /* the common case with a constant stride */
        else ngb_lu_sample_pk2(&c, sb, s, lane, tpg, V, Rs, Z, P, mask);
    }
    else ngb_lu_sample_packed(&c, sb, s, lane, tpg, V, Rs, Z, Z /* unused: A is read from global memory */, P, mask);
}

__global__ void __launch_bounds__(128)
ngb_k_tran_control(const NgbTranCtx c)
{
    NGB_PDL_WAIT();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < c.S) ngb_tran_control(&c, s);
}

__global__ void ngb_k_fill_f64(double *p, double value, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = value;
}

__global__ void ngb_k_clear_i32(int *p, int value, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = value;
}


/* FP64-pipe peak of this GPU, measured: eight independent register chains per thread, no memory.
 * fused = 1: DFMA (2 flop per instruction); fused = 0: alternating DADD / DMUL, the instruction mix of code compiled
 * with -fmad=false (1 flop per instruction) -- the denominator of roofline.fp64_frac */
template <int FUSED>
__global__ void __launch_bounds__(256)
ngb_k_fp64_peak(double *out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0 - 1e-9 * seed, c = 1e-7;
    for (int i = 0; i < iters; i++) {
        if (FUSED) {
            a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
            a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
        } else {
            a0 = __dmul_rn(a0, m); a1 = __dadd_rn(a1, c); a2 = __dmul_rn(a2, m); a3 = __dadd_rn(a3, c);
            a4 = __dmul_rn(a4, m); a5 = __dadd_rn(a5, c); a6 = __dmul_rn(a6, m); a7 = __dadd_rn(a7, c);
        }
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 12345.678) out[0] = r;      /* keeps the chains alive, never true */
}

/* ------------------------------------------------------------------ runtime */
extern "C" {

const char *ngb_dev_backend(void) { return "cuda-sm_100a"; }

int ngb_dev_init(int device)
{
    int count = 0;
    if (g_stream && g_device == device) return 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        ngb_set_error("no CUDA device available");
        return NGB_E_PANIC;
    }
    CUDA_OK(cudaSetDevice(device));
    if (g_stream) cudaStreamDestroy(g_stream);
    CUDA_OK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
    g_cur = g_stream;
    if (!g_side_ready) {
        for (int i = 0; i < NGB_SIDE; i++) {
            CUDA_OK(cudaStreamCreateWithFlags(&g_side[i], cudaStreamNonBlocking));
            CUDA_OK(cudaEventCreateWithFlags(&g_join[i], cudaEventDisableTiming));
        }
        CUDA_OK(cudaEventCreateWithFlags(&g_fork, cudaEventDisableTiming));
        g_side_ready = 1;
    }
    CUDA_OK(cudaDeviceGetAttribute(&g_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CUDA_OK(cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, device));
    CUDA_OK(cudaFuncSetAttribute(ngb_k_lu_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin));
    CUDA_OK(cudaFuncSetAttribute(ngb_k_lu_block, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin));
    CUDA_OK(cudaFuncSetAttribute(ngb_k_lu_packed<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin));
    CUDA_OK(cudaFuncSetAttribute(ngb_k_lu_packed<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, g_smem_optin));
    g_device = device;
    return 0;
}

void *ngb_dev_malloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 8) != cudaSuccess) return nullptr;
    cudaMemsetAsync(p, 0, bytes ? bytes : 8, g_stream);
    return p;
}
void ngb_dev_free(void *p) { if (p) cudaFree(p); }
int ngb_dev_h2d(void *dst, const void *src, size_t bytes)
{
    CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_stream));
    CUDA_OK(cudaStreamSynchronize(g_stream));
    return 0;
}
int ngb_dev_d2h(void *dst, const void *src, size_t bytes)
{
    CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_stream));
    CUDA_OK(cudaStreamSynchronize(g_stream));
    return 0;
}
int ngb_dev_memset(void *dst, int byte, size_t bytes)
{
    CUDA_OK(cudaMemsetAsync(dst, byte, bytes, g_stream));
    return 0;
}
int ngb_dev_sync(void)
{
    CUDA_OK(cudaStreamSynchronize(g_stream));
    CUDA_OK(cudaGetLastError());
    return 0;
}
long ngb_dev_launch_count(void) { return g_launches; }
void *ngb_dev_stream(void) { return (void *)g_stream; }

/* live timing of the dominant kernel: CUDA events around sampled bsim4_load launches */
#define NGB_PROF_MAX 2048
static thread_local int g_prof_on = 0, g_prof_every = 1, g_prof_n = 0;
static thread_local long g_prof_seen = 0;
static thread_local cudaEvent_t g_prof_ev[2 * NGB_PROF_MAX];
static thread_local int g_prof_created = 0;

void ngb_dev_profile(int enable, int every)
{
    if (enable && !g_prof_created) {
        for (int i = 0; i < 2 * NGB_PROF_MAX; i++) cudaEventCreate(&g_prof_ev[i]);
        g_prof_created = 1;
    }
    g_prof_on = enable; g_prof_every = every > 0 ? every : 1; g_prof_n = 0; g_prof_seen = 0;
}
int ngb_dev_profile_read(double *ms_sum, long *count)
{
    double sum = 0.0;
    if (g_stream) cudaStreamSynchronize(g_stream);
    for (int i = 0; i < g_prof_n; i++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_prof_ev[2 * i], g_prof_ev[2 * i + 1]) == cudaSuccess) sum += ms;
    }
    if (ms_sum) *ms_sum = sum;
    if (count) *count = g_prof_n;
    g_prof_n = 0; g_prof_seen = 0;
    return 0;
}
int ngb_dev_set_stream(void *stream)
{
    if (g_device < 0) return NGB_E_PANIC;
    g_stream = (cudaStream_t)stream; g_cur = g_stream;
    return 0;
}

static thread_local int g_capturing = 0, g_prof_force = 0;

/* per-stage timing of sampled Newton steps (the steps ngb_dev_profile_due sends kernel by kernel, on one stream): an
 * event after every stage; the time between two marks is booked on the later one */
#define NGB_STAGE_SLOTS 8
#define NGB_STAGE_TICKS 256
static thread_local cudaEvent_t g_stage_ev[NGB_STAGE_TICKS][NGB_STAGE_SLOTS];
static thread_local unsigned g_stage_mask[NGB_STAGE_TICKS];
static thread_local int g_stage_created = 0, g_stage_cur = -1, g_stage_n = 0;
void ngb_dev_stage_begin(void)
{
    g_stage_cur = -1;
    if (!g_prof_on || !g_prof_force || g_capturing || g_stage_n >= NGB_STAGE_TICKS) return;
    if (!g_stage_created) {
        for (int i = 0; i < NGB_STAGE_TICKS; i++) for (int k = 0; k < NGB_STAGE_SLOTS; k++) cudaEventCreate(&g_stage_ev[i][k]);
        g_stage_created = 1;
    }
    g_stage_cur = g_stage_n++;
    g_stage_mask[g_stage_cur] = 1u;
    cudaEventRecord(g_stage_ev[g_stage_cur][0], g_stream);
}
void ngb_dev_stage_mark(int slot)
{
    if (g_stage_cur < 0 || slot <= 0 || slot >= NGB_STAGE_SLOTS) return;
    cudaEventRecord(g_stage_ev[g_stage_cur][slot], g_stream);
    g_stage_mask[g_stage_cur] |= 1u << slot;
}
/* ms[slot] summed over the sampled steps; returns their number */
int ngb_dev_stage_read(double ms[NGB_STAGE_SLOTS])
{
    if (g_stream) cudaStreamSynchronize(g_stream);
    for (int k = 0; k < NGB_STAGE_SLOTS; k++) ms[k] = 0.0;
    for (int i = 0; i < g_stage_n; i++) {
        int prev = 0;
        for (int k = 1; k < NGB_STAGE_SLOTS; k++) {
            if (!(g_stage_mask[i] >> k & 1u)) continue;
            float t = 0.f;
            if (cudaEventElapsedTime(&t, g_stage_ev[i][prev], g_stage_ev[i][k]) == cudaSuccess) ms[k] += t;
            prev = k;
        }
    }
    const int n = g_stage_n;
    g_stage_n = 0; g_stage_cur = -1;
    return n;
}
int ngb_dev_profile_due(void)
{
    if (!g_prof_on || g_prof_n >= NGB_PROF_MAX) return 0;
    if (g_prof_seen++ % g_prof_every == 0) { g_prof_force = 1; return 1; }
    return 0;
}
int ngb_dev_graph_begin(void)
{
    if (g_capturing) return NGB_E_PANIC;
    if (cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); return NGB_E_PANIC; }
    g_capturing = 1;
    return 0;
}
int ngb_dev_graph_end(void **exec, int *nodes)
{
    cudaGraph_t g = NULL; cudaGraphExec_t x = NULL; size_t n = 0;
    g_capturing = 0;
    if (cudaStreamEndCapture(g_stream, &g) != cudaSuccess || !g) { cudaGetLastError(); return NGB_E_PANIC; }
    cudaGraphGetNodes(g, NULL, &n);
    if (cudaGraphInstantiate(&x, g, 0) != cudaSuccess) { cudaGetLastError(); cudaGraphDestroy(g); return NGB_E_PANIC; }
    cudaGraphDestroy(g);
    *exec = (void *)x; *nodes = (int)n;
    return 0;
}
int ngb_dev_graph_launch(void *exec, int nodes)
{
    CUDA_OK(cudaGraphLaunch((cudaGraphExec_t)exec, g_stream));
    g_launches += nodes;
    return 0;
}
void ngb_dev_graph_destroy(void *exec) { if (exec) cudaGraphExecDestroy((cudaGraphExec_t)exec); }

/* L2 persistence for a block that every launch re-reads (per-sample parameter rows): an access-policy window on the launch
 * streams, inherited by the kernel nodes of captured graphs.  NGB_L2_PERSIST=0 switches it off */
void ngb_dev_l2_persist(const void *p, size_t bytes)
{
    static int on = -1;
    if (on < 0) { const char *e = getenv("NGB_L2_PERSIST"); on = (e && !atoi(e)) ? 0 : 1; }
    if (!on || !g_stream) return;
    int dev = 0, maxwin = 0, maxpersist = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&maxwin, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    cudaDeviceGetAttribute(&maxpersist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaStreamAttrValue av;
    memset(&av, 0, sizeof av);
    if (p && bytes && maxwin > 0 && maxpersist > 0) {
        size_t win = bytes < (size_t)maxwin ? bytes : (size_t)maxwin;
        size_t carve = win < (size_t)maxpersist ? win : (size_t)maxpersist;
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
        av.accessPolicyWindow.base_ptr = const_cast<void *>(p);
        av.accessPolicyWindow.num_bytes = win;
        av.accessPolicyWindow.hitRatio = (float)((double)carve / (double)win);
        av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    } else {
        av.accessPolicyWindow.num_bytes = 0;
        cudaCtxResetPersistingL2Cache();
    }
    cudaStreamSetAttribute(g_stream, cudaStreamAttributeAccessPolicyWindow, &av);
    for (int i = 0; i < NGB_SIDE; i++) if (g_side_ready) cudaStreamSetAttribute(g_side[i], cudaStreamAttributeAccessPolicyWindow, &av);
    cudaGetLastError();
}

/* fork / join around the device-type loads of one CKTload.  A step whose bsim4_load is being timed
 * (ngb_dev_profile_due) stays on one stream so that the kernel is measured alone. */
int ngb_dev_branch_begin(void)
{
    if (g_branch_off < 0) { const char *e = getenv("NGB_NO_BRANCH"); g_branch_off = (e && atoi(e)) ? 1 : 0; }
    g_branching = 0;
    if (g_branch_off || !g_side_ready || g_prof_force) return 0;
    if (cudaEventRecord(g_fork, g_stream) != cudaSuccess) { cudaGetLastError(); return 0; }
    for (int i = 0; i < NGB_SIDE; i++) g_side_used[i] = 0;
    g_branching = 1;
    return 1;
}
void ngb_dev_branch(int i)
{
    if (!g_branching || i < 0) { g_cur = g_stream; return; }
    i %= NGB_SIDE;
    if (!g_side_used[i]) { cudaStreamWaitEvent(g_side[i], g_fork, 0); g_side_used[i] = 1; }
    g_cur = g_side[i];
}
int ngb_dev_branch_end(void)
{
    g_cur = g_stream;
    if (!g_branching) return 0;
    g_branching = 0;
    for (int i = 0; i < NGB_SIDE; i++)
        if (g_side_used[i]) {
            CUDA_OK(cudaEventRecord(g_join[i], g_side[i]));
            CUDA_OK(cudaStreamWaitEvent(g_stream, g_join[i], 0));
        }
    return 0;
}


/* out[0] = DFMA flop/s (2 per instruction), out[1] = DADD/DMUL flop/s (1 per instruction), out[2] = SM clock (kHz) the
 * driver reports; best of `reps` launches of 148 x 8 CTAs x 256 threads x 8 chains */
int ngb_dev_fp64_peak(double out[3])
{
    if (!g_stream) { ngb_set_error("ngbInit first"); return NGB_E_PANIC; }
    double *d = nullptr;
    CUDA_OK(cudaMalloc(&d, 8));
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
    const int iters = 1 << 14, grid = g_sm_count * 8, reps = 5;
    for (int fused = 1; fused >= 0; fused--) {
        double best = 0.0;
        for (int r = 0; r < reps + 1; r++) {
            cudaEventRecord(e0, g_stream);
            if (fused) ngb_k_fp64_peak<1><<<grid, 256, 0, g_stream>>>(d, iters, 1.0 + r);
            else ngb_k_fp64_peak<0><<<grid, 256, 0, g_stream>>>(d, iters, 1.0 + r);
            cudaEventRecord(e1, g_stream);
            CUDA_OK(cudaEventSynchronize(e1));
            float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1);
            const double rate = (double)grid * 256 * 8 * iters * (fused ? 2.0 : 1.0) / (ms * 1e-3);
            if (r > 0 && rate > best) best = rate;      /* launch 0 is the warm-up */
        }
        out[fused ? 0 : 1] = best;
    }
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, g_device);
    out[2] = khz;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    g_launches += 2 * (reps + 1);
    return 0;
}

static int post_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (!g_capturing) g_launches++;
    if (e != cudaSuccess) { ngb_set_error("launch of %s failed: %s", what, cudaGetErrorString(e)); return NGB_E_PANIC; }
    return 0;
}

int ngb_launch_bsim4_load(const B4Ctx *c, int *errflag)
{
    if (c->T <= 0) return 0;
    /* the kernel is bound by per-warp latency, not throughput: a batch that does not fill the GPU with 256-thread CTAs
     * (fewer than two per SM) is spread over all SMs in smaller ones -- fewer warps share an SM's L1 and instruction cache */
    int cta = NGB_B4_CTA;
    while (cta > 64 && ((size_t)c->T + cta - 1) / cta < 2 * (size_t)g_sm_count) cta >>= 1;
    const unsigned grid = (unsigned)(((size_t)c->T + cta - 1) / cta);
    /* sampled timing: forced by ngb_dev_profile_due (transient driver) or by this launch's own turn */
    const int rec = !g_capturing && g_prof_on && g_prof_n < NGB_PROF_MAX && (g_prof_force || (g_prof_seen++ % g_prof_every == 0));
    g_prof_force = 0;
    if (rec) cudaEventRecord(g_prof_ev[2 * g_prof_n], g_cur);
    {
        int launched = 0;
#define X(k) if (!launched && c->variant == (k)) { ngb_k_bsim4_load<(k)><<<grid, cta, 0, g_cur>>>(*c, errflag); launched = 1; }
        NGB_B4_VARIANT_KEYS(X)
#undef X
        if (!launched) ngb_k_bsim4_load<NGB_B4_GENERIC><<<grid, cta, 0, g_cur>>>(*c, errflag);
    }
    if (rec) { cudaEventRecord(g_prof_ev[2 * g_prof_n + 1], g_cur); g_prof_n++; }
    return post_launch("bsim4_load");
}
int ngb_launch_bsim4_lte(const B4Ctx *c)
{
    if (c->T <= 0) return 0;
    static int flat = -1;
    if (flat < 0) { const char *e = getenv("NGB_LTE_FLAT"); flat = (e && !atoi(e)) ? 0 : 1; }
    if (flat && c->S >= 32)
        ngb_launch_dep(1, ngb_k_bsim4_lte_flat, (unsigned)(((size_t)c->T + 255) / 256), 256u, 0, g_stream, *c);
    else
        ngb_k_bsim4_lte<<<(unsigned)(((size_t)c->S * 32 + 127) / 128), 128, 0, g_stream>>>(*c);
    return post_launch("bsim4_lte");
}
int ngb_launch_cap_load(const NgbCapCtx *c, int *errflag)
{
    if (c->T <= 0) return 0;
    const unsigned grid = (unsigned)(((size_t)c->T + 255) / 256);
    ngb_k_cap_load<<<grid, 256, 0, g_cur>>>(*c, errflag);
    return post_launch("cap_load");
}
int ngb_launch_bsim3_load(const B3Ctx *c, int *errflag)
{
    if (c->T <= 0) return 0;
    const unsigned grid = (unsigned)(((size_t)c->T + 255) / 256);
    ngb_k_bsim3_load<<<grid, 256, 0, g_cur>>>(*c, errflag);
    return post_launch("bsim3_load");
}
int ngb_launch_vbic_load(const NgbVbicCtx *c, int *errflag)
{
    if (c->T <= 0) return 0;
    const unsigned grid = (unsigned)(((size_t)c->T + 127) / 128);
    ngb_k_vbic_load<<<grid, 128, 0, g_cur>>>(*c, errflag);
    return post_launch("vbic_load");
}
int ngb_launch_dio_load(const NgbDioCtx *c, int *errflag)
{
    if (c->T <= 0) return 0;
    const unsigned grid = (unsigned)(((size_t)c->T + 255) / 256);
    ngb_k_dio_load<<<grid, 256, 0, g_cur>>>(*c, errflag);
    return post_launch("dio_load");
}
int ngb_launch_src_load(const NgbSrcCtx *c)
{
    if (c->T <= 0) return 0;
    const unsigned grid = (unsigned)(((size_t)c->T + 255) / 256);
    ngb_k_src_load<<<grid, 256, 0, g_cur>>>(*c);
    return post_launch("src_load");
}
__global__ void __launch_bounds__(128)
ngb_k_node_override(const NgbAsmCtx c)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < c.S) ngb_override_thread(&c, s);
}

int ngb_launch_assemble(const NgbAsmCtx *c)
{
    const size_t total = (size_t)(c->nnz + c->neq1) * c->S;
    const unsigned grid = (unsigned)((total + 255) / 256);
    static int tiled = -1;
    if (tiled < 0) { const char *e = getenv("NGB_ASM_TILED"); tiled = (e && !atoi(e)) ? 0 : 1; }
    if (tiled && c->S >= 32) {
        const int nts = (c->S + 31) / 32, ntt = (c->nnz + c->neq1 + 31) / 32;
        ngb_launch_dep(2, ngb_k_assemble_tiled, (unsigned)nts * (unsigned)ntt, 256u, 0, g_stream, *c, nts);
    } else {
        ngb_k_assemble<<<grid, 256, 0, g_stream>>>(*c, total);
    }
    for (int li = 0; li < c->nlong; li++) {
        const int e = post_launch("assemble");
        if (e) return e;
        int n = (c->long_len_host[li] + NGB_ASM_CHUNK - 1) / NGB_ASM_CHUNK, half = 0;
        if (!c->long_part || (size_t)n * c->S > (size_t)c->long_cap) { ngb_set_error("long-target scratch too small"); return NGB_E_PANIC; }
        ngb_k_assemble_long1<<<(unsigned)(((size_t)n * c->S + 255) / 256), 256, 0, g_stream>>>(*c, li, c->long_part, n);
        while (n > 1 || half == 0) {           /* at least one second-level pass: it is the one that stores the total */
            const int m = (n + NGB_ASM_CHUNK - 1) / NGB_ASM_CHUNK;
            const int e2 = post_launch("assemble_long");
            if (e2) return e2;
            ngb_k_assemble_long2<<<(unsigned)(((size_t)m * c->S + 255) / 256), 256, 0, g_stream>>>(*c, li, c->long_part + (size_t)half * c->long_cap,
                                                                                                 c->long_part + (size_t)(1 - half) * c->long_cap, n, m);
            n = m; half = 1 - half;
            if (m == 1) break;
        }
    }
    if (c->nov > 0) {
        const int e = post_launch("assemble");
        if (e) return e;
        ngb_k_node_override<<<(unsigned)((c->S + 127) / 128), 128, 0, g_stream>>>(*c);
        return post_launch("node_override");
    }
    return post_launch("assemble");
}
int ngb_launch_lu(const NgbLuCtx *c)
{
    const int per = c->sch.nV + c->sch.n + c->sch.ntask;
    const size_t bytes1 = (size_t)per * sizeof(double);
    if (c->pk.ok) {
        const int per = c->sch.nV + c->sch.n + c->sch.ntask + c->pk.maxlp;
        const size_t bytes1 = (size_t)per * sizeof(double);
        /* schedule blob + per-sample values in shared memory.  Many samples: one warp each, one CTA
         * per SM holding as many samples as its shared memory takes (the blob is paid once per CTA);
         * few samples: one CTA of 256 threads per sample. */
        static int v1 = -1;
        if (v1 < 0) { const char *e = getenv("NGB_LU_V1"); v1 = (e && atoi(e)) ? 1 : 0; }   /* first packing, for comparison */
        const int v2 = c->pk.ok2 && !v1;
        const size_t blob = (size_t)(((v2 ? c->pk.blob2_u16 : c->pk.blob_u16) + 3) / 4) * sizeof(double);
        const size_t budget = (size_t)g_smem_optin - 1024;
        if (c->S >= 64 && blob + 4 * bytes1 <= budget) {
            int groups = (int)((budget - blob) / bytes1);
            if (groups > 32) groups = 32;
            {   /* one CTA per SM and wave: the fewest waves the shared memory allows, then the smallest group
                 * count that still needs only that many, so that the last wave is as full as the first */
                const int waves = (c->S + groups * g_sm_count - 1) / (groups * g_sm_count);
                const int even = (c->S + waves * g_sm_count - 1) / (waves * g_sm_count);
                if (even >= 4 && even < groups) groups = even;
            }
            const unsigned grid = (unsigned)((c->S + groups - 1) / groups);
            static int tpg = 0;
            if (!tpg) { const char *e = getenv("NGB_LU_TPG"); tpg = e ? atoi(e) : 32; if (tpg != 4 && tpg != 8 && tpg != 16) tpg = 32; }
            if (v2) ngb_launch_dep(1, ngb_k_lu_packed<1>, grid, (unsigned)((groups * tpg + 31) / 32 * 32), blob + bytes1 * groups, g_stream, *c, groups, tpg, per);
            else ngb_launch_dep(1, ngb_k_lu_packed<0>, grid, (unsigned)((groups * tpg + 31) / 32 * 32), blob + bytes1 * groups, g_stream, *c, groups, tpg, per);
            return post_launch("lu_packed");
        }
        if (blob + bytes1 <= (size_t)g_smem_optin) {
            /* few samples: one CTA per sample.  The levels of a circuit matrix are narrow (RO-101: 27 values and 19 products
             * per level on average), so the width that pays is small: NGB_LU_CTA threads (default below), one warp
             * synchronises with __syncwarp instead of a CTA barrier */
            static int cta = 0;
            if (!cta) { const char *e = getenv("NGB_LU_CTA"); cta = e ? atoi(e) : NGB_LU_CTA_DEFAULT; if (cta < 32 || cta > 1024 || (cta & 31)) cta = NGB_LU_CTA_DEFAULT; }
            if (v2) ngb_k_lu_packed<1><<<(unsigned)c->S, cta, blob + bytes1, g_stream>>>(*c, 1, cta, per);
            else ngb_k_lu_packed<0><<<(unsigned)c->S, cta, blob + bytes1, g_stream>>>(*c, 1, cta, per);
            return post_launch("lu_packed");
        }
    }
    /* warp per sample while four samples fit one CTA's shared memory with room for 2+ CTAs/SM */
    if (bytes1 * 4 <= (size_t)g_smem_optin / 2) {
        const int warps = 4;
        const unsigned grid = (unsigned)((c->S + warps - 1) / warps);
        ngb_k_lu_warp<<<grid, warps * 32, bytes1 * warps, g_stream>>>(*c, per);
    } else if (bytes1 <= (size_t)g_smem_optin) {
        ngb_k_lu_block<<<(unsigned)c->S, 256, bytes1, g_stream>>>(*c);
    } else {
        /* grid-wide LU: one CTA per SM (every CTA must be resident for the barrier; the levels of a circuit matrix are
         * narrower than 148 x 256 lanes anyway) */
        NgbLuCtx cc = *c;
        void *args[1] = { (void *)&cc };
        if (!c->gV || !c->gRs || !c->gZ) { ngb_set_error("grid-wide LU without its work arrays"); return NGB_E_PANIC; }
        if (cudaLaunchCooperativeKernel((const void *)ngb_k_lu_grid, dim3((unsigned)g_sm_count), dim3(256), args, 0, g_stream) != cudaSuccess) {
            ngb_set_error("cooperative launch of the grid-wide LU failed: %s", cudaGetErrorString(cudaGetLastError()));
            return NGB_E_PANIC;
        }
    }
    return post_launch("lu");
}
int ngb_launch_clear_i32(int *p, int value, int n)
{
    if (n <= 0) return 0;
    ngb_k_clear_i32<<<(unsigned)((n + 255) / 256), 256, 0, g_stream>>>(p, value, n);
    return post_launch("clear_i32");
}

int ngb_launch_tran_control(const NgbTranCtx *c)
{
    ngb_launch_dep(1, ngb_k_tran_control, (unsigned)((c->S + 127) / 128), 128u, 0, g_stream, *c);
    return post_launch("tran_control");
}
int ngb_launch_fill_f64(double *p, double value, int n)
{
    if (n <= 0) return 0;
    ngb_k_fill_f64<<<(unsigned)((n + 255) / 256), 256, 0, g_stream>>>(p, value, n);
    return post_launch("fill_f64");
}

}  /* extern "C" */
