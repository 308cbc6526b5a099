/* ngb_host.h -- host-side objects behind the opaque handles of include/ngb200.h */
#ifndef NGB_HOST_H
#define NGB_HOST_H
#include "ngb_types.h"
#include "bsim4_eval.cuh"
#include "dio_eval.cuh"
#include "bsim3_eval.cuh"
#include "vbic_types.h"

#ifdef __cplusplus
extern "C" {
#endif

struct ngb_circuit {
    int neq;                       /* CKTmaxEqNum */
    int *node_type;                /* [neq+1] */
    NgbOpts opt;
    int finalized, have_lu, exact_order;
    /* device tables (host copies, instance order = reference list order) */
    int b4_n, b4_nrows; int *b4_nodes, *b4_flags, *b4_prow; double *b4_inst, *b4_mtab, *b4_ptab;
    int *b4_spos, *b4_slots;
    int b4_row0;                   /* first stamp row of the BSIM4 block: position k of instance i is row b4_row0 + k * b4_n + i */
    int res_n; int *res_nodes; double *res_g; int *res_spos;
    int cap_n; int *cap_nodes; double *cap_par; int *cap_spos;
    int b3_n, b3_nrows; int *b3_nodes, *b3_flags, *b3_prow; double *b3_inst, *b3_mtab, *b3_ptab; int *b3_spos;
    int vb_n; int *vb_nodes, *vb_flags; double *vb_par, *vb_aux; int *vb_spos;
    int dio_n; int *dio_nodes, *dio_flags; double *dio_par; int *dio_spos;
    int vs_n; int *vs_nodes, *vs_fn; double *vs_par; int *vs_spos, *vs_cspos;
    int *vs_pwl_ptr, *vs_pwl_rep, *vs_pwl_len, vs_pwl_n; double *vs_pwl, *vs_pwl_rdelay;
    int *is_pwl_ptr, *is_pwl_len, is_pwl_n; double *is_pwl;      /* PWL current sources: corner lists only (isrcload.c has no delay / repetition) */
    int is_n; int *is_nodes, *is_fn; double *is_par; int *is_spos;
    /* CSC pattern (SMPconvertCOOtoCSC) */
    int n, nnz; int *Ap, *Ai, *eq2col, *col2eq, *slot_diag, *diag_slot;
    /* stamp rows and per-target contribution lists */
    int nstamp_rows, ntgt; int *tgt_ptr, *tgt_rows;
    int nlong; int *long_tgt;     /* targets with more than NGB_ASM_LONG contributions */
    int nconst; int *const_row; double *const_val;
    /* .nodeset / .ic rows (ngbCircuitSetNodeOverrides) */
    int ov_n; int *ov_eq, *ov_kind, *ov_cur, *ov_diag, *ov_zptr, *ov_zslot; double *ov_val;
    /* LU: imported or own symbolic objects + task schedule */
    int klu_nblocks; int *klu_Q, *klu_R, *klu_Pnum;
    int *klu_P;                    /* symbolic row permutation (ngbCircuitSetSymbolic / ngbCircuitAnalyze), NULL when only finished factors were imported */
    int *pat_Lp, *pat_Li, *pat_Up, *pat_Ui, *pat_Offp, *pat_Offi;    /* the factor last passed to ngbCircuitSetLuPattern */
    double pivtol;                 /* threshold of the own pivoting factor (CKTpivotRelTol) */
    int pivot_mode;                /* ngbCircuitSetPivotMode: -1 default (1 when a symbolic analysis is there), 0 batch orders, 1 per sample */
    int lnz, unz, nzoff, npairs, nsolvepairs;
    NgbLuSched sch;                /* host arrays (set being built) */
    NgbLuPacked pk;                /* host arrays, level-contiguous 16-bit form */
    /* finished pattern sets, one per distinct pivoting factor of a run (NIiter re-pivots in the INITJCT
     * iteration, in the iteration after it, and in the first two iterations of the first time point); lu_target selects which one ngbCircuitSetLuPattern fills */
    struct ngb_luset { NgbLuSched sch; NgbLuPacked pk; int npairs, nsolvepairs, lnz, unz, nzoff, valid;
                      unsigned long long sig; /* hash of the factor (row order + L / U patterns): equal factors share a set */ } lu[NGB_LU_SETS];
    int lu_target;
    int lu_event[NGB_LU_EVENTS], lu_event_set;   /* pivoting event -> pattern set (ngbCircuitSetLuEvents) */
};

#define NGB_MAX_ARR 64
struct ngb_tran;
struct ngb_batch {
    struct ngb_circuit *c;
    int S, neq1, failed, op_full, have_lu;
    NgbCtl ctl;
    double *x, *Ax, *stamp;
    int *errflag;
    int *d_node_type, *d_tgt_ptr, *d_tgt_rows, *d_slot_diag, *d_long_tgt;
    int *long_len; double *long_part; int long_cap;      /* long assembly targets: lengths (host), chunk-total scratch (device) */
    int *lu_verify;                   /* [S] device, owned by the transient driver: pivoting event due (NgbLuCtx.verify) */
    double *Zw; int lu_ntask_cap;     /* solve-task scratch of the grid-wide LU, [S][lu_ntask_cap] */
    int lte_deferred;                 /* transient driver: BSIM4trunc in its own launch after the solve */
    double *b4_inst, *b4_state, *b4_op, *b4_mtab, *b4_ptab; int *b4_prow, *b4_prow_t, *b4_flags, *b4_nodes, *b4_spos;
    void *b4_rows_block;              /* per-sample rows (ngbBatchSetBsim4Rows): b4_mtab and b4_ptab live in this one allocation */
    double *cap_par, *cap_state; int *cap_nodes, *cap_spos;
    double *b3_inst, *b3_state, *b3_von, *b3_mtab, *b3_ptab; int *b3_prow, *b3_flags, *b3_nodes, *b3_spos;
    double *vb_par, *vb_aux, *vb_state; int *vb_nodes, *vb_flags, *vb_spos;
    int *ov_eq, *ov_kind, *ov_cur, *ov_diag, *ov_zptr, *ov_zslot; double *ov_val;
    double *dio_par, *dio_state; int *dio_nodes, *dio_flags, *dio_spos;
    double *vs_par; int *vs_fn, *vs_spos; int *vs_pwl_ptr, *vs_pwl_rep; double *vs_pwl, *vs_pwl_rdelay; int *is_pwl_ptr; double *is_pwl;
    double *is_par; int *is_fn, *is_spos;
    struct { NgbLuSched dsch; NgbLuPacked dpk; int valid; } dlu[NGB_LU_SETS];   /* device arrays per pattern set */
    int lu_which;                  /* set used by the direct ngbLuFac/ngbSolve calls */
    double *V, *Rs; int *nodeconv, *singular;
    struct { const char *name; void *ptr; size_t bytes; } arr[NGB_MAX_ARR];
    int narr;
    struct ngb_tran *tran;
    int load_lte;              /* ngbBatchSetLoadLte: direct ngbLoad calls evaluate DEVtrunc's bounds inside the load (off by default) */
    unsigned b4_key;           /* variant key of the BSIM4 instances (bsim4_variants.h); NGB_B4_GENERIC when they differ */
    int b4_overlay;            /* per-sample rows are read as an overlay over the rows of sample 0 (ngbBatchSetBsim4Rows) */
    int b4_mvary[B4M_COUNT], b4_pvary[B4P_COUNT];   /* byte pitch of a row for the columns that differ between samples, else 0 */
    int b4_force_generic;      /* ngbBatchSetBsim4Generic / NGB_B4_GENERIC=1: run the generic kernel whatever the key */
    /* measurement clauses for the next ngbTranRun (ngbTranSetMeasures) */
    int ms_n; int *ms_eq, *ms_kind, *ms_count; double *ms_val, *ms_td;
};

void ngb_set_error(const char *fmt, ...);
void ngb_fill_b4ctx(struct ngb_batch *b, B4Ctx *x);
void ngb_fill_capctx(struct ngb_batch *b, NgbCapCtx *x);
void ngb_fill_dioctx(struct ngb_batch *b, NgbDioCtx *x);
void ngb_fill_b3ctx(struct ngb_batch *b, B3Ctx *x);
void ngb_fill_vbctx(struct ngb_batch *b, NgbVbicCtx *x);
void ngb_lu_events(const struct ngb_circuit *c, int ev[NGB_LU_EVENTS]);
void ngb_fill_srcctx(struct ngb_batch *b, NgbSrcCtx *x, int is_current);
void ngb_fill_asmctx(struct ngb_batch *b, NgbAsmCtx *x);
void ngb_fill_luctx(struct ngb_batch *b, NgbLuCtx *x, int do_factor, int do_solve, int which);
int ngb_enqueue_load(struct ngb_batch *b);
/* own pivoting factor of one sample's matrix (ngb_repivot_compute: no shared state, may run on many host threads) and its
 * placement among the circuit's pattern sets (ngb_repivot_commit: serial) */
typedef struct NgbRepivot { int rc, sing; int *Pnum, *Lp, *Up, *Offp, *Li, *Ui, *Offi; } NgbRepivot;
void ngb_repivot_compute(const struct ngb_circuit *c, const double *Ax, NgbRepivot *r);
int ngb_repivot_commit(struct ngb_batch *b, NgbRepivot *r, int *set_out);
void ngb_repivot_free(NgbRepivot *r);
void ngb_tran_free(struct ngb_batch *b);
int ngb_pivot_factor(int n, const int *Ap, const int *Ai, const double *Ax, int nblocks, const int *P, const int *Q,
                     const int *R, double tol, int *Pnum, int *Lp, int **Li_out, int *Up, int **Ui_out,
                     int *Offp, int **Offi_out, int *singular_col);
int ngbBatchSetBsim4Rows(struct ngb_batch *b, const int *prow_t, int nrows, const double *mtab, const double *ptab);

#ifdef __cplusplus
}
#endif
#endif
