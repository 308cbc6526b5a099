/* ngb_pivot.c -- pivoting numeric factor of one sample's matrix on the host: the role klu_factor plays behind
 * SMPreorder (src/maths/KLU/klusmp.c:700-760 -> klu_factor.c:384 -> klu_kernel.c:642).
 *
 * Why it exists here: the batched refactor on the device keeps ONE recorded pivot order per pattern set.  When a
 * sample's refactor meets an exact zero pivot the reference factors that same matrix again from scratch, with
 * pivoting, inside the same Newton iteration (niiter.c:162-195); ngb_tran.c does the same for the sample through this
 * file, and callers without a recorded run (bench.py, ngbCircuitFactor) obtain their pattern sets from it.
 *
 * What is reused and what is restated: the symbolic analysis (block triangular form P, Q, R from klu_analyze) is
 * an INPUT.  The numeric part is written here from the published algorithm:
 * left-looking Gilbert-Peierls -- for column k, the pattern of L \ A(:,k) by a depth-first search over the
 * columns of L already computed, the values by a sparse forward substitution in the search's topological order --
 * with Eisenstat-Liu symmetric pruning of the search and threshold partial pivoting that keeps the diagonal when
 * |a_kk| >= tol * max|a_ik|.  The ORDER in which row indices end up inside every column of L and U decides the order
 * of the subtractions of every later refactor (klu_refactor.c:285-426), i.e. the bits of the solution; the search
 * order, the pruning partition and the pivot swap therefore follow klu_kernel.c step for step (dfs :17-113,
 * lsolve_symbolic :117-300, lsolve_numeric :310-345, lpivot :355-520, prune :529-640), and the result is pinned
 * against every pivoting factor the reference recorded (tests/test_pivot_factor.py).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "ngb_host.h"

#define NP_EMPTY (-1)
#define NP_FLIP(i) (-(i) - 2)

typedef struct {            /* one column of L or U under construction */
    int *i; double *x; int len, cap;
} NpCol;

static int col_reserve(NpCol *c, int need)
{
    if (need <= c->cap) return 0;
    int cap = c->cap ? c->cap : 8;
    while (cap < need) cap *= 2;
    int *ni = (int *)realloc(c->i, sizeof(int) * (size_t)cap);
    if (!ni) return -1;
    c->i = ni;
    double *nx = (double *)realloc(c->x, sizeof(double) * (size_t)cap);
    if (!nx) return -1;
    c->x = nx; c->cap = cap;
    return 0;
}

/* depth-first search from pivotal row j through the (pruned) columns of L; non-pivotal rows met on the way go
 * straight into the new column of L, finished pivotal rows onto the output stack from the top down */
static int np_dfs(int j, int k, const int *Pinv, NpCol *L, const int *Lpend, int *Stack, int *Flag, int *Ap_pos,
                  int top, NpCol *Lk)
{
    int head = 0;
    Stack[0] = j;
    while (head >= 0) {
        j = Stack[head];
        const int jnew = Pinv[j];
        if (Flag[j] != k) {
            Flag[j] = k;
            Ap_pos[head] = (Lpend[jnew] == NP_EMPTY) ? L[jnew].len : Lpend[jnew];
        }
        const int *Li = L[jnew].i;
        int pos;
        for (pos = --Ap_pos[head]; pos >= 0; --pos) {
            const int i = Li[pos];
            if (Flag[i] != k) {
                if (Pinv[i] >= 0) {
                    Ap_pos[head] = pos;
                    Stack[++head] = i;
                    break;
                }
                Flag[i] = k;
                Lk->i[Lk->len++] = i;
            }
        }
        if (pos == -1) {
            head--;
            Stack[--top] = j;
        }
    }
    return top;
}

/* Factor the matrix (Ap, Ai, Ax: CSC, order n) on the block structure P, Q, R (nblocks blocks).  tol is the pivot
 * threshold (CKTpivotRelTol through SMPreorder; KLU's default 0.001).  On success fills the caller's arrays:
 * Pnum[n], Lp[n+1], Up[n+1], Offp[n+1] and allocates *Li, *Ui, *Offi (caller frees); L / U row indices are in
 * pivotal numbering of the whole matrix, in the order the factorization produced them.
 * Returns 0, or NGB_E_SINGULAR with *singular_col = the column (original numbering) whose pivot is zero
 * (halt_if_singular, klu_defaults.c:32), or NGB_E_PANIC when memory runs out. */
int ngb_pivot_factor(int n, const int *Ap, const int *Ai, const double *Ax, int nblocks, const int *P, const int *Q,
                     const int *R, double tol, int *Pnum, int *Lp, int **Li_out, int *Up, int **Ui_out,
                     int *Offp, int **Offi_out, int *singular_col)
{
    int rc = NGB_OK, b, k, p;
    const int nz = Ap[n];
    double *Rs = (double *)calloc((size_t)n, sizeof(double));
    int *PSinv = (int *)malloc(sizeof(int) * (size_t)n);
    int *Offi = (int *)malloc(sizeof(int) * (size_t)(nz > 0 ? nz : 1));
    double *X = (double *)calloc((size_t)n, sizeof(double));
    int *Pinv = (int *)malloc(sizeof(int) * (size_t)n), *Pblk = (int *)malloc(sizeof(int) * (size_t)n);
    int *Stack = (int *)malloc(sizeof(int) * (size_t)n), *Flag = (int *)malloc(sizeof(int) * (size_t)n);
    int *Ap_pos = (int *)malloc(sizeof(int) * (size_t)n), *Lpend = (int *)malloc(sizeof(int) * (size_t)n);
    NpCol *L = (NpCol *)calloc((size_t)n, sizeof(NpCol)), *U = (NpCol *)calloc((size_t)n, sizeof(NpCol));
    int *Pfin = NULL;
    if (singular_col) *singular_col = -1;
    *Li_out = *Ui_out = *Offi_out = NULL;
    if (!Rs || !PSinv || !Offi || !X || !Pinv || !Pblk || !Stack || !Flag || !Ap_pos || !Lpend || !L || !U) { rc = NGB_E_PANIC; goto done; }

    /* KLU_scale, scale = 2: largest magnitude of every row, 1 for an empty row (klu_scale.c:83-152) */
    for (p = 0; p < nz; p++) {
        const double a = fabs(Ax[p]);
        if (a > Rs[Ai[p]]) Rs[Ai[p]] = a;
    }
    for (k = 0; k < n; k++) if (Rs[k] == 0.0) Rs[k] = 1.0;
    for (k = 0; k < n; k++) PSinv[P[k]] = k;
    Offp[0] = 0;

    for (b = 0; b < nblocks && rc == NGB_OK; b++) {
        const int k1 = R[b], k2 = R[b + 1], nk = k2 - k1;
        if (nk == 1) {
            /* singleton: the pivot is the entry itself, everything above it belongs to the off-diagonal part */
            int poff = Offp[k1];
            const int oldcol = Q[k1];
            double s = 0.0;
            for (p = Ap[oldcol]; p < Ap[oldcol + 1]; p++) {
                const int oldrow = Ai[p];
                if (PSinv[oldrow] < k1) Offi[poff++] = oldrow;
                else s = Ax[p] / Rs[oldrow];
            }
            if (s == 0.0) { if (singular_col) *singular_col = oldcol; rc = NGB_E_SINGULAR; break; }
            Offp[k1 + 1] = poff;
            Pnum[k1] = P[k1];
            continue;
        }
        /* ---- one block, columns k1 .. k2-1, rows in the block's own numbering 0 .. nk-1 ---- */
        NpCol *Lb = L + k1, *Ub = U + k1;
        int firstrow = 0;
        (void)firstrow;
        for (k = 0; k < nk; k++) { X[k] = 0.0; Flag[k] = NP_EMPTY; Lpend[k] = NP_EMPTY; Pblk[k] = k; Pinv[k] = NP_FLIP(k); }
        for (k = 0; k < nk && rc == NGB_OK; k++) {
            NpCol *Lk = &Lb[k], *Uk = &Ub[k];
            int top = nk, poff = Offp[k + k1];
            const int oldcol = Q[k + k1];
            if (col_reserve(Lk, nk)) { rc = NGB_E_PANIC; break; }
            Lk->len = 0;
            /* pattern of the column (search) and scatter of the scaled entries into X */
            for (p = Ap[oldcol]; p < Ap[oldcol + 1]; p++) {
                const int oldrow = Ai[p];
                const int i = PSinv[oldrow] - k1;
                double aik = Ax[p];
                aik /= Rs[oldrow];
                if (i < 0) { Offi[poff++] = oldrow; continue; }
                if (Flag[i] != k) {
                    if (Pinv[i] >= 0) top = np_dfs(i, k, Pinv, Lb, Lpend, Stack, Flag, Ap_pos, top, Lk);
                    else { Flag[i] = k; Lk->i[Lk->len++] = i; }
                }
                X[i] = aik;
            }
            Offp[k + k1 + 1] = poff;
            /* values: x = L \ A(:,k), columns in the search's topological order */
            for (int s = top; s < nk; s++) {
                const int j = Stack[s];
                const NpCol *Lj = &Lb[Pinv[j]];
                const double xj = X[j];
                for (p = 0; p < Lj->len; p++) X[Lj->i[p]] -= Lj->x[p] * xj;
            }
            /* pivot: the largest candidate, unless the diagonal reaches tol times it */
            const int diagrow = Pblk[k];
            int pivrow;
            double pivot;
            if (Lk->len == 0) { if (singular_col) *singular_col = oldcol; rc = NGB_E_SINGULAR; break; }   /* structurally singular */
            {
                int pdiag = NP_EMPTY, ppivrow = NP_EMPTY;
                double abs_pivot = NP_EMPTY, xabs;
                const int last_row_index = Lk->i[Lk->len - 1];
                Lk->len -= 1;
                for (p = 0; p < Lk->len; p++) {
                    const int i = Lk->i[p];
                    const double x = X[i];
                    X[i] = 0.0;
                    Lk->x[p] = x;
                    xabs = fabs(x);
                    if (i == diagrow) pdiag = p;
                    if (xabs > abs_pivot) { abs_pivot = xabs; ppivrow = p; }
                }
                xabs = fabs(X[last_row_index]);
                if (xabs > abs_pivot) { abs_pivot = xabs; ppivrow = NP_EMPTY; }
                if (last_row_index == diagrow) {
                    if (xabs >= tol * abs_pivot) { abs_pivot = xabs; ppivrow = NP_EMPTY; }
                } else if (pdiag != NP_EMPTY) {
                    xabs = fabs(Lk->x[pdiag]);
                    if (xabs >= tol * abs_pivot) { abs_pivot = xabs; ppivrow = pdiag; }
                }
                if (ppivrow != NP_EMPTY) {
                    pivrow = Lk->i[ppivrow];
                    pivot = Lk->x[ppivrow];
                    Lk->i[ppivrow] = last_row_index;
                    Lk->x[ppivrow] = X[last_row_index];
                } else {
                    pivrow = last_row_index;
                    pivot = X[last_row_index];
                }
                X[last_row_index] = 0.0;
                if (pivot == 0.0) { if (singular_col) *singular_col = oldcol; rc = NGB_E_SINGULAR; break; }
                for (p = 0; p < Lk->len; p++) Lk->x[p] /= pivot;
            }
            /* column of U: the stack, top down, in pivotal numbering */
            if (col_reserve(Uk, nk - top > 0 ? nk - top : 1)) { rc = NGB_E_PANIC; break; }
            Uk->len = nk - top;
            for (p = top; p < nk; p++) {
                const int j = Stack[p];
                Uk->i[p - top] = Pinv[j];
                Uk->x[p - top] = X[j];
                X[j] = 0.0;
            }
            /* log the pivot row; a displaced diagonal becomes the "diagonal" of the column whose row was taken */
            if (pivrow != diagrow && Pinv[diagrow] < 0) {
                const int kbar = NP_FLIP(Pinv[pivrow]);
                Pblk[kbar] = diagrow;
                Pinv[diagrow] = NP_FLIP(kbar);
            }
            Pblk[k] = pivrow;
            Pinv[pivrow] = k;
            /* symmetric pruning: a column j of L that holds the new pivot row and is used by column k of U never
             * needs its non-pivotal rows searched again; its pivotal rows move to the front */
            for (p = 0; p < Uk->len; p++) {
                const int j = Uk->i[p];
                if (Lpend[j] != NP_EMPTY) continue;
                NpCol *Lj = &Lb[j];
                for (int p2 = 0; p2 < Lj->len; p2++) {
                    if (Lj->i[p2] != pivrow) continue;
                    int phead = 0, ptail = Lj->len;
                    while (phead < ptail) {
                        const int i = Lj->i[phead];
                        if (Pinv[i] >= 0) phead++;
                        else {
                            ptail--;
                            Lj->i[phead] = Lj->i[ptail]; Lj->i[ptail] = i;
                            const double x = Lj->x[phead]; Lj->x[phead] = Lj->x[ptail]; Lj->x[ptail] = x;
                        }
                    }
                    Lpend[j] = ptail;
                    break;
                }
            }
        }
        if (rc != NGB_OK) break;
        /* rows of L in pivotal order; the block's row order joins the symbolic one */
        for (k = 0; k < nk; k++) {
            for (p = 0; p < Lb[k].len; p++) Lb[k].i[p] = Pinv[Lb[k].i[p]];
            Pnum[k + k1] = P[Pblk[k] + k1];
        }
    }
    if (rc != NGB_OK) goto done;

    /* flatten: whole-matrix pivotal numbering (block offset added), singletons have empty columns */
    {
        int lnz = 0, unz = 0;
        for (b = 0; b < nblocks; b++)
            for (k = R[b]; k < R[b + 1]; k++) {
                Lp[k] = lnz; Up[k] = unz;
                if (R[b + 1] - R[b] > 1) { lnz += L[k].len; unz += U[k].len; }
            }
        Lp[n] = lnz; Up[n] = unz;
        int *Li = (int *)malloc(sizeof(int) * (size_t)(lnz > 0 ? lnz : 1)), *Ui = (int *)malloc(sizeof(int) * (size_t)(unz > 0 ? unz : 1));
        Pfin = (int *)malloc(sizeof(int) * (size_t)n);
        if (!Li || !Ui || !Pfin) { free(Li); free(Ui); rc = NGB_E_PANIC; goto done; }
        for (b = 0; b < nblocks; b++) {
            const int k1 = R[b];
            if (R[b + 1] - k1 == 1) continue;
            for (k = k1; k < R[b + 1]; k++) {
                for (p = 0; p < L[k].len; p++) Li[Lp[k] + p] = L[k].i[p] + k1;
                for (p = 0; p < U[k].len; p++) Ui[Up[k] + p] = U[k].i[p] + k1;
            }
        }
        for (k = 0; k < n; k++) Pfin[Pnum[k]] = k;
        for (p = 0; p < Offp[n]; p++) Offi[p] = Pfin[Offi[p]];
        *Li_out = Li; *Ui_out = Ui; *Offi_out = Offi; Offi = NULL;
    }
done:
    if (L) for (k = 0; k < n; k++) { free(L[k].i); free(L[k].x); }
    if (U) for (k = 0; k < n; k++) { free(U[k].i); free(U[k].x); }
    free(L); free(U); free(Rs); free(PSinv); free(Offi); free(X); free(Pinv); free(Pblk); free(Stack); free(Flag);
    free(Ap_pos); free(Lpend); free(Pfin);
    return rc;
}

/* ---------------------------------------------------------------------------------------------- C ABI */
#include "../../include/ngb200.h"

int ngbCircuitSetSymbolic(ngb_circuit *c, int n, int nblocks, const int *P, const int *Q, const int *R)
{
    if (!c->finalized) { ngb_set_error("circuit not finalized"); return NGB_E_PANIC; }
    if (n != c->n || nblocks < 1 || R[0] != 0 || R[nblocks] != n) { ngb_set_error("symbolic analysis does not fit the matrix (n %d, order %d)", n, c->n); return NGB_E_PANIC; }
    free(c->klu_P); free(c->klu_Q); free(c->klu_R);
    c->klu_P = (int *)malloc(sizeof(int) * (size_t)n); c->klu_Q = (int *)malloc(sizeof(int) * (size_t)n);
    c->klu_R = (int *)malloc(sizeof(int) * ((size_t)nblocks + 1));
    if (!c->klu_P || !c->klu_Q || !c->klu_R) return NGB_E_PANIC;
    memcpy(c->klu_P, P, sizeof(int) * (size_t)n); memcpy(c->klu_Q, Q, sizeof(int) * (size_t)n);
    memcpy(c->klu_R, R, sizeof(int) * ((size_t)nblocks + 1));
    c->klu_nblocks = nblocks;
    return NGB_OK;
}

/* Own symbolic analysis: ONE block (no block triangular form) and a greedy minimum-degree ordering of the pattern of
 * A + A' applied to rows and columns alike, ties to the lower index.  It is not klu_analyze's BTF + AMD result, so a
 * run on it agrees with the reference to rounding, not bit for bit; bit-identical runs import klu_analyze's P, Q, R
 * through ngbCircuitSetSymbolic.  The next node comes from a binary heap of (degree, index)
 * pairs with lazy deletion (an entry whose degree is out of date is skipped when it surfaces). */
typedef struct { int deg, idx; } NpHeapItem;
static int np_less(NpHeapItem a, NpHeapItem b) { return a.deg < b.deg || (a.deg == b.deg && a.idx < b.idx); }
static int np_heap_push(NpHeapItem **h, int *len, int *cap, int deg, int idx)
{
    int i;
    if (*len == *cap) {
        NpHeapItem *t = (NpHeapItem *)realloc(*h, sizeof(NpHeapItem) * (size_t)(*cap ? 2 * *cap : 1024));
        if (!t) return 1;
        *h = t; *cap = *cap ? 2 * *cap : 1024;
    }
    i = (*len)++;
    (*h)[i].deg = deg; (*h)[i].idx = idx;
    while (i > 0 && np_less((*h)[i], (*h)[(i - 1) / 2])) { NpHeapItem t = (*h)[i]; (*h)[i] = (*h)[(i - 1) / 2]; (*h)[(i - 1) / 2] = t; i = (i - 1) / 2; }
    return 0;
}
static NpHeapItem np_heap_pop(NpHeapItem *h, int *len)
{
    NpHeapItem top = h[0];
    int i = 0;
    h[0] = h[--(*len)];
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < *len && np_less(h[l], h[m])) m = l;
        if (r < *len && np_less(h[r], h[m])) m = r;
        if (m == i) break;
        { NpHeapItem t = h[i]; h[i] = h[m]; h[m] = t; }
        i = m;
    }
    return top;
}
int ngbCircuitAnalyze(ngb_circuit *c)
{
    const int n = c->n;
    int i, j, k, p, rc = NGB_OK;
    NpHeapItem *heap = NULL; int hlen = 0, hcap = 0;
    if (!c->finalized) { ngb_set_error("circuit not finalized"); return NGB_E_PANIC; }
    /* adjacency as sorted-free integer lists with a mark array */
    int **adj = (int **)calloc((size_t)n, sizeof(int *)), *deg = (int *)calloc((size_t)n, sizeof(int)), *cap = (int *)calloc((size_t)n, sizeof(int));
    int *mark = (int *)malloc(sizeof(int) * (size_t)n), *gone = (int *)calloc((size_t)n, sizeof(int));
    int *ord = (int *)malloc(sizeof(int) * (size_t)n), R[2] = { 0, n };
    if (!adj || !deg || !cap || !mark || !gone || !ord) { rc = NGB_E_PANIC; goto done; }
#define NP_ADD(a, b) do { if (deg[a] == cap[a]) { cap[a] = cap[a] ? 2 * cap[a] : 8; int *t_ = (int *)realloc(adj[a], sizeof(int) * (size_t)cap[a]); \
                          if (!t_) { rc = NGB_E_PANIC; goto done; } adj[a] = t_; } adj[a][deg[a]++] = (b); } while (0)
    for (i = 0; i < n; i++) mark[i] = -1;
    for (j = 0; j < n; j++)                       /* pattern of A + A' without the diagonal, duplicates removed below */
        for (p = c->Ap[j]; p < c->Ap[j + 1]; p++) {
            i = c->Ai[p];
            if (i != j) { NP_ADD(i, j); NP_ADD(j, i); }
        }
    for (i = 0; i < n; i++) {                     /* unique */
        int m = 0;
        for (p = 0; p < deg[i]; p++) if (mark[adj[i][p]] != i) { mark[adj[i][p]] = i; adj[i][m++] = adj[i][p]; }
        deg[i] = m;
    }
    for (i = 0; i < n; i++) mark[i] = -1;
    for (i = 0; i < n; i++) if (np_heap_push(&heap, &hlen, &hcap, deg[i], i)) { rc = NGB_E_PANIC; goto done; }
    for (k = 0; k < n; k++) {
        int best = -1;
        while (hlen > 0) {           /* lowest degree, ties to the lower index */
            const NpHeapItem it = np_heap_pop(heap, &hlen);
            if (!gone[it.idx] && deg[it.idx] == it.deg) { best = it.idx; break; }
        }
        if (best < 0) { rc = NGB_E_PANIC; goto done; }
        ord[k] = best; gone[best] = 1;
        /* eliminate: the remaining neighbours of `best` become a clique */
        for (p = 0; p < deg[best]; p++) {
            const int a = adj[best][p];
            if (gone[a]) continue;
            int m = 0, q;
            for (q = 0; q < deg[a]; q++) if (!gone[adj[a][q]]) { adj[a][m++] = adj[a][q]; mark[adj[a][q]] = n + a + k * 0; }
            deg[a] = m;
            for (q = 0; q < deg[a]; q++) mark[adj[a][q]] = -2 - a;
            for (q = 0; q < deg[best]; q++) {
                const int b2 = adj[best][q];
                if (b2 == a || gone[b2] || mark[b2] == -2 - a) continue;
                mark[b2] = -2 - a;
                NP_ADD(a, b2);
            }
            if (np_heap_push(&heap, &hlen, &hcap, deg[a], a)) { rc = NGB_E_PANIC; goto done; }
        }
    }
#undef NP_ADD
    rc = ngbCircuitSetSymbolic(c, n, 1, ord, ord, R);
done:
    if (adj) for (i = 0; i < n; i++) free(adj[i]);
    free(adj); free(deg); free(cap); free(mark); free(gone); free(ord); free(heap);
    return rc;
}

/* SMPreorder for one sample: pivoting factor of the matrix values Ax (CSC slot order of ngbCircuitGetPattern, what
 * CKTload + LoadGmin left) on the symbolic analysis; the result becomes the pattern set selected by
 * ngbCircuitSelectLuSet.  pivtol <= 0 selects KLU's default threshold 0.001 */
int ngbCircuitFactor(ngb_circuit *c, const double *Ax, double pivtol)
{
    const int n = c->n;
    int rc, sing = -1;
    if (!c->klu_P) { ngb_set_error("no symbolic analysis: call ngbCircuitSetSymbolic or ngbCircuitAnalyze first"); return NGB_E_PANIC; }
    int *Pnum = (int *)malloc(sizeof(int) * (size_t)n), *Lp = (int *)malloc(sizeof(int) * ((size_t)n + 1));
    int *Up = (int *)malloc(sizeof(int) * ((size_t)n + 1)), *Offp = (int *)malloc(sizeof(int) * ((size_t)n + 1));
    int *Li = NULL, *Ui = NULL, *Offi = NULL;
    int *P = (int *)malloc(sizeof(int) * (size_t)n), *Q = (int *)malloc(sizeof(int) * (size_t)n), *R = (int *)malloc(sizeof(int) * ((size_t)c->klu_nblocks + 1));
    const int nb = c->klu_nblocks;
    if (!Pnum || !Lp || !Up || !Offp || !P || !Q || !R) { rc = NGB_E_PANIC; goto done; }
    memcpy(P, c->klu_P, sizeof(int) * (size_t)n); memcpy(Q, c->klu_Q, sizeof(int) * (size_t)n); memcpy(R, c->klu_R, sizeof(int) * ((size_t)nb + 1));
    c->pivtol = pivtol > 0 ? pivtol : 0.001;
    rc = ngb_pivot_factor(n, c->Ap, c->Ai, Ax, nb, P, Q, R, c->pivtol, Pnum, Lp, &Li, Up, &Ui, Offp, &Offi, &sing);
    if (rc == NGB_E_SINGULAR) { ngb_set_error("matrix is singular: zero pivot in column %d", sing); goto done; }
    if (rc) { ngb_set_error("pivoting factor: out of memory"); goto done; }
    rc = ngbCircuitSetLuPattern(c, n, nb, Q, R, Pnum, Lp, Li, Up, Ui, Offp, Offi);
done:
    free(Pnum); free(Lp); free(Up); free(Offp); free(Li); free(Ui); free(Offi); free(P); free(Q); free(R);
    return rc;
}

/* the factor last given to (or computed for) ngbCircuitSetLuPattern: Pnum [n], Lp / Up / Offp [n+1], Li [Lp[n]], Ui [Up[n]],
 * Offi [Offp[n]]; any pointer may be NULL */
int ngbCircuitGetLuPattern(const ngb_circuit *c, int *Pnum, int *Lp, int *Li, int *Up, int *Ui, int *Offp, int *Offi)
{
    const int n = c->n;
    if (!c->pat_Lp) { ngb_set_error("no LU pattern on the circuit"); return NGB_E_PANIC; }
    if (Pnum) memcpy(Pnum, c->klu_Pnum, sizeof(int) * (size_t)n);
    if (Lp) memcpy(Lp, c->pat_Lp, sizeof(int) * ((size_t)n + 1));
    if (Up) memcpy(Up, c->pat_Up, sizeof(int) * ((size_t)n + 1));
    if (Offp) memcpy(Offp, c->pat_Offp, sizeof(int) * ((size_t)n + 1));
    if (Li) memcpy(Li, c->pat_Li, sizeof(int) * (size_t)c->pat_Lp[n]);
    if (Ui) memcpy(Ui, c->pat_Ui, sizeof(int) * (size_t)c->pat_Up[n]);
    if (Offi) memcpy(Offi, c->pat_Offi, sizeof(int) * (size_t)c->pat_Offp[n]);
    return NGB_OK;
}
