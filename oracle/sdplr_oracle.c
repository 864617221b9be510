/*
 * sdplr_oracle.c -- CPU restatement of the SDPLRPlus.jl per-iteration hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (sdplrplus.jl_b200/, the
 * C-ABI library) links, imports or executes this file.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may use it, and only as the checker / the timed CPU baseline.
 *
 * Parity status: the reference is pure Julia and no Julia toolchain exists in
 * this image, so the oracle cannot be compared with outputs of the reference
 * itself.  It is pinned instead against (a) the hand-derived golden index maps
 * of SURVEY.md Appendix C (tests/golden/), (b) the dense-reference identities
 * and finite-difference checks of the reference's own test/coreop.jl and
 * (c) the K2 known answers of test/maxcut.jl and test/minimumbisection.jl.
 * Random streams (Julia Xoshiro) are "parity unpinned": R0 / Lanczos start
 * vectors are always injected.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference checkout).  Index arrays exported by orc_export_* are 1-based
 * int64 exactly as Julia would hold them; internally everything is 0-based.
 *
 * Layout: Rt is r x n column-major (Julia), i.e. vertex j's r numbers are
 * contiguous at Rt[j*r .. j*r+r).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;

typedef struct {
    i64 gid;   /* 0-based global slot in the (m+1)-vectors */
    i64 s;     /* inner dimension */
    double *B; /* n x s column-major */
    double *D; /* s */
} orc_lowrank;

typedef struct {
    i64 n, m, nA;
    /* aggregated pattern, 0-based (src/preprocess.jl:24-169) */
    i64 nnzT, nnzF, Ec;
    i64 *triu_colptr, *triu_rowval; /* n+1, nnzT */
    i64 *matptr;                    /* nA+1 */
    i64 *nzind;                     /* Ec: slot in triu CSC */
    double *nzval_one, *nzval_two;  /* Ec */
    i64 *full_colptr, *full_rowval; /* n+1, nnzF */
    i64 *mapped;                    /* nnzF: full slot -> triu slot */
    i64 *sparse_gid;                /* nA: 0-based global slot (m for C) */
    double *triuS, *S, *UVt;        /* nnzT, nnzF, nnzT (src/structs.jl:346-360) */
    /* low-rank matrices (src/structs.jl:11-24) */
    i64 nlr;
    orc_lowrank *lr;
    /* problem data (src/structs.jl:150-162) */
    double *b;
    /* SolverVars (src/structs.jl:194-223) */
    i64 r;
    double sigma, obj;
    double *Rt, *Gt;
    double *lambda, *lambda_ub, *y, *pvio_raw, *pvio_lb, *pvio, *A_RD, *A_DD;
    /* L-BFGS history (src/lbfgs.jl:4-28) */
    i64 h, latest; /* latest is 1-based like the reference */
    double **hs, **hy, *hrho, *ha;
} orc_ctx;

static void *xcalloc(size_t n, size_t sz) {
    void *p = calloc(n ? n : 1, sz);
    if (!p) { fprintf(stderr, "oracle: out of memory\n"); abort(); }
    return p;
}

void orc_set_threads(int t) {
#ifdef _OPENMP
    if (t > 0) omp_set_num_threads(t);
#else
    (void)t;
#endif
}
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

orc_ctx *orc_create(i64 n, i64 m) {
    orc_ctx *c = (orc_ctx *)xcalloc(1, sizeof(orc_ctx));
    c->n = n; c->m = m;
    c->b = (double *)xcalloc(m, sizeof(double));
    c->lambda = (double *)xcalloc(m, sizeof(double));
    c->lambda_ub = (double *)xcalloc(m, sizeof(double));
    c->pvio_lb = (double *)xcalloc(m, sizeof(double));
    c->pvio = (double *)xcalloc(m, sizeof(double));
    c->y = (double *)xcalloc(m + 1, sizeof(double));
    c->pvio_raw = (double *)xcalloc(m + 1, sizeof(double));
    c->A_RD = (double *)xcalloc(m + 1, sizeof(double));
    c->A_DD = (double *)xcalloc(m + 1, sizeof(double));
    for (i64 i = 0; i < m; i++) { c->lambda_ub[i] = INFINITY; c->pvio_lb[i] = -INFINITY; }
    c->sigma = 2.0;
    return c;
}

static void free_lbfgs(orc_ctx *c) {
    for (i64 j = 0; j < c->h; j++) { free(c->hs[j]); free(c->hy[j]); }
    free(c->hs); free(c->hy); free(c->hrho); free(c->ha);
    c->hs = c->hy = NULL; c->hrho = c->ha = NULL; c->h = 0;
}

void orc_destroy(orc_ctx *c) {
    if (!c) return;
    free(c->triu_colptr); free(c->triu_rowval); free(c->matptr); free(c->nzind);
    free(c->nzval_one); free(c->nzval_two); free(c->full_colptr); free(c->full_rowval);
    free(c->mapped); free(c->sparse_gid); free(c->triuS); free(c->S); free(c->UVt);
    for (i64 i = 0; i < c->nlr; i++) { free(c->lr[i].B); free(c->lr[i].D); }
    free(c->lr);
    free(c->b); free(c->lambda); free(c->lambda_ub); free(c->pvio_lb); free(c->pvio);
    free(c->y); free(c->pvio_raw); free(c->A_RD); free(c->A_DD);
    free(c->Rt); free(c->Gt);
    free_lbfgs(c);
    free(c);
}

/* ------------------------------------------------------------------------- */
/* sparse(I,J,1,n,n): SparseArrays stdlib contract used at
 * src/preprocess.jl:90,93 -- CSC, rows ascending within a column, duplicate
 * coordinates merged (the pattern is the union; values are irrelevant).
 * Two stable counting sorts (by row, then by column) + unique. 0-based in/out. */
static void build_csc_pattern(i64 n, i64 nnz, const i64 *I, const i64 *J,
                              i64 **colptr_out, i64 **rowval_out, i64 *nnz_out) {
    i64 *cnt = (i64 *)xcalloc(n + 1, sizeof(i64));
    i64 *perm = (i64 *)xcalloc(nnz, sizeof(i64));
    for (i64 k = 0; k < nnz; k++) cnt[I[k] + 1]++;
    for (i64 i = 0; i < n; i++) cnt[i + 1] += cnt[i];
    for (i64 k = 0; k < nnz; k++) perm[cnt[I[k]]++] = k; /* sorted by row */
    i64 *cnt2 = (i64 *)xcalloc(n + 1, sizeof(i64));
    for (i64 k = 0; k < nnz; k++) cnt2[J[k] + 1]++;
    for (i64 i = 0; i < n; i++) cnt2[i + 1] += cnt2[i];
    i64 *rows = (i64 *)xcalloc(nnz, sizeof(i64));
    i64 *cols = (i64 *)xcalloc(nnz, sizeof(i64));
    for (i64 q = 0; q < nnz; q++) { /* stable by column */
        i64 k = perm[q];
        i64 p = cnt2[J[k]]++;
        rows[p] = I[k]; cols[p] = J[k];
    }
    i64 *colptr = (i64 *)xcalloc(n + 1, sizeof(i64));
    i64 u = 0;
    for (i64 p = 0; p < nnz; p++) {
        if (p > 0 && rows[p] == rows[p - 1] && cols[p] == cols[p - 1]) continue;
        rows[u] = rows[p]; cols[u] = cols[p];
        colptr[cols[u] + 1]++;
        u++;
    }
    for (i64 i = 0; i < n; i++) colptr[i + 1] += colptr[i];
    i64 *rowval = (i64 *)xcalloc(u, sizeof(i64));
    memcpy(rowval, rows, (size_t)u * sizeof(i64));
    free(cnt); free(cnt2); free(perm); free(rows); free(cols);
    *colptr_out = colptr; *rowval_out = rowval; *nnz_out = u;
}

/* binary search of `row` inside column `col` of a CSC pattern; -1 if absent
 * (src/preprocess.jl:109-123 and :146-157). */
static i64 csc_find(const i64 *colptr, const i64 *rowval, i64 col, i64 row) {
    i64 low = colptr[col], high = colptr[col + 1] - 1;
    while (low <= high) {
        i64 mid = (low + high) / 2;
        if (rowval[mid] == row) return mid;
        else if (rowval[mid] < row) low = mid + 1;
        else high = mid - 1;
    }
    return -1;
}

static void alloc_S(orc_ctx *c) {
    free(c->triuS); free(c->S); free(c->UVt);
    c->triuS = (double *)xcalloc(c->nnzT, sizeof(double));
    c->S = (double *)xcalloc(c->nnzF, sizeof(double));
    c->UVt = (double *)xcalloc(c->nnzT, sizeof(double));
}

/* preprocess_sparsecons (src/preprocess.jl:24-169) fed with what
 * SolverAuxiliary (src/structs.jl:296-361) collects: the nA sparse matrices
 * (sparse/diagonal A_i in order of appearance, then C if sparse) as
 * concatenated 1-based triplets in `findnz` order, with their global ids
 * (1-based; m+1 for C).  triu() of each matrix keeps entries with i<=j in
 * stored order (src/preprocess.jl:4-16). */
int orc_preprocess(orc_ctx *c, i64 nA, const i64 *mat_off, const i64 *I1, const i64 *J1,
                   const double *V, const i64 *sparse_gid1) {
    i64 n = c->n;
    i64 total_nnz = mat_off[nA];
    c->nA = nA;
    /* count (src/preprocess.jl:47-52) */
    i64 total_triu = 0;
    for (i64 k = 0; k < total_nnz; k++)
        if (I1[k] <= J1[k]) total_triu++;
    /* concatenate coordinates with value one (src/preprocess.jl:56-82) */
    i64 *aI = (i64 *)xcalloc(total_nnz, sizeof(i64)), *aJ = (i64 *)xcalloc(total_nnz, sizeof(i64));
    i64 *tI = (i64 *)xcalloc(total_triu, sizeof(i64)), *tJ = (i64 *)xcalloc(total_triu, sizeof(i64));
    i64 t = 0;
    for (i64 k = 0; k < total_nnz; k++) {
        if (I1[k] < 1 || I1[k] > n || J1[k] < 1 || J1[k] > n) { free(aI); free(aJ); free(tI); free(tJ); return -2; }
        aI[k] = I1[k] - 1; aJ[k] = J1[k] - 1;
        if (I1[k] <= J1[k]) { tI[t] = I1[k] - 1; tJ[t] = J1[k] - 1; t++; }
    }
    free(c->triu_colptr); free(c->triu_rowval); free(c->full_colptr); free(c->full_rowval);
    /* the two sparse() calls (src/preprocess.jl:90,93) */
    build_csc_pattern(n, total_triu, tI, tJ, &c->triu_colptr, &c->triu_rowval, &c->nnzT);
    build_csc_pattern(n, total_nnz, aI, aJ, &c->full_colptr, &c->full_rowval, &c->nnzF);
    free(aI); free(aJ); free(tI); free(tJ);

    /* per-matrix entry list (src/preprocess.jl:95-135) */
    free(c->matptr); free(c->nzind); free(c->nzval_one); free(c->nzval_two); free(c->sparse_gid);
    c->Ec = total_triu;
    c->matptr = (i64 *)xcalloc(nA + 1, sizeof(i64));
    c->nzind = (i64 *)xcalloc(total_triu, sizeof(i64));
    c->nzval_one = (double *)xcalloc(total_triu, sizeof(double));
    c->nzval_two = (double *)xcalloc(total_triu, sizeof(double));
    c->sparse_gid = (i64 *)xcalloc(nA, sizeof(i64));
    /* matptr first (counts of kept entries per matrix), then the entries of the matrices in parallel: each matrix fills
     * its own span in stored order, so the result is the sequential loop's */
    for (i64 i = 0; i < nA; i++) {
        i64 cnt = 0;
        for (i64 k = mat_off[i]; k < mat_off[i + 1]; k++) cnt += (I1[k] <= J1[k]);
        c->matptr[i + 1] = c->matptr[i] + cnt;
        c->sparse_gid[i] = sparse_gid1[i] - 1;
    }
#pragma omp parallel for schedule(dynamic, 4096)
    for (i64 i = 0; i < nA; i++) {
        i64 cumul = c->matptr[i];
        for (i64 k = mat_off[i]; k < mat_off[i + 1]; k++) {
            i64 row = I1[k] - 1, col = J1[k] - 1;
            if (row > col) continue;
            i64 slot = csc_find(c->triu_colptr, c->triu_rowval, col, row);
            c->nzind[cumul] = slot;
            c->nzval_one[cumul] = V[k];
            c->nzval_two[cumul] = (row == col) ? V[k] : 2.0 * V[k]; /* :125-132 */
            cumul++;
        }
    }
    c->matptr[nA] = total_triu;
    /* full -> triu map (src/preprocess.jl:137-159); an unmatched lower entry
     * stays 0 in Julia, i.e. -1 here. */
    free(c->mapped);
    c->mapped = (i64 *)xcalloc(c->nnzF, sizeof(i64));
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(|:bad)
    for (i64 col = 0; col < n; col++) {
        for (i64 nzi = c->full_colptr[col]; nzi < c->full_colptr[col + 1]; nzi++) {
            i64 row = c->full_rowval[nzi];
            i64 rr = row < col ? row : col, cc = row < col ? col : row;
            i64 slot = csc_find(c->triu_colptr, c->triu_rowval, cc, rr);
            c->mapped[nzi] = slot;
            if (slot < 0) bad = 1;
        }
    }
    alloc_S(c);
    return bad ? 1 : 0; /* 1 = asymmetric storage detected (maps still exported) */
}

/* adopt maps computed elsewhere (1-based, Julia layout); used by the bench
 * CPU-baseline leg so that the 10M-vertex sample does not pay a CPU sort. */
int orc_set_maps(orc_ctx *c, i64 nA, i64 nnzT, i64 nnzF, i64 Ec, const i64 *triu_colptr,
                 const i64 *triu_rowval, const i64 *matptr, const i64 *nzind, const double *one,
                 const double *two, const i64 *full_colptr, const i64 *full_rowval,
                 const i64 *mapped, const i64 *sparse_gid1) {
    i64 n = c->n;
    c->nA = nA; c->nnzT = nnzT; c->nnzF = nnzF; c->Ec = Ec;
    free(c->triu_colptr); free(c->triu_rowval); free(c->full_colptr); free(c->full_rowval);
    free(c->matptr); free(c->nzind); free(c->nzval_one); free(c->nzval_two); free(c->sparse_gid);
    free(c->mapped);
    c->triu_colptr = (i64 *)xcalloc(n + 1, sizeof(i64));
    c->full_colptr = (i64 *)xcalloc(n + 1, sizeof(i64));
    c->triu_rowval = (i64 *)xcalloc(nnzT, sizeof(i64));
    c->full_rowval = (i64 *)xcalloc(nnzF, sizeof(i64));
    c->mapped = (i64 *)xcalloc(nnzF, sizeof(i64));
    c->matptr = (i64 *)xcalloc(nA + 1, sizeof(i64));
    c->nzind = (i64 *)xcalloc(Ec, sizeof(i64));
    c->nzval_one = (double *)xcalloc(Ec, sizeof(double));
    c->nzval_two = (double *)xcalloc(Ec, sizeof(double));
    c->sparse_gid = (i64 *)xcalloc(nA, sizeof(i64));
    for (i64 i = 0; i <= n; i++) { c->triu_colptr[i] = triu_colptr[i] - 1; c->full_colptr[i] = full_colptr[i] - 1; }
    for (i64 k = 0; k < nnzT; k++) c->triu_rowval[k] = triu_rowval[k] - 1;
    for (i64 k = 0; k < nnzF; k++) { c->full_rowval[k] = full_rowval[k] - 1; c->mapped[k] = mapped[k] - 1; }
    for (i64 i = 0; i <= nA; i++) c->matptr[i] = matptr[i] - 1;
    for (i64 k = 0; k < Ec; k++) { c->nzind[k] = nzind[k] - 1; c->nzval_one[k] = one[k]; c->nzval_two[k] = two[k]; }
    for (i64 i = 0; i < nA; i++) c->sparse_gid[i] = sparse_gid1[i] - 1;
    alloc_S(c);
    return 0;
}

void orc_pattern_sizes(const orc_ctx *c, i64 *nnzT, i64 *nnzF, i64 *Ec) {
    *nnzT = c->nnzT; *nnzF = c->nnzF; *Ec = c->Ec;
}

/* 1-based export, Julia layout */
void orc_pattern_export(const orc_ctx *c, i64 *triu_colptr, i64 *triu_rowval, i64 *matptr, i64 *nzind,
                        double *one, double *two, i64 *full_colptr, i64 *full_rowval, i64 *mapped) {
    for (i64 i = 0; i <= c->n; i++) { triu_colptr[i] = c->triu_colptr[i] + 1; full_colptr[i] = c->full_colptr[i] + 1; }
    for (i64 k = 0; k < c->nnzT; k++) triu_rowval[k] = c->triu_rowval[k] + 1;
    for (i64 i = 0; i <= c->nA; i++) matptr[i] = c->matptr[i] + 1;
    for (i64 k = 0; k < c->Ec; k++) { nzind[k] = c->nzind[k] + 1; one[k] = c->nzval_one[k]; two[k] = c->nzval_two[k]; }
    for (i64 k = 0; k < c->nnzF; k++) { full_rowval[k] = c->full_rowval[k] + 1; mapped[k] = c->mapped[k] + 1; }
}

/* SymLowRankMatrix(D, B) (src/structs.jl:11-24); gid1 is the 1-based global
 * id (m+1 for the objective) recorded at src/structs.jl:310-312,326-328. */
int orc_add_symlowrank(orc_ctx *c, i64 gid1, i64 s, const double *B, const double *D) {
    c->lr = (orc_lowrank *)realloc(c->lr, (size_t)(c->nlr + 1) * sizeof(orc_lowrank));
    orc_lowrank *L = &c->lr[c->nlr++];
    L->gid = gid1 - 1; L->s = s;
    L->B = (double *)xcalloc(c->n * s, sizeof(double));
    L->D = (double *)xcalloc(s, sizeof(double));
    memcpy(L->B, B, (size_t)(c->n * s) * sizeof(double));
    memcpy(L->D, D, (size_t)s * sizeof(double));
    return 0;
}

/* SDPData.b and constraint types -> lambda_ub / primal_vio_lb
 * (src/structs.jl:228,248-249): inequality => ub 0, lb 0; equality => Inf, -Inf */
void orc_set_problem(orc_ctx *c, const double *b, const uint8_t *is_ineq) {
    for (i64 i = 0; i < c->m; i++) {
        c->b[i] = b[i];
        int q = is_ineq ? is_ineq[i] : 0;
        c->lambda_ub[i] = q ? 0.0 : INFINITY;
        c->pvio_lb[i] = q ? 0.0 : -INFINITY;
    }
}

/* SolverVars(Rt0, lambda0, lambda_ub, r, sigma0) (src/structs.jl:242-263);
 * lambda0 is clipped to lambda_ub as at src/structs.jl:233. */
void orc_init_vars(orc_ctx *c, i64 r, const double *Rt0, const double *lambda0, double sigma0) {
    i64 N = c->n * r;
    free(c->Rt); free(c->Gt);
    c->r = r;
    c->Rt = (double *)xcalloc(N, sizeof(double));
    c->Gt = (double *)xcalloc(N, sizeof(double));
    memcpy(c->Rt, Rt0, (size_t)N * sizeof(double));
    for (i64 i = 0; i < c->m; i++) {
        double l = lambda0 ? lambda0[i] : 0.0;
        c->lambda[i] = l < c->lambda_ub[i] ? l : c->lambda_ub[i];
        c->pvio[i] = 0.0;
    }
    for (i64 i = 0; i <= c->m; i++) c->y[i] = c->pvio_raw[i] = c->A_RD[i] = c->A_DD[i] = 0.0;
    c->sigma = sigma0; c->obj = 0.0;
}

/* raw accessors for the ctypes harness */
double *orc_ptr(orc_ctx *c, int which) {
    switch (which) {
    case 0: return c->Rt;
    case 1: return c->Gt;
    case 2: return c->lambda;
    case 3: return c->lambda_ub;
    case 4: return c->y;
    case 5: return c->pvio_raw;
    case 6: return c->pvio_lb;
    case 7: return c->pvio;
    case 8: return c->A_RD;
    case 9: return c->A_DD;
    case 10: return c->b;
    case 11: return c->triuS;
    case 12: return c->S;
    default: return NULL;
    }
}
double orc_get_sigma(const orc_ctx *c) { return c->sigma; }
void orc_set_sigma(orc_ctx *c, double s) { c->sigma = s; }
double orc_get_obj(const orc_ctx *c) { return c->obj; }
void orc_set_obj(orc_ctx *c, double o) { c->obj = o; }
i64 orc_get_rank(const orc_ctx *c) { return c->r; }

/* ------------------------------------------------------------------------- */
/* mydot (src/coreop.jl:153-160) */
static inline double mydot1(const double *Ut, i64 r, i64 a, i64 b) {
    const double *x = Ut + a * r, *y = Ut + b * r;
    double s = 0.0;
    for (i64 i = 0; i < r; i++) s += x[i] * y[i];
    return s;
}
/* mydot two-argument (src/coreop.jl:162-172) */
static inline double mydot2(const double *Ut, const double *Vt, i64 r, i64 a, i64 b) {
    double s = 0.0;
    const double *ua = Ut + a * r, *vb = Vt + b * r, *va = Vt + a * r, *ub = Ut + b * r;
    for (i64 i = 0; i < r; i++) s += ua[i] * vb[i];
    for (i64 i = 0; i < r; i++) s += va[i] * ub[i];
    return s / 2.0;
}

/* A_sparse_formUUt! / formUVt! (src/coreop.jl:174-203): one sampled dot per
 * slot of the upper-triangular aggregated pattern. */
static void form_UVt(orc_ctx *c, i64 r, const double *Ut, const double *Vt) {
    const i64 *colptr = c->triu_colptr, *rowval = c->triu_rowval;
    double *out = c->UVt;
#pragma omp parallel for schedule(dynamic, 1024)
    for (i64 col = 0; col < c->n; col++)
        for (i64 nzi = colptr[col]; nzi < colptr[col + 1]; nzi++) {
            i64 row = rowval[nzi];
            out[nzi] = Vt ? mydot2(Ut, Vt, r, col, row) : mydot1(Ut, r, col, row);
        }
}

/* A_sparse! (src/coreop.jl:72-113): v = UVt' * CSC(nnzT x nA; matptr, nzind,
 * nzval_two) -- a segmented reduction -- then scatter to the global slots. */
static void A_sparse(orc_ctx *c, double *out, i64 r, const double *Ut, const double *Vt) {
    form_UVt(c, r, Ut, Vt);
#pragma omp parallel for schedule(dynamic, 4096)
    for (i64 i = 0; i < c->nA; i++) {
        double s = 0.0;
        for (i64 k = c->matptr[i]; k < c->matptr[i + 1]; k++) s += c->UVt[c->nzind[k]] * c->nzval_two[k];
        out[c->sparse_gid[i]] = s;
    }
}

/* Ut * B  (r x n times n x s -> r x s), used by tr_UtAU / tr_UtAV and the
 * low-rank mul! (src/coreop.jl:115-130, src/structs.jl:135-145). */
static void UtB(const orc_ctx *c, const orc_lowrank *L, i64 r, const double *Ut, double *out /* r*s */) {
    for (i64 q = 0; q < r * L->s; q++) out[q] = 0.0;
    for (i64 k = 0; k < L->s; k++)
        for (i64 j = 0; j < c->n; j++) {
            double bjk = L->B[j + k * c->n];
            const double *u = Ut + j * r;
            for (i64 i = 0; i < r; i++) out[i + k * r] += u[i] * bjk;
        }
}

/* A_symlowrank! (src/coreop.jl:132-151) with tr_UtAU (:115-120) / tr_UtAV (:122-130) */
static void A_symlowrank(orc_ctx *c, double *out, i64 r, const double *Ut, const double *Vt) {
    for (i64 q = 0; q < c->nlr; q++) {
        orc_lowrank *L = &c->lr[q];
        double *ub = (double *)xcalloc(r * L->s, sizeof(double));
        double *vb = Vt ? (double *)xcalloc(r * L->s, sizeof(double)) : ub;
        UtB(c, L, r, Ut, ub);
        if (Vt) UtB(c, L, r, Vt, vb);
        double tot = 0.0;
        for (i64 k = 0; k < L->s; k++)
            for (i64 i = 0; i < r; i++) tot += ub[i + k * r] * vb[i + k * r] * L->D[k];
        out[L->gid] = tot;
        if (Vt) free(vb);
        free(ub);
    }
}

/* A!(out, aux, Ut)  (src/coreop.jl:36-49): out has m+1 slots */
void orc_A_uu(orc_ctx *c, double *out, i64 r, const double *Ut) {
    for (i64 i = 0; i <= c->m; i++) out[i] = 0.0;
    if (c->nA > 0) A_sparse(c, out, r, Ut, NULL);
    if (c->nlr > 0) A_symlowrank(c, out, r, Ut, NULL);
}
/* A!(out, aux, Ut, Vt) = A((UV' + VU')/2)  (src/coreop.jl:54-70) */
void orc_A_uv(orc_ctx *c, double *out, i64 r, const double *Ut, const double *Vt) {
    for (i64 i = 0; i <= c->m; i++) out[i] = 0.0;
    if (c->nA > 0) A_sparse(c, out, r, Ut, Vt);
    if (c->nlr > 0) A_symlowrank(c, out, r, Ut, Vt);
}

/* f! (src/coreop.jl:11-31) */
double orc_f(orc_ctx *c) {
    i64 m = c->m;
    orc_A_uu(c, c->pvio_raw, c->r, c->Rt);
    c->obj = c->pvio_raw[m];
    double sigma = c->sigma, L = c->obj;
    for (i64 i = 0; i < m; i++) {
        c->pvio_raw[i] -= c->b[i];
        double v = c->pvio_raw[i];
        c->pvio[i] = v > c->pvio_lb[i] ? v : c->pvio_lb[i];
    }
    for (i64 i = 0; i < m; i++) {
        double yi = c->lambda[i] - sigma * c->pvio_raw[i];
        if (c->lambda_ub[i] < yi) yi = c->lambda_ub[i];
        L += (yi * yi - c->lambda[i] * c->lambda[i]) / (2.0 * sigma);
    }
    return L;
}

/* copy2y_lambda_sub_pvio! (src/coreop.jl:229-236) */
void orc_copy2y(orc_ctx *c) {
    for (i64 i = 0; i < c->m; i++) {
        double t = c->lambda[i] - c->sigma * c->pvio_raw[i];
        if (c->lambda_ub[i] < t) t = c->lambda_ub[i];
        c->y[i] = -t;
    }
    c->y[c->m] = 1.0;
}

/* At_preprocess! (src/coreop.jl:248-258) -> At_preprocess_sparse! (:205-227):
 * triuS.nzval = CSC(nnzT x nA; matptr,nzind,nzval_one) * y[sparse inds], then
 * S.nzval[k] = triuS.nzval[mapped[k]]. */
void orc_At_preprocess(orc_ctx *c) {
    if (c->nA <= 0) return;
    for (i64 k = 0; k < c->nnzT; k++) c->triuS[k] = 0.0;
    for (i64 i = 0; i < c->nA; i++) {
        double v = c->y[c->sparse_gid[i]];
        for (i64 k = c->matptr[i]; k < c->matptr[i + 1]; k++) c->triuS[c->nzind[k]] += c->nzval_one[k] * v;
    }
#pragma omp parallel for schedule(static)
    for (i64 k = 0; k < c->nnzF; k++) c->S[k] = c->triuS[c->mapped[k]];
}

/* Y (+)= alpha * (X*B) * D * B'   (src/structs.jl:135-145), X is r x n */
static void lowrank_left(const orc_ctx *c, const orc_lowrank *L, i64 r, double *Y, const double *X, double alpha) {
    double *xb = (double *)xcalloc(r * L->s, sizeof(double));
    UtB(c, L, r, X, xb);
    for (i64 k = 0; k < L->s; k++)
        for (i64 i = 0; i < r; i++) xb[i + k * r] *= L->D[k];
#pragma omp parallel for schedule(static)
    for (i64 j = 0; j < c->n; j++)
        for (i64 k = 0; k < L->s; k++) {
            double bjk = alpha * L->B[j + k * c->n];
            for (i64 i = 0; i < r; i++) Y[j * r + i] += xb[i + k * r] * bjk;
        }
    free(xb);
}

/* At!(Y, X, aux, var): Y = X*S + sum_g y[g] X (B D B')   (src/coreop.jl:260-279).
 * X, Y are r x n column-major. S is CSC: Y[:,j] += X[:,rowval[k]]*S[k]. */
void orc_At_left(orc_ctx *c, double *Y, i64 r, const double *X) {
    i64 n = c->n;
    memset(Y, 0, (size_t)(n * r) * sizeof(double));
    if (c->nA > 0) {
#pragma omp parallel for schedule(dynamic, 1024)
        for (i64 j = 0; j < n; j++) {
            double *yj = Y + j * r;
            for (i64 k = c->full_colptr[j]; k < c->full_colptr[j + 1]; k++) {
                const double *x = X + c->full_rowval[k] * r;
                double s = c->S[k];
                for (i64 i = 0; i < r; i++) yj[i] += x[i] * s;
            }
        }
    }
    for (i64 q = 0; q < c->nlr; q++) lowrank_left(c, &c->lr[q], r, Y, X, c->y[c->lr[q].gid]);
}

/* At!(y, aux, x, var): y = S*x + sum_g y[g] (B D B') x   (src/coreop.jl:281-300).
 * x, y are n x nc column-major. */
void orc_At_right(orc_ctx *c, double *y, i64 nc, const double *x) {
    i64 n = c->n;
    memset(y, 0, (size_t)(n * nc) * sizeof(double));
    if (c->nA > 0) {
        for (i64 q = 0; q < nc; q++) {
            const double *xq = x + q * n;
            double *yq = y + q * n;
            /* CSC S*x: column j scatters into rows; S is symmetric so the
             * row-gather form gives the same sums in the same k order. */
#pragma omp parallel for schedule(dynamic, 1024)
            for (i64 j = 0; j < n; j++) {
                double s = 0.0;
                for (i64 k = c->full_colptr[j]; k < c->full_colptr[j + 1]; k++) s += c->S[k] * xq[c->full_rowval[k]];
                yq[j] = s;
            }
        }
    }
    for (i64 g = 0; g < c->nlr; g++) {
        orc_lowrank *L = &c->lr[g];
        double coeff = c->y[L->gid];
        for (i64 q = 0; q < nc; q++) {
            const double *xq = x + q * n;
            double *yq = y + q * n;
            for (i64 k = 0; k < L->s; k++) {
                double t = 0.0;
                for (i64 j = 0; j < n; j++) t += L->B[j + k * n] * xq[j];
                t *= L->D[k] * coeff;
                for (i64 j = 0; j < n; j++) yq[j] += L->B[j + k * n] * t;
            }
        }
    }
}

/* g! (src/coreop.jl:305-317) */
void orc_g(orc_ctx *c) {
    orc_copy2y(c);
    orc_At_preprocess(c);
    orc_At_left(c, c->Gt, c->r, c->Rt);
    i64 N = c->n * c->r;
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < N; i++) c->Gt[i] *= 2.0;
}

static double nrm2(const double *x, i64 N) {
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+:s)
    for (i64 i = 0; i < N; i++) s += x[i] * x[i];
    return sqrt(s);
}
static double ddot(const double *x, const double *y, i64 N) {
    double s = 0.0;
#pragma omp parallel for schedule(static) reduction(+:s)
    for (i64 i = 0; i < N; i++) s += x[i] * y[i];
    return s;
}
static void daxpy(double a, const double *x, double *y, i64 N) {
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < N; i++) y[i] += a * x[i];
}
static void dscal(double a, double *x, i64 N) {
#pragma omp parallel for schedule(static)
    for (i64 i = 0; i < N; i++) x[i] *= a;
}
double orc_dot(const double *x, const double *y, i64 N) { return ddot(x, y, N); }
double orc_nrm2(const double *x, i64 N) { return nrm2(x, N); }
void orc_axpy(double a, const double *x, double *y, i64 N) { daxpy(a, x, y, N); }
void orc_scal(double a, double *x, i64 N) { dscal(a, x, N); }

/* fg! (src/coreop.jl:323-349): out = (L, ||G||_F (/normC), ||pvio||_2 (/normb));
 * pass normC = normb = 1 for the :absolute modes. */
void orc_fg(orc_ctx *c, double normC, double normb, double *out3) {
    double L = orc_f(c);
    orc_g(c);
    for (i64 i = 0; i < c->m; i++) {
        double v = c->pvio_raw[i];
        c->pvio[i] = v > c->pvio_lb[i] ? v : c->pvio_lb[i];
    }
    out3[0] = L;
    out3[1] = nrm2(c->Gt, c->n * c->r) / normC;
    out3[2] = nrm2(c->pvio, c->m) / normb;
}

/* ------------------------------------------------------------------------- */
static double cubic_newton(const double *c, double x, int iters) {
    for (int it = 0; it < iters; it++) {
        double f = ((c[3] * x + c[2]) * x + c[1]) * x + c[0];
        double df = (3.0 * c[3] * x + 2.0 * c[2]) * x + c[1];
        if (df == 0.0 || !isfinite(f / df)) break;
        double xn = x - f / df;
        if (xn == x) break;
        x = xn;
    }
    return x;
}

/* real roots of c0 + c1 x + c2 x^2 + c3 x^3 (c3 != 0).
 * Stands in for PolynomialRoots.roots (src/linesearch.jl:82,94): the
 * reference keeps only roots with |imag| < eps, i.e. the real ones.
 * Closed form for the root of LARGEST magnitude (the one Cardano / the
 * trigonometric form deliver without cancellation), Newton-polished on the
 * original coefficients; backward deflation by it leaves a quadratic whose
 * roots come from the cancellation-free quadratic formula, so the small roots
 * survive a nearly vanishing leading coefficient. */
static int cubic_real_roots(const double *c, double *roots) {
    double a = c[2] / c[3], b = c[1] / c[3], d = c[0] / c[3];
    double p = b - a * a / 3.0, q = 2.0 * a * a * a / 27.0 - a * b / 3.0 + d;
    double disc = q * q / 4.0 + p * p * p / 27.0;
    double r1;
    if (disc > 0) {
        double sq = sqrt(disc);
        double u = cbrt(-q / 2.0 + sq), v = cbrt(-q / 2.0 - sq);
        r1 = u + v - a / 3.0;
    } else if (p == 0.0) {
        r1 = -a / 3.0;
    } else {
        double m = 2.0 * sqrt(-p / 3.0);
        double arg = 3.0 * q / (p * m);
        if (arg > 1) arg = 1;
        if (arg < -1) arg = -1;
        double th = acos(arg) / 3.0;
        r1 = 0.0;
        for (int k = 0; k < 3; k++) {
            double t = m * cos(th - 2.0 * M_PI * k / 3.0) - a / 3.0;
            if (fabs(t) > fabs(r1)) r1 = t;
        }
    }
    if (!isfinite(r1)) return 0;
    r1 = cubic_newton(c, r1, 30);
    int nr = 0;
    roots[nr++] = r1;
    if (r1 == 0.0) return nr;
    /* (x - r1)(c3 x^2 + b1 x + b0), BACKWARD deflation (from the constant term): the stable direction for the
     * root of largest magnitude */
    double b0 = -c[0] / r1, b1 = (b0 - c[1]) / r1;
    double dq = b1 * b1 - 4.0 * c[3] * b0;
    if (dq >= 0) {
        double sq = sqrt(dq);
        double t = -0.5 * (b1 + (b1 >= 0 ? sq : -sq));
        double x1 = t / c[3], x2 = (t != 0.0) ? b0 / t : 0.0;
        roots[nr++] = cubic_newton(c, x1, 8);
        roots[nr++] = cubic_newton(c, x2, 8);
    }
    return nr;
}

/* quartic coefficients of the augmented Lagrangian along Dt
 * (src/linesearch.jl:36-56); A_RD (already x2) and A_DD must be filled. */
void orc_biquadratic(const orc_ctx *c, double *bq) {
    i64 m = c->m;
    double p0 = c->obj, p1 = c->A_RD[m], p2 = c->A_DD[m], sigma = c->sigma;
    double lq0 = 0, q0q0 = 0, lq1 = 0, q0q1 = 0, yq2 = 0, q1q1 = 0, q1q2 = 0, q2q2 = 0;
    for (i64 i = 0; i < m; i++) {
        double l = c->lambda[i], q0 = c->pvio_raw[i], q1 = c->A_RD[i], q2 = c->A_DD[i];
        lq0 += l * q0; q0q0 += q0 * q0; lq1 += l * q1; q0q1 += q0 * q1;
        yq2 += (l - sigma * q0) * q2; q1q1 += q1 * q1; q1q2 += q1 * q2; q2q2 += q2 * q2;
    }
    bq[0] = p0 - lq0 + sigma * q0q0 / 2.0;
    bq[1] = p1 - lq1 + sigma * q0q1;
    bq[2] = p2 - yq2 + sigma * q1q1 / 2.0;
    bq[3] = sigma * q1q2;
    bq[4] = sigma * q2q2 / 2.0;
}

static double poly4(const double *b, double x) { return (((b[4] * x + b[3]) * x + b[2]) * x + b[1]) * x + b[0]; }

/* candidate selection (src/linesearch.jl:58-112). returns -1 where the
 * reference throws (cubic[1] > eps). */
int orc_pick_alpha(const double *bq, double alpha_max, double *alpha_out, double *f_out) {
    const double eps = 2.220446049250313e-16;
    double cubic[4] = {bq[1], 2.0 * bq[2], 3.0 * bq[3], 4.0 * bq[4]};
    if (cubic[0] > eps) return -1;
    double roots[4];
    int nr = 0;
    if (fabs(cubic[3]) < eps) {
        /* quadratic fallback (:70-83): cubic[1:3] ./ cubic[3] */
        double qa = cubic[2], qb = cubic[1], qc = cubic[0];
        if (qa != 0.0) {
            double disc = qb * qb - 4.0 * qa * qc;
            if (disc >= 0) {
                double sq = sqrt(disc);
                double t = -0.5 * (qb + (qb >= 0 ? sq : -sq));
                if (t != 0.0) { roots[nr++] = t / qa; roots[nr++] = qc / t; }
                else { roots[nr++] = 0.0; roots[nr++] = -qb / qa; }
            }
        } else if (qb != 0.0) {
            roots[nr++] = -qc / qb;
        }
    } else {
        nr = cubic_real_roots(cubic, roots);
    }
    roots[nr++] = alpha_max;
    double a_star = 0.0, f_star = bq[0];
    for (int i = 0; i < nr; i++) {
        double x = roots[i];
        if (!(x >= 0.0) || x > alpha_max) continue;
        double fx = poly4(bq, x);
        if (fx < f_star) { f_star = fx; a_star = x; }
    }
    *alpha_out = a_star; *f_out = f_star;
    return 0;
}

/* residual recurrence and bookkeeping after alpha is known
 * (src/linesearch.jl:118-124, shared by the Armijo variant :182-187) */
void orc_commit_step(orc_ctx *c, double alpha) {
    i64 m = c->m;
    for (i64 i = 0; i <= m; i++) c->pvio_raw[i] += alpha * (alpha * c->A_DD[i] + c->A_RD[i]);
    c->obj = c->pvio_raw[m];
    for (i64 i = 0; i < m; i++) {
        double v = c->pvio_raw[i];
        c->pvio[i] = v > c->pvio_lb[i] ? v : c->pvio_lb[i];
    }
}

/* the two A passes shared by both line searches (src/linesearch.jl:9-16) */
void orc_linesearch_passes(orc_ctx *c, const double *Dt) {
    orc_A_uv(c, c->A_RD, c->r, c->Rt, Dt);
    for (i64 i = 0; i <= c->m; i++) c->A_RD[i] *= 2.0;
    orc_A_uv(c, c->A_DD, c->r, Dt, Dt);
}

/* linesearch! (src/linesearch.jl:4-127) */
int orc_linesearch(orc_ctx *c, const double *Dt, double alpha_max, double *alpha, double *Lval, double *bq_out) {
    double bq[5];
    orc_linesearch_passes(c, Dt);
    orc_biquadratic(c, bq);
    if (bq_out) memcpy(bq_out, bq, sizeof(bq));
    int rc = orc_pick_alpha(bq, alpha_max, alpha, Lval);
    if (rc) return rc;
    orc_commit_step(c, *alpha);
    return 0;
}

/* eval_AL closure of linesearch_armijo! (src/linesearch.jl:158-166) */
double orc_eval_AL(const orc_ctx *c, double a) {
    i64 m = c->m;
    double sigma = c->sigma;
    double L = c->obj + a * c->A_RD[m] + a * a * c->A_DD[m];
    for (i64 i = 0; i < m; i++) {
        double g = c->pvio_raw[i] + a * c->A_RD[i] + a * a * c->A_DD[i];
        double t = c->lambda[i] - sigma * g;
        if (c->lambda_ub[i] < t) t = c->lambda_ub[i];
        L += (t * t - c->lambda[i] * c->lambda[i]) / (2.0 * sigma);
    }
    return L;
}

/* linesearch_armijo! (src/linesearch.jl:139-191) */
int orc_linesearch_armijo(orc_ctx *c, const double *Dt, double alpha_max, double *alpha, double *Lval) {
    i64 m = c->m;
    orc_linesearch_passes(c, Dt);
    double L0 = orc_eval_AL(c, 0.0);
    double slope = c->A_RD[m];
    for (i64 i = 0; i < m; i++) slope += c->y[i] * c->A_RD[i];
    double a = alpha_max, La = orc_eval_AL(c, a);
    for (int it = 0; it < 50; it++) {
        if (La <= L0 + 1e-4 * a * slope) break;
        a /= 2.0;
        La = orc_eval_AL(c, a);
    }
    orc_commit_step(c, a);
    *alpha = a; *Lval = La;
    return 0;
}

/* ------------------------------------------------------------------------- */
/* lbfgs_init (src/lbfgs.jl:35-47) */
void orc_lbfgs_init(orc_ctx *c, i64 h) {
    free_lbfgs(c);
    i64 N = c->n * c->r;
    c->h = h; c->latest = h;
    c->hs = (double **)xcalloc(h, sizeof(double *));
    c->hy = (double **)xcalloc(h, sizeof(double *));
    c->hrho = (double *)xcalloc(h, sizeof(double));
    c->ha = (double *)xcalloc(h, sizeof(double));
    for (i64 j = 0; j < h; j++) {
        c->hs[j] = (double *)xcalloc(N, sizeof(double));
        c->hy[j] = (double *)xcalloc(N, sizeof(double));
    }
}
/* lbfgs_clear! (src/lbfgs.jl:52-59) -- note: `latest` is NOT reset */
void orc_lbfgs_clear(orc_ctx *c) {
    i64 N = c->n * c->r;
    for (i64 j = 0; j < c->h; j++) {
        memset(c->hs[j], 0, (size_t)N * sizeof(double));
        memset(c->hy[j], 0, (size_t)N * sizeof(double));
        c->hrho[j] = 0.0; c->ha[j] = 0.0;
    }
}
/* lbfgs_dir! (src/lbfgs.jl:77-124); indices j are 1-based as in the reference */
void orc_lbfgs_dir(orc_ctx *c, double *dir, const double *grad, int negate) {
    i64 N = c->n * c->r, m = c->h;
    memcpy(dir, grad, (size_t)N * sizeof(double));
    if (m == 0) return;
    i64 lst = c->latest, j = lst;
    for (i64 q = 0; q < m; q++) {
        double a = c->hrho[j - 1] * ddot(c->hs[j - 1], dir, N);
        daxpy(-a, c->hy[j - 1], dir, N);
        c->ha[j - 1] = a;
        j -= 1;
        if (j == 0) j = m;
    }
    j = lst % m + 1;
    for (i64 q = 0; q < m; q++) {
        double be = c->hrho[j - 1] * ddot(c->hy[j - 1], dir, N);
        double ga = c->ha[j - 1] - be;
        daxpy(ga, c->hs[j - 1], dir, N);
        j += 1;
        if (j == m + 1) j = 1;
    }
    if (negate) dscal(-1.0, dir, N);
    j = c->latest % m + 1;
    memcpy(c->hy[j - 1], grad, (size_t)N * sizeof(double));
    dscal(-1.0, c->hy[j - 1], N);
}
/* lbfgs_update! (src/lbfgs.jl:129-149): scales dir in place */
void orc_lbfgs_update(orc_ctx *c, double *dir, const double *grad, double step) {
    i64 N = c->n * c->r, m = c->h;
    if (m == 0) return;
    i64 j = c->latest % m + 1;
    dscal(step, dir, N);
    memcpy(c->hs[j - 1], dir, (size_t)N * sizeof(double));
    daxpy(1.0, grad, c->hy[j - 1], N);
    c->hrho[j - 1] = 1.0 / ddot(c->hy[j - 1], c->hs[j - 1], N);
    c->latest = j;
}
double *orc_lbfgs_ptr(orc_ctx *c, int which, i64 j0) { return which == 0 ? c->hs[j0] : c->hy[j0]; }
double orc_lbfgs_rho(const orc_ctx *c, i64 j0) { return c->hrho[j0]; }
i64 orc_lbfgs_latest(const orc_ctx *c) { return c->latest; }

/* ------------------------------------------------------------------------- */
/* smallest eigenvalue of SymTridiagonal(d[0..k), e[0..k-1)) by Sturm bisection
 * (replaces GenericArpack.symeigs(B,1; which=:SA, tol=1e-4), src/coreop.jl:509-511,
 * with an exact answer; see SURVEY 8c: parity "to solver tolerance"). */
double orc_tridiag_mineig(const double *d, const double *e, i64 k) {
    if (k == 1) return d[0];
    double lo = INFINITY, hi = -INFINITY;
    for (i64 i = 0; i < k; i++) {
        double rad = (i > 0 ? fabs(e[i - 1]) : 0.0) + (i < k - 1 ? fabs(e[i]) : 0.0);
        if (d[i] - rad < lo) lo = d[i] - rad;
        if (d[i] + rad > hi) hi = d[i] + rad;
    }
    for (int it = 0; it < 200; it++) {
        double mid = 0.5 * (lo + hi);
        if (mid == lo || mid == hi) break;
        /* count eigenvalues < mid */
        int cnt = 0;
        double q = d[0] - mid;
        if (q < 0) cnt++;
        for (i64 i = 1; i < k; i++) {
            if (q == 0.0) q = 1e-300;
            q = d[i] - mid - e[i - 1] * e[i - 1] / q;
            if (q < 0) cnt++;
        }
        if (cnt >= 1) hi = mid; else lo = mid;
    }
    return 0.5 * (lo + hi);
}

/* approx_mineigval_lanczos (src/coreop.jl:461-514) with the start vector
 * injected (v0, unnormalised: the reference draws randn(n) then normalises).
 * alpha_out/beta_out (length q) receive the UNSHIFTED recurrence coefficients. */
double orc_lanczos(orc_ctx *c, i64 q, const double *v0, i64 *iters_out, double *alpha_out, double *beta_out) {
    i64 n = c->n;
    if (q > n - 1) q = n - 1;
    if (q < 1) q = 1;
    double *alpha = (double *)xcalloc(q, sizeof(double)), *beta = (double *)xcalloc(q, sizeof(double));
    double *v = (double *)xcalloc(n, sizeof(double)), *Av = (double *)xcalloc(n, sizeof(double));
    double *vp = (double *)xcalloc(n, sizeof(double));
    double nv = nrm2(v0, n);
    for (i64 i = 0; i < n; i++) v[i] = v0[i] / nv;
    i64 iter = 0;
    const double eps = 2.220446049250313e-16;
    for (i64 i = 0; i < q; i++) {
        iter++;
        orc_At_right(c, Av, 1, v);
        alpha[i] = ddot(v, Av, n);
        if (i == 0) {
            for (i64 k = 0; k < n; k++) Av[k] -= alpha[i] * v[k];
        } else {
            for (i64 k = 0; k < n; k++) Av[k] -= alpha[i] * v[k] + beta[i - 1] * vp[k];
        }
        beta[i] = nrm2(Av, n);
        if (fabs(beta[i]) < sqrt((double)n) * eps) break;
        for (i64 k = 0; k < n; k++) { Av[k] /= beta[i]; vp[k] = v[k]; v[k] = Av[k]; }
    }
    if (alpha_out) memcpy(alpha_out, alpha, (size_t)q * sizeof(double));
    if (beta_out) memcpy(beta_out, beta, (size_t)q * sizeof(double));
    if (iters_out) *iters_out = iter;
    /* shift by +1, smallest eigenvalue, shift back (:502-513) */
    for (i64 i = 0; i < iter; i++) alpha[i] += 1.0;
    double res = (iter == 1) ? alpha[0] - 1.0 : orc_tridiag_mineig(alpha, beta, iter) - 1.0;
    free(alpha); free(beta); free(v); free(Av); free(vp);
    return res;
}

/* dual_obj, Lanczos branch (src/coreop.jl:376-415). eig_iter follows :402. */
double orc_dual_obj(orc_ctx *c, double trace_bound, i64 iter, const double *v0, double *mineig_out, i64 *q_out) {
    orc_copy2y(c);
    orc_At_preprocess(c);
    double it = (double)(iter > 100 ? iter : 100);
    i64 q = (i64)(2.0 * ceil(sqrt(it) * log((double)c->n)));
    double lam = orc_lanczos(c, q, v0, q_out, NULL, NULL);
    double s = 0.0;
    for (i64 i = 0; i < c->m; i++) s += c->y[i] * c->b[i];
    if (mineig_out) *mineig_out = lam;
    return -s + trace_bound * (lam < 0.0 ? lam : 0.0);
}

/* copy2y_lambda! (src/coreop.jl:238-246): y_i = -lambda_i, y_{m+1} = 1 */
void orc_copy2y_lambda(orc_ctx *c) {
    for (i64 i = 0; i < c->m; i++) c->y[i] = -c->lambda[i];
    c->y[c->m] = 1.0;
}

/* dual value of dual_obj (src/coreop.jl:407-412) for a lambda_min(S) obtained elsewhere: the highprecision branch
 * (:386-400) calls SDP_S_eigval = GenericArpack.symeigs (third-party, see pyoracle.S_eigval) */
double orc_dual_value(const orc_ctx *c, double trace_bound, double mineig) {
    double s = 0.0;
    for (i64 i = 0; i < c->m; i++) s += c->y[i] * c->b[i];
    return -s + trace_bound * (mineig < 0.0 ? mineig : 0.0);
}

/* DIMACS_errors (src/coreop.jl:426-453) after `copy2y_lambda!; At_preprocess!` (:435-436), for the smallest
 * eigenvalue `mineig` of S delivered by SDP_S_eigval (:438-440).  err2 = err3 = 0 (:432-433); err6 takes the sparse
 * part of S only, `dot(var.Rt, var.Rt * aux.sparse_S)` (:448-451). */
void orc_dimacs_errors(const orc_ctx *c, double normb, double normC, double mineig, double *errs) {
    double v2 = 0.0, lb = 0.0, xz = 0.0;
    for (i64 i = 0; i < c->m; i++) { v2 += c->pvio_raw[i] * c->pvio_raw[i]; lb += c->lambda[i] * c->b[i]; }
    if (c->nA > 0) {
        for (i64 j = 0; j < c->n; j++)
            for (i64 k = c->full_colptr[j]; k < c->full_colptr[j + 1]; k++) {
                const double *x = c->Rt + c->full_rowval[k] * c->r, *yj = c->Rt + j * c->r;
                double d = 0.0;
                for (i64 i = 0; i < c->r; i++) d += x[i] * yj[i];
                xz += d * c->S[k];
            }
    }
    double den = 1.0 + fabs(c->obj) + fabs(lb);
    errs[0] = sqrt(v2) / (1.0 + normb);
    errs[1] = 0.0;
    errs[2] = 0.0;
    errs[3] = (mineig < 0.0 ? -mineig : 0.0) / (1.0 + normC);
    errs[4] = (c->obj - lb) / den;
    errs[5] = xz / den;
}

/* dual update (src/sdplr.jl:358-362) */
void orc_dual_update(orc_ctx *c) {
    for (i64 i = 0; i < c->m; i++) {
        double t = c->lambda[i] - c->sigma * c->pvio_raw[i];
        c->lambda[i] = c->lambda_ub[i] < t ? c->lambda_ub[i] : t;
    }
}

/* norm(A, 2) / norm(A, Inf) of B*D*B' (src/structs.jl:61-82), used for normC
 * and by test/symlowrank.jl:12-13 */
double orc_symlowrank_norm(i64 n, i64 s, const double *B, const double *D, int inf) {
    double res = 0.0;
    double *col = (double *)xcalloc(n, sizeof(double));
    for (i64 i = 0; i < n; i++) {
        for (i64 j = 0; j < n; j++) col[j] = 0.0;
        for (i64 k = 0; k < s; k++) {
            double t = D[k] * B[i + k * n];
            for (i64 j = 0; j < n; j++) col[j] += B[j + k * n] * t;
        }
        for (i64 j = 0; j < n; j++) {
            if (inf) { if (fabs(col[j]) > res) res = fabs(col[j]); }
            else res += col[j] * col[j];
        }
    }
    free(col);
    return inf ? res : sqrt(res);
}
