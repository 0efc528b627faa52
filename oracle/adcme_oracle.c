/*
 * adcme_oracle.c — CPU restatement of ADCME's sparse linear-algebra custom-op path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  libadcme_b200.so never links,
 * loads or calls anything in this directory.
 *
 * The reference cannot be compiled in the build container (needs Julia, TensorFlow 1.15 headers,
 * Eigen 3.3.7, MPI, Hypre: SURVEY.md F7), so each function below restates the algorithm of a
 * reference function and cites it.  Where the arithmetic lives in a third-party dependency
 * that is not under /root/reference the published algorithm is restated:
 *   - Eigen 3.3.7 (deps/build-standard.jl:104-117)  SparseMatrix::setFromTriplets
 *     (internal::set_from_triplets): bucket triplets by row of a row-major temporary in input
 *     order, collapseDuplicates() adds later duplicates onto the FIRST occurrence in input
 *     order, the transposed copy sorts inner indices;
 *   - tensorflow==1.15.0 (deps/linux.yml:105-106)  SparseTensorDenseMatMul CPU functor and
 *     python/ops/sparse_grad.py::_SparseTensorDenseMatMulGrad.
 * The oracle is pinned against every golden vector the reference's own tests hold for this
 * path (tests/golden/reference_vectors.json, tests/test_oracle_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EXPORT __attribute__((visibility("default")))

/* =====================================================================================
 * SparseAccum — deps/CustomOps/SparseAccumulate/Impl.cpp:11-48, SparseAccumulate.h:18-34
 * ===================================================================================== */
typedef struct {
    int n;
    double tol;
    int** jj; double** vv; int* len; int* cap;
} orc_accum;

EXPORT void* orc_accum_create(int nrow, double tol) {          /* SparseAccum::initialize, Impl.cpp:50-56 */
    orc_accum* a = (orc_accum*)calloc(1, sizeof(orc_accum));
    a->n = nrow; a->tol = tol;
    a->jj = (int**)calloc(nrow > 0 ? nrow : 1, sizeof(int*));
    a->vv = (double**)calloc(nrow > 0 ? nrow : 1, sizeof(double*));
    a->len = (int*)calloc(nrow > 0 ? nrow : 1, sizeof(int));
    a->cap = (int*)calloc(nrow > 0 ? nrow : 1, sizeof(int));
    return a;
}
/* SparseAccum::push_back, Impl.cpp:11-27: keep iff abs(v) > tol; a row beyond n aborts the call */
EXPORT int orc_accum_add(void* h, int row, const int* cols, const double* vals, int N) {
    orc_accum* a = (orc_accum*)h;
    for (int i = 0; i < N; i++) {
        if (fabs(vals[i]) > a->tol) {
            if (row > a->n || row < 1) return -1;
            int r = row - 1;
            if (a->len[r] == a->cap[r]) {
                a->cap[r] = a->cap[r] ? 2 * a->cap[r] : 4;
                a->jj[r] = (int*)realloc(a->jj[r], sizeof(int) * a->cap[r]);
                a->vv[r] = (double*)realloc(a->vv[r], sizeof(double) * a->cap[r]);
            }
            a->jj[r][a->len[r]] = cols[i];
            a->vv[r][a->len[r]] = vals[i];
            a->len[r]++;
        }
    }
    return 0;
}
EXPORT int64_t orc_accum_nnz(void* h) {                          /* SparseAccum::get_n, Impl.cpp:30-35 */
    orc_accum* a = (orc_accum*)h;
    int64_t k = 0;
    for (int i = 0; i < a->n; i++) k += a->len[i];
    return k;
}
/* SparseAccum::copy_to + destroy, Impl.cpp:37-48,107-118: row by row, insertion order, 1-based rows */
EXPORT void orc_accum_copy(void* h, int* rows, int* cols, double* vals) {
    orc_accum* a = (orc_accum*)h;
    int64_t j = 0;
    for (int i = 0; i < a->n; i++) {
        for (int k = 0; k < a->len[i]; k++) {
            rows[j] = i + 1; cols[j] = a->jj[i][k]; vals[j] = a->vv[i][k]; j++;
        }
        free(a->jj[i]); free(a->vv[i]);
    }
    free(a->jj); free(a->vv); free(a->len); free(a->cap); free(a);
}

/* =====================================================================================
 * SparseCompressor — deps/CustomOps/SparseCompress/SparseCompress.h:21-54
 * indices: row-major [N,2] int64, sorted; equal NEIGHBOURS are merged, values added in order.
 * ===================================================================================== */
EXPORT int64_t orc_compress(int64_t N, const int64_t* idx2, const double* v, int64_t* out_idx2, double* out_v) {
    int64_t k = 0;
    for (int64_t i = 0; i < N; i++) {
        if (k > 0 && out_idx2[2 * (k - 1)] == idx2[2 * i] && out_idx2[2 * (k - 1) + 1] == idx2[2 * i + 1]) {
            out_v[k - 1] += v[i];
        } else {
            out_idx2[2 * k] = idx2[2 * i]; out_idx2[2 * k + 1] = idx2[2 * i + 1];
            out_v[k] = v[i];
            k++;
        }
    }
    return k;
}
/* SparseCompressor::backward, SparseCompress.h:45-54: grad_v[i] = grad_nv[segment(i)] */
EXPORT void orc_compress_grad(int64_t N, const int64_t* idx2, const double* grad_nv, double* grad_v) {
    int64_t k = -1;
    for (int64_t i = 0; i < N; i++) {
        if (i == 0 || idx2[2 * i] != idx2[2 * i - 2] || idx2[2 * i + 1] != idx2[2 * i - 1]) k++;
        grad_v[i] = grad_nv[k];
    }
}

/* =====================================================================================
 * MPICreateMatrix — deps/Plugin/MPITensor/MPITensor.h:1-36
 * ===================================================================================== */
EXPORT int64_t orc_coo_to_rows_count(int64_t n, const int64_t* idx2) {     /* MPITensor.h:1-13 */
    if (n == 0) return 0;
    int64_t cur = idx2[0], nrow = 1;
    for (int64_t i = 1; i < n; i++)
        if (idx2[2 * i] != cur) { nrow++; cur = idx2[2 * i]; }
    return nrow;
}
EXPORT void orc_coo_to_rows(int64_t n, const int64_t* idx2, int64_t ilower, int32_t* rows, int32_t* ncols, int32_t* cols) {
    if (n == 0) return;                                                    /* MPITensor.h:15-36 */
    int64_t cur = idx2[0], nrow = 1;
    rows[0] = (int32_t)(cur + ilower); ncols[0] = 1; cols[0] = (int32_t)idx2[1];
    for (int64_t i = 1; i < n; i++) {
        if (idx2[2 * i] == cur) ncols[nrow - 1]++;
        else { rows[nrow++] = (int32_t)(idx2[2 * i] + ilower); ncols[nrow - 1] = 1; cur = idx2[2 * i]; }
        cols[i] = (int32_t)idx2[2 * i + 1];
    }
}

/* =====================================================================================
 * setFromTriplets semantics -> CSR (call sites SparseSolver.h:16-22, SparseFactorization.h:8-12)
 * Bucket by row in input order; a later duplicate (same row, same col) is ADDED onto the first
 * occurrence (collapseDuplicates, left-to-right in input order); then each row is sorted by col.
 * slot[t] (optional) = stored entry of input triplet t.  Returns nnz, or -1 on a range error.
 * ===================================================================================== */
typedef struct { int32_t col; int32_t first; } orc_ent;
static int orc_ent_cmp(const void* a, const void* b) {
    int32_t x = ((const orc_ent*)a)->col, y = ((const orc_ent*)b)->col;
    return (x > y) - (x < y);
}
EXPORT int64_t orc_coo_to_csr(int64_t T, const int64_t* ii, const int64_t* jj, const double* vv, int base,
                              int64_t m, int64_t n, int32_t* rowptr, int32_t* colind, double* values, int32_t* slot) {
    int64_t* cnt = (int64_t*)calloc(m + 1, sizeof(int64_t));
    for (int64_t t = 0; t < T; t++) {
        int64_t r = ii[t] - base, c = jj[t] - base;
        if (r < 0 || r >= m || c < 0 || c >= n) { free(cnt); return -1; }
        cnt[r + 1]++;
    }
    for (int64_t r = 0; r < m; r++) cnt[r + 1] += cnt[r];
    /* pass 2: row buckets in input order (trMat.insertBackUncompressed) */
    int64_t* pos = (int64_t*)malloc(sizeof(int64_t) * (m > 0 ? m : 1));
    memcpy(pos, cnt, sizeof(int64_t) * m);
    int64_t* bucket = (int64_t*)malloc(sizeof(int64_t) * (T > 0 ? T : 1));   /* triplet ids */
    for (int64_t t = 0; t < T; t++) bucket[pos[ii[t] - base]++] = t;
    /* pass 3: collapse duplicates per row, keeping first-occurrence order, summing in input order */
    int32_t* seen = (int32_t*)malloc(sizeof(int32_t) * (n > 0 ? n : 1));
    for (int64_t c = 0; c < n; c++) seen[c] = -1;
    int64_t nnz = 0;
    int64_t* first_of = (int64_t*)malloc(sizeof(int64_t) * (T > 0 ? T : 1)); /* triplet -> entry within row */
    rowptr[0] = 0;
    orc_ent* ents = NULL; double* evals = NULL; int64_t ecap = 0;
    for (int64_t r = 0; r < m; r++) {
        int64_t a = cnt[r], b = cnt[r + 1], ne = 0;
        if (b - a > ecap) { ecap = 2 * (b - a); ents = (orc_ent*)realloc(ents, sizeof(orc_ent) * ecap); evals = (double*)realloc(evals, sizeof(double) * ecap); }
        for (int64_t k = a; k < b; k++) {
            int64_t t = bucket[k];
            int32_t c = (int32_t)(jj[t] - base);
            if (seen[c] >= 0) { evals[seen[c]] += vv[t]; first_of[t] = seen[c]; }
            else { seen[c] = (int32_t)ne; ents[ne].col = c; ents[ne].first = (int32_t)ne; evals[ne] = vv[t]; first_of[t] = ne; ne++; }
        }
        for (int64_t e = 0; e < ne; e++) seen[ents[e].col] = -1;
        /* pass 4: sort inner indices (the transposed copy in Eigen) */
        qsort(ents, ne, sizeof(orc_ent), orc_ent_cmp);
        int32_t* where = (int32_t*)malloc(sizeof(int32_t) * (ne > 0 ? ne : 1));
        for (int64_t e = 0; e < ne; e++) {
            colind[nnz + e] = ents[e].col;
            values[nnz + e] = evals[ents[e].first];
            where[ents[e].first] = (int32_t)(nnz + e);
        }
        if (slot) for (int64_t k = a; k < b; k++) slot[bucket[k]] = where[first_of[bucket[k]]];
        free(where);
        nnz += ne;
        rowptr[r + 1] = (int32_t)nnz;
    }
    free(cnt); free(pos); free(bucket); free(seen); free(first_of); free(ents); free(evals);
    return nnz;
}

/* =====================================================================================
 * tf.sparse.sparse_dense_matmul (TF 1.15 SparseTensorDenseMatMul CPU functor), call sites
 * src/sparse.jl:236,251:  out = 0; for k: out[row_k,:] += val_k * B[col_k,:]   (stored order)
 * and its gradient (_SparseTensorDenseMatMulGrad):
 *   grad_B = A' grad ;  grad_val[k] = sum_c grad[row_k,c] * B[col_k,c]
 * indices: [T,2] int64 0-based; B,out row-major [*,k].
 * ===================================================================================== */
EXPORT void orc_spmm_coo(int64_t T, const int64_t* idx2, const double* val, int64_t m, int64_t k,
                         const double* B, double* out, int adjoint_a) {
    memset(out, 0, sizeof(double) * m * k);
    for (int64_t t = 0; t < T; t++) {
        int64_t r = idx2[2 * t + (adjoint_a ? 1 : 0)], c = idx2[2 * t + (adjoint_a ? 0 : 1)];
        for (int64_t q = 0; q < k; q++) out[r * k + q] += val[t] * B[c * k + q];
    }
}
EXPORT void orc_spmm_grad_values(int64_t T, const int64_t* idx2, int64_t k, const double* grad, const double* B, double* gval) {
    for (int64_t t = 0; t < T; t++) {
        double s = 0;
        for (int64_t q = 0; q < k; q++) s += grad[idx2[2 * t] * k + q] * B[idx2[2 * t + 1] * k + q];
        gval[t] = s;
    }
}

/* CSR SpMV / SpMM used by the CPU baseline and the large-n checks (same per-row sequential order) */
EXPORT void orc_csr_spmm(int64_t m, const int32_t* rowptr, const int32_t* colind, const double* values,
                         int64_t k, const double* X, double* Y) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < m; r++) {
        for (int64_t q = 0; q < k; q++) Y[r * k + q] = 0.0;
        for (int32_t j = rowptr[r]; j < rowptr[r + 1]; j++) {
            const double v = values[j];
            const double* xr = X + (int64_t)colind[j] * k;
            for (int64_t q = 0; q < k; q++) Y[r * k + q] += v * xr[q];
        }
    }
}

/* =====================================================================================
 * SparseSolver forward/backward — deps/CustomOps/SparseSolver/SparseSolver.h:13-70.
 * Small-n direct solve (dense LU with partial pivoting stands in for Eigen::SparseLU; the
 * Python side of the oracle uses scipy SuperLU for mid-size n).  Returns 0 on success, -1 when
 * the factorisation fails (the reference returns false -> "Sparse solver factorization failed.").
 * transpose != 0 assembles the swapped triplets exactly like backward() (SparseSolver.h:53-58).
 * ===================================================================================== */
static int orc_dense_lu_solve(int64_t d, double* A, double* x) {
    for (int64_t k = 0; k < d; k++) {
        int64_t piv = k; double best = fabs(A[k * d + k]);
        for (int64_t r = k + 1; r < d; r++) if (fabs(A[r * d + k]) > best) { best = fabs(A[r * d + k]); piv = r; }
        if (!(best > 0.0)) return -1;
        if (piv != k) {
            for (int64_t c = 0; c < d; c++) { double t = A[k * d + c]; A[k * d + c] = A[piv * d + c]; A[piv * d + c] = t; }
            double t = x[k]; x[k] = x[piv]; x[piv] = t;
        }
        for (int64_t r = k + 1; r < d; r++) {
            double f = A[r * d + k] / A[k * d + k];
            if (f != 0.0) { for (int64_t c = k; c < d; c++) A[r * d + c] -= f * A[k * d + c]; x[r] -= f * x[k]; }
        }
    }
    for (int64_t k = d - 1; k >= 0; k--) {
        double s = x[k];
        for (int64_t c = k + 1; c < d; c++) s -= A[k * d + c] * x[c];
        x[k] = s / A[k * d + k];
    }
    return 0;
}
EXPORT int orc_sparse_solver(const int64_t* ii, const int64_t* jj, const double* vv, int64_t nv,
                             const double* f, int64_t d, int transpose, double* u) {
    double* A = (double*)calloc((size_t)d * d, sizeof(double));
    for (int64_t k = 0; k < nv; k++) {
        int64_t r = (transpose ? jj[k] : ii[k]) - 1, c = (transpose ? ii[k] : jj[k]) - 1;
        A[r * d + c] += vv[k];                     /* setFromTriplets sums duplicates */
    }
    memcpy(u, f, sizeof(double) * d);
    int rc = orc_dense_lu_solve(d, A, u);
    free(A);
    return rc;
}
/* backward, SparseSolver.h:46-70: x = A^{-T} grad_u; grad_f = x; grad_vv[k] = -x[ii_k-1]*u[jj_k-1] */
EXPORT int orc_sparse_solver_grad(const double* grad_u, const double* u, const int64_t* ii, const int64_t* jj,
                                  const double* vv, int64_t nv, int64_t d, double* grad_vv, double* grad_f) {
    int rc = orc_sparse_solver(ii, jj, vv, nv, grad_u, d, 1, grad_f);
    if (rc) return rc;
    for (int64_t k = 0; k < nv; k++) grad_vv[k] = 0.0;
    for (int64_t k = 0; k < nv; k++) grad_vv[k] -= grad_f[ii[k] - 1] * u[jj[k] - 1];
    return 0;
}
/* the per-triplet gradient formula alone (x supplied by any exact solver) */
EXPORT void orc_solver_grad_vv(const double* x, const double* u, const int64_t* ii, const int64_t* jj, int64_t nv, double* grad_vv) {
    for (int64_t k = 0; k < nv; k++) grad_vv[k] = -x[ii[k] - 1] * u[jj[k] - 1];
}

/* =====================================================================================
 * GetPoissonMatrixForward / GetPoissonGrad —
 * docs/src/assets/Codes/MPI/ccode/GetPoissonMatrix.h:15-72, for the single-rank case where
 * colext(i,j) = (i-1)*n + (j-1) inside the grid and -1 outside.  Emits the reference's own
 * per-row order (i+1, i-1, j+1, j-1, centre); cols/val sized 5*n*n, returns entries written;
 * ncols[row] receives the per-row count.
 * ===================================================================================== */
EXPORT int64_t orc_poisson2d_matrix(int64_t n, const double* kext, int64_t* cols, double* val, int32_t* ncols) {
#define KE(i, j) kext[(i) * (n + 2) + (j)]
#define CE(i, j) (((i) >= 1 && (i) <= n && (j) >= 1 && (j) <= n) ? ((i)-1) * n + ((j)-1) : -1)
    int64_t p = 0;
    for (int64_t i = 1; i < n + 1; i++)
        for (int64_t j = 1; j < n + 1; j++) {
            int64_t p0 = p;
            if (CE(i + 1, j) >= 0) { cols[p] = CE(i + 1, j); val[p++] = KE(i + 1, j) + KE(i, j); }
            if (CE(i - 1, j) >= 0) { cols[p] = CE(i - 1, j); val[p++] = KE(i - 1, j) + KE(i, j); }
            if (CE(i, j + 1) >= 0) { cols[p] = CE(i, j + 1); val[p++] = KE(i, j) + KE(i, j + 1); }
            if (CE(i, j - 1) >= 0) { cols[p] = CE(i, j - 1); val[p++] = KE(i, j) + KE(i, j - 1); }
            cols[p] = CE(i, j);
            val[p++] = -(KE(i, j + 1) + KE(i, j - 1) + KE(i + 1, j) + KE(i - 1, j) + 4 * KE(i, j));
            ncols[(i - 1) * n + (j - 1)] = (int32_t)(p - p0);
        }
    return p;
#undef CE
}
EXPORT void orc_poisson2d_grad(int64_t n, const double* xgrad, const double* uext, double* kext) {
#define UE(i, j) uext[(i) * (n + 2) + (j)]
#define XGR(i, j) xgrad[((i)-1) * n + ((j)-1)]
    memset(kext, 0, sizeof(double) * (n + 2) * (n + 2));
    for (int64_t i = 1; i < n + 1; i++)
        for (int64_t j = 1; j < n + 1; j++) {
            KE(i + 1, j) += XGR(i, j) * (UE(i + 1, j) - UE(i, j));
            KE(i - 1, j) += XGR(i, j) * (UE(i - 1, j) - UE(i, j));
            KE(i, j + 1) += XGR(i, j) * (UE(i, j + 1) - UE(i, j));
            KE(i, j - 1) += XGR(i, j) * (UE(i, j - 1) - UE(i, j));
            KE(i, j) += XGR(i, j) * (-4 * UE(i, j) + UE(i + 1, j) + UE(i - 1, j) + UE(i, j + 1) + UE(i, j - 1));
        }
#undef UE
#undef XGR
#undef KE
}

/* =====================================================================================
 * CPU baseline for BASELINE.json's metric ("CG iters/s"): the reference has no iterative
 * solver on this path, so the timed CPU arm is this port of the SAME
 * Jacobi-PCG recurrence the GPU runs, OpenMP over all host threads.
 * ===================================================================================== */
EXPORT void orc_poisson3d_csr_count(int64_t nx, int64_t ny, int64_t nz, int64_t* nnz) {
    *nnz = 7 * nx * ny * nz - 2 * (nx * ny + ny * nz + nx * nz);
}
EXPORT void orc_poisson3d_csr(int64_t nx, int64_t ny, int64_t nz, int32_t* rowptr, int32_t* colind, double* values) {
    const int64_t N = nx * ny * nz, sxy = nx * ny;
    rowptr[0] = 0;
    for (int64_t r = 0; r < N; r++) {
        int64_t x = r % nx, y = (r / nx) % ny, z = r / sxy;
        rowptr[r + 1] = rowptr[r] + 1 + (x > 0) + (x < nx - 1) + (y > 0) + (y < ny - 1) + (z > 0) + (z < nz - 1);
    }
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < N; r++) {
        int64_t x = r % nx, y = (r / nx) % ny, z = r / sxy;
        int32_t p = rowptr[r];
        if (z > 0) { colind[p] = (int32_t)(r - sxy); values[p++] = -1.0; }
        if (y > 0) { colind[p] = (int32_t)(r - nx); values[p++] = -1.0; }
        if (x > 0) { colind[p] = (int32_t)(r - 1); values[p++] = -1.0; }
        colind[p] = (int32_t)r; values[p++] = 6.0;
        if (x < nx - 1) { colind[p] = (int32_t)(r + 1); values[p++] = -1.0; }
        if (y < ny - 1) { colind[p] = (int32_t)(r + nx); values[p++] = -1.0; }
        if (z < nz - 1) { colind[p] = (int32_t)(r + sxy); values[p++] = -1.0; }
    }
}
EXPORT void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
EXPORT int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* Jacobi-PCG, x0 = 0.  tol < 0: run exactly maxit iterations.  Returns iterations done;
 * *relres = ||r||/||b|| from the recurrence.  work: 4*n doubles (r, p, Ap, dinv). */
EXPORT int orc_pcg_jacobi(int64_t n, const int32_t* rowptr, const int32_t* colind, const double* values,
                          const double* b, double* x, double tol, int maxit, double* relres, double* work) {
    double *r = work, *p = work + n, *Ap = work + 2 * n, *dinv = work + 3 * n;
    double rz = 0, bb = 0;
#pragma omp parallel for schedule(static) reduction(+ : rz, bb)
    for (int64_t i = 0; i < n; i++) {
        double d = 0;
        for (int32_t j = rowptr[i]; j < rowptr[i + 1]; j++) if (colind[j] == i) d += values[j];
        dinv[i] = d != 0.0 ? 1.0 / d : 1.0;
        x[i] = 0; r[i] = b[i]; p[i] = dinv[i] * b[i];
        rz += r[i] * p[i]; bb += b[i] * b[i];
    }
    double rr = bb;
    int it = 0;
    if (bb == 0) { if (relres) *relres = 0; return 0; }
    for (; it < maxit;) {
        double pAp = 0;
#pragma omp parallel for schedule(static) reduction(+ : pAp)
        for (int64_t i = 0; i < n; i++) {
            double s = 0;
            for (int32_t j = rowptr[i]; j < rowptr[i + 1]; j++) s += values[j] * p[colind[j]];
            Ap[i] = s; pAp += s * p[i];
        }
        if (!(pAp > 0)) { it = -it - 1; break; }
        double alpha = rz / pAp, rznew = 0; rr = 0;
#pragma omp parallel for schedule(static) reduction(+ : rznew, rr)
        for (int64_t i = 0; i < n; i++) {
            x[i] += alpha * p[i]; r[i] -= alpha * Ap[i];
            rznew += r[i] * dinv[i] * r[i]; rr += r[i] * r[i];
        }
        it++;
        if (tol >= 0 && rr <= tol * tol * bb) break;
        double beta = rznew / rz; rz = rznew;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; i++) p[i] = dinv[i] * r[i] + beta * p[i];
    }
    if (relres) *relres = sqrt(rr / bb);
    return it;
}
