/*
 * adcme_b200.h — flat C-ABI of libadcme_b200.so
 *
 * B200-native (sm_100a) replacement for ADCME.jl's sparse linear-algebra
 * custom-op hot path.  Every entry point takes plain pointers and sizes; no
 * torch / TensorFlow types cross this boundary.  Host code (Julia `ccall`,
 * Python `ctypes`, C) binds exactly these symbols.
 *
 * Conventions (mirroring the reference's own native-op conventions,
 * /root/reference/src/extra.jl:160-235 and deps/CustomOpsTemplate):
 *   - every function returns 0 on success and a negative AB_E* code otherwise;
 *     ab_last_error() returns the thread-local message for the last failure
 *     (the reference surfaces a TF `Status` string as PyCall.PyError,
 *     test/sparse.jl:335-339);
 *   - all array arguments are CALLER-OWNED.  Each pointer may be a device
 *     pointer or a host pointer (detected with cudaPointerGetAttributes);
 *     host buffers are staged over PCIe inside the call;
 *   - outputs are caller-allocated after an explicit size query
 *     (the reference pattern: SparseCompressor.nout, SAMAP[h]->get_n(),
 *     deps/CustomOps/SparseAccumulate/SparseAccumulator.cpp:202);
 *   - cross-call state lives in a process-global table keyed by an integer
 *     handle (reference: SAMAP, deps/CustomOps/SparseAccumulate/SparseAccumulator.cpp:8;
 *     cache1/cache2, deps/CustomOps/SparseFactorizationSolve/lru_cache.cpp:215-216);
 *   - values are fp64; there is no fp32 path (reference: src/variable.jl:742-764).
 *   - There is NO CPU fallback: every compute entry point fails with
 *     AB_ECUDA when no sm_100 device is usable.
 */
#ifndef ADCME_B200_H
#define ADCME_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes --------------------------------------------------- */
#define AB_OK            0
#define AB_EINVAL       -1  /* bad argument (null pointer, negative size, unknown method …)     */
#define AB_EHANDLE      -2  /* unknown / destroyed handle                                       */
#define AB_ECUDA        -3  /* CUDA runtime failure (message has the cudaError string)          */
#define AB_ERANGE       -4  /* a row/column index outside [base, base+dim)                      */
#define AB_ENOCONV      -5  /* iterative solve hit maxit without reaching tol                   */
#define AB_EBREAKDOWN   -6  /* solver breakdown: singular / non-SPD / NaN (reference: "Sparse
                               solver factorization failed.", SparseSolver.cpp:119-138)         */
#define AB_EUNSUPPORTED -7  /* size beyond 32-bit local index space, unknown option …           */
#define AB_ENCCL        -8  /* NCCL failure on the multi-GPU path                                */

/* ---- lifecycle, errors, stream -------------------------------------- */
/* Replaces mpi_init/mpi_finalize/mpi_size (deps/CustomOps/MPI/MPI_Utils.cpp:9-69). */
int         ab_init(int device);            /* bind this process to one GPU; idempotent           */
int         ab_finalize(void);              /* drop every handle and cached workspace             */
int         ab_device_count(void);          /* number of visible CUDA devices (0 if none)         */
const char* ab_last_error(void);
const char* ab_version(void);
int         ab_set_stream(void* cuda_stream); /* launches go to this cudaStream_t (default: 0)   */
int         ab_synchronize(void);
/* tunables: "tol" (default solve tolerance), "maxit", "check_every", "verbose", "refine", and A/B switches of the
 * kernel forms (every one defaults to the measured-best setting; none changes a result beyond fp64 rounding of the
 * fused inner products, and the SpMV output itself is bit-identical under all of them):
 *   SpMV form of the solver   "spmv_rowpat", "spmv_code", "spmv_off8", "spmv_rowwin", "spmv_rowmarch", "spmv_zmarch",
 *                             "spmv_zmarch_g" (tiles per chain position), "spmv_zmarch_len", "spmv_zmarch_balanced",
 *                             "spmv_zmarch_tma" (1: cp.async.bulk, 0: 16-byte cp.async), "spmv_variant"
 *   vector kernels            "ew_ctas_per_sm" (default 4), "cg_defer_x", "cg_scaled", "pdl" (programmatic dependent launch)
 *   distributed iteration     "dist_p2p", "dist_fused_wait", "dist_k3_halo" (K3 stores the next halos), "dist_ll" (all-reduce with
 *                             the flag inside the payload), "dist_epilogue_reduce" (all-reduces in the last CTA of the
 *                             producing kernel), "dist_fused_reduce", "dist_fused_push" (measured slower: off)
 *   assembly                  "asm_rowsort" (row bits first, columns inside each row)
 *   solver fallback           "dense_lu_max" (largest n the dense partial-pivot LU may take, default 8192)
 *   SpMM / SDDMM              "spmm_group", "sddmm_entry"
 *   device                    "l2_fetch_granularity" (cudaLimitMaxL2FetchGranularity; applied at once) */
int         ab_set_option(const char* key, double value);
int         ab_get_option(const char* key, double* value);

/* ---- SparseAssembler (reference a3) ---------------------------------
 * src/sparse.jl:480-534; deps/CustomOps/SparseAccumulate/Impl.cpp:11-118.
 * handle/nrow/row/cols are int32 and rows/cols 1-based exactly as in the
 * reference; a contribution is kept iff abs(v) > tol (strict, applied per
 * contribution BEFORE accumulation, Impl.cpp:16); duplicates are NOT summed;
 * ab_asm_copy dumps row by row in insertion order and destroys the handle. */
int     ab_asm_create(int32_t h, int32_t nrow, double tol);
int     ab_asm_add(int32_t h, int32_t row, const int32_t* cols, const double* vals, int32_t k);
int     ab_asm_add_coo(int32_t h, const int32_t* rows, const int32_t* cols,
                       const double* vals, int64_t k);          /* bulk form of ab_asm_add     */
int64_t ab_asm_nnz(int32_t h);                                   /* <0: error code              */
int     ab_asm_copy(int32_t h, int32_t* rows, int32_t* cols, double* vals);
int     ab_asm_destroy(int32_t h);

/* ---- COO -> CSR assembly with duplicate summation (a1 + a4 + a5) ----
 * SparseTensor(ii,jj,vv,m,n) (src/sparse.jl:62-76) followed by what every
 * consumer does to it: Eigen setFromTriplets (SparseSolver.h:16-22) — sum
 * duplicates in input order, sorted column indices within each row.
 * `base` is 1 for the op boundary (find(), SparseSolver.h:19) and 0 for
 * tf.SparseTensor indices.  The handle keeps rowptr/colind/values AND the
 * triplet -> stored-entry map, so gradients can be delivered per INPUT triplet
 * in the caller's original order (SparseCompress.h:45-54).
 * Errors: AB_ERANGE if an index is out of range (the reference is UB there). */
int ab_csr_from_coo(int64_t T, const int64_t* ii, const int64_t* jj, const double* vv,
                    int64_t m, int64_t n, int base, int64_t* out_h);
int ab_csr_from_coo32(int64_t T, const int32_t* ii, const int32_t* jj, const double* vv,
                      int64_t m, int64_t n, int base, int64_t* out_h);
/* tf.sparse.reorder of the constructor (src/sparse.jl:62-76): stable row-major sort of the triplets,
 * duplicates kept — the order find()/rows()/cols()/values() hand to every op.  Outputs may be NULL;
 * perm[k] = input position of output k (so values with gradients can be permuted by the host AD). */
int ab_coo_reorder(int64_t T, const int64_t* ii, const int64_t* jj, const double* vv, int64_t m, int64_t n,
                   int base, int64_t* oii, int64_t* ojj, double* ovv, int64_t* perm);
/* adopt an existing 0-based CSR (copied); T == nnz and the triplet map is the identity */
int ab_csr_create(int64_t m, int64_t n, int64_t nnz, const int32_t* rowptr,
                  const int32_t* colind, const double* values, int64_t* out_h);
int ab_csr_query(int64_t h, int64_t* m, int64_t* n, int64_t* nnz, int64_t* T);
int ab_csr_export(int64_t h, int32_t* rowptr /*[m+1]*/, int32_t* colind /*[nnz]*/,
                  double* values /*[nnz]*/);              /* any output may be NULL              */
int ab_csr_export_map(int64_t h, int32_t* slot /*[T]*/); /* slot[t] = stored entry of triplet t */
/* re-reduce new triplet values through the cached map: no re-sort (a11-style reuse) */
int ab_csr_update_values(int64_t h, const double* vv /*[T]*/);
int ab_csr_transpose(int64_t h, int64_t* out_h);          /* explicit A' (src/sparse.jl:220-225) */
int ab_csr_destroy(int64_t h);

/* ---- compress (a4) ----------------------------------------------------
 * deps/CustomOps/SparseCompress/SparseCompress.h:21-54.  idx2 is the 0-based
 * row-major [N,2] int64 index matrix of a tf.SparseTensor, SORTED (precondition
 * inherited from the reference); equal neighbours are merged, values summed
 * left to right. */
int ab_compress_count(int64_t N, const int64_t* idx2, int64_t* nout);
int ab_compress(int64_t N, const int64_t* idx2, const double* v,
                int64_t* out_idx2 /*[nout,2]*/, double* out_v /*[nout]*/);
int ab_compress_grad(int64_t N, const int64_t* idx2, const double* grad_nv /*[nout]*/,
                     double* grad_v /*[N]*/);

/* ---- mpi_create_matrix (a6): sorted COO -> (rows, ncols, cols) --------
 * deps/Plugin/MPITensor/MPITensor.h:1-36. */
int ab_coo_to_rows_count(int64_t n, const int64_t* idx2, int64_t* nrow);
int ab_coo_to_rows(int64_t n, const int64_t* idx2, int64_t ilower,
                   int32_t* rows /*[nrow]*/, int32_t* ncols /*[nrow]*/, int32_t* cols /*[n]*/);

/* ---- SpMV / SpMM and gradients (a7, a8) -------------------------------
 * src/sparse.jl:227-264 -> tf.sparse.sparse_dense_matmul.  X, Y are ROW-MAJOR
 * [ncols,k] / [nrows,k] fp64 (k == 1 is SpMV).  transA != 0 multiplies by A'
 * (the `o*s` path, sparse.jl:251).
 * grad_values is per INPUT triplet: dvv[t] = <dY[i_t,:], X[j_t,:]>.
 * grad_dense is dX = A' dY. */
int ab_spmm(int64_t h, int transA, const double* X, int64_t k, double* Y);
int ab_spmm_grad_values(int64_t h, const double* dY, const double* X, int64_t k,
                        double* dvv /*[T]*/);
int ab_spmm_grad_dense(int64_t h, const double* dY, int64_t k, double* dX);

/* ---- sparse `\` and its adjoint (a9, a10, a11) ------------------------
 * src/sparse.jl:413-445; SparseSolver.h:13-70; Solve.h:3-33.
 * method: "CG" | "BiCGSTAB" | "auto", plus the reference's strings mapped as
 *   "SimplicialLLT","SimplicialLDLT" -> CG (caller asserts SPD);
 *   "SparseLU","SparseQR","auto"    -> CG FIRST when the matrix is bitwise symmetric with a
 *       positive diagonal (one cached check per set of values — the FEM / stencil systems
 *       this path exists for; the adjoint then needs no transpose), BiCGSTAB otherwise or
 *       when CG breaks down / stalls, and a one-CTA dense partial-pivot LU when BiCGSTAB
 *       fails too and n <= 2048, the same elimination spread over the whole GPU for n <= "dense_lu_max"
 *       (option, default 8192, at most 32768; O(n^3): about a second at 8192 — a last resort, a direct
 *       solver never fails on a regular matrix).
 * Both Krylov solvers are Jacobi-preconditioned: the supported matrix class above "dense_lu_max"
 * unknowns is "Jacobi-Krylov converges" (SPD, diagonally dominant, mildly unsymmetric);
 * strongly indefinite / saddle-point systems of that size return AB_ENOCONV / AB_EBREAKDOWN
 * where the reference's SparseLU would succeed.  tol <= 0 / maxit <= 0 pick the options
 * "tol" / "maxit".  The iteration stops on the recurrence residual; the answer is then VETTED
 * by the true residual ||f - A u||_2 / ||f||_2 (one more SpMV) and refined by correction
 * solves when it had drifted: AB_OK means true relres <= 10*tol, and *relres receives the
 * true value.  iters / relres may be NULL.  Non-convergence returns AB_ENOCONV (u still
 * holds the last iterate), breakdown AB_EBREAKDOWN — never a silent NaN.
 * ab_solve_grad: x = A^{-T} grad_u ; grad_f = x ; grad_vv[t] = -x[i_t]*u[j_t]
 * for every input triplet t (duplicates each get the same formula). */
int ab_solve(int64_t h, const char* method, const double* f, double* u,
             double tol, int maxit, int* iters, double* relres);
int ab_solve_grad(int64_t h, const char* method, const double* grad_u, const double* u,
                  double* grad_vv /*[T] or NULL*/, double* grad_f /*[n] or NULL*/,
                  double tol, int maxit, int* iters, double* relres);
/* run exactly `iters` CG iterations with the convergence test disabled
 * (benchmark entry: BASELINE.json "CG iters/s"); returns relres at the end */
int ab_cg_fixed(int64_t h, const double* f, double* u, int iters, double* relres);

/* One-shot forms with the reference op signatures (1-based int64 ii/jj):
 * SparseSolver / SparseSolverGrad, deps/CustomOps/SparseSolver/SparseSolver.cpp:21-60 */
int ab_sparse_solver(const int64_t* ii, const int64_t* jj, const double* vv, int64_t nv,
                     const double* f, int64_t d, const char* method, double* u);
int ab_sparse_solver_grad(const double* grad_u, const double* u,
                          const int64_t* ii, const int64_t* jj, const double* vv, int64_t nv,
                          const double* f, int64_t d, const char* method,
                          double* grad_vv /*[nv]*/, double* grad_f /*[d]*/);

/* factorize / solve with an id-keyed cache (a11):
 * SparseFactorization.h:4-34, Solve.h:3-33.  "Factorizing" here assembles and
 * caches CSR(A), CSR(A'), the Jacobi diagonals and the solver workspace. */
int ab_factorize(const int64_t* ii, const int64_t* jj, const double* vv, int64_t nv,
                 int64_t d, int64_t max_cache_size, int64_t* id);
int ab_factorized_solve(int64_t id, const double* rhs, int64_t d, double* out);
int ab_factorized_solve_grad(int64_t id, const double* grad_out, const double* out,
                             const int64_t* ii, const int64_t* jj, int64_t nv, int64_t d,
                             double* grad_rhs, double* grad_vv);

/* ---- stencil matrix generators (N1) -----------------------------------
 * 3-D 7-point Dirichlet Laplacian (diag 6, off -1) on an nx*ny*nz grid, rows
 * [row_lo,row_hi) only (a z-slab when row_lo/row_hi are multiples of nx*ny),
 * built direct-to-CSR on the device.  Column ids are GLOBAL. */
int ab_poisson3d_csr(int64_t nx, int64_t ny, int64_t nz, int64_t row_lo, int64_t row_hi,
                     int64_t* out_h);
/* the same pattern with variable coefficients: finite-volume -div(kappa grad u), Dirichlet walls,
 * face weight (kappa_i + kappa_j)/2 (wall faces: kappa_i); kappa is the GLOBAL cell field
 * [nx*ny*nz] (3-D analogue of GetPoissonMatrix.h:24-45, sign chosen SPD).          */
int ab_poisson3d_kappa_csr(int64_t nx, int64_t ny, int64_t nz, const double* kappa,
                           int64_t row_lo, int64_t row_hi, int64_t* out_h);
/* 2-D 5-point variable-coefficient operator from a halo-extended kext[(n+2)*(n+2)]:
 * docs/src/assets/Codes/MPI/ccode/GetPoissonMatrix.h:15-51, with sign so that
 * sign=+1 reproduces the reference values and sign=-1 the SPD matrix (-A).
 * Dirichlet: neighbours outside the n*n grid are dropped.                      */
int ab_poisson2d_csr(int64_t n, const double* kext, int sign, int64_t* out_h);
int ab_poisson2d_update(int64_t h, int64_t n, const double* kext, int sign);
/* adjoint of the generator: GetPoissonMatrix.h:55-72 (xgrad = A^{-T}g, uext = halo-extended u) */
int ab_poisson2d_grad(int64_t n, const double* xgrad, const double* uext, int sign,
                      double* grad_kext /*[(n+2)*(n+2)]*/);

/* ---- structural edits between assembly and solve and the remaining
 *      SparseTensor algebra (N4) — csrc/structure.cu ---------------------------------
 * The reference ops below return triplet lists whose length is known only after the op ran (they
 * size allocate_output from op-private state, e.g. IJV.get_size()).  Here such a result is parked
 * on the device under a TRIPLET HANDLE: ab_trip_size -> caller allocates -> ab_trip_fetch (which
 * frees the handle, like ab_asm_copy), or ab_csr_from_trip assembles it without leaving the GPU. */
int64_t ab_trip_size(int64_t th);                                   /* <0: error code            */
int     ab_trip_fetch(int64_t th, int64_t* ii, int64_t* jj, double* vv);   /* outputs may be NULL */
int     ab_trip_fetch_idx2(int64_t th, int64_t* idx2 /*[n,2]*/, double* vv);
int     ab_trip_destroy(int64_t th);
int     ab_csr_from_trip(int64_t th, int64_t m, int64_t n, int base, int64_t* out_h); /* consumes th */

/* zero_out_row (src/sparse.jl:790-797; deps/CustomOps/ZeroOutRow/ZeroOutRow.h:28-57):
 * idx2 = 0-based [N,2] indices, bd = 1-based rows to drop; kept entries keep their order.
 * grad: grad_v[i] = grad_oval[k] for the k-th kept entry, 0 for dropped ones. */
int ab_zero_out_row(int64_t N, const int64_t* idx2, const double* v, const int64_t* bd, int64_t nbd,
                    int64_t* out_th);
int ab_zero_out_row_grad(int64_t N, const int64_t* idx2, const int64_t* bd, int64_t nbd,
                         const double* grad_oval, int64_t nout, double* grad_v /*[N]*/);

/* sparse_scatter_update, A[ii,jj] = B (src/sparse.jl:340-364;
 * deps/CustomOps/SparseScatterUpdate/SparseScatterUpdate.h:20-60).  All indices 1-based;
 * (uii,ujj) are positions inside the block.  Output = entries of A outside ii x jj in input
 * order, followed by (ii[uii-1], jj[ujj-1], uvv) in input order.
 * grad: grad_ovv = grad_out for kept entries / 0 for replaced ones; grad_uvv = the tail. */
int ab_scatter_update(int64_t on, const int64_t* oii, const int64_t* ojj, const double* ovv,
                      int64_t un, const int64_t* uii, const int64_t* ujj, const double* uvv,
                      int64_t m, int64_t n, const int64_t* ii, int64_t ni, const int64_t* jj, int64_t nj,
                      int64_t* out_th);
int ab_scatter_update_grad(int64_t on, const int64_t* oii, const int64_t* ojj, int64_t un,
                           int64_t m, int64_t n, const int64_t* ii, int64_t ni, const int64_t* jj, int64_t nj,
                           const double* grad_out, int64_t nout,
                           double* grad_ovv /*[on] or NULL*/, double* grad_uvv /*[un] or NULL*/);

/* sparse_indexing, A[ii,jj] (src/sparse.jl:303-327; deps/CustomOps/SparseIndexing/SparseIndexing.h:23-83).
 * h = assembled A (duplicates summed, as the op's setFromTriplets does); ii, jj 1-based, without
 * repeats.  Output triplets are 1-based positions inside the block, ordered by position in jj, then
 * by ascending row of A (the reference's column walk).
 * grad: the gradient of output k goes to the LAST input triplet of A at (ii[ii0-1], jj[jj0-1]). */
int ab_sparse_indexing(int64_t h, const int64_t* ii, int64_t ni, const int64_t* jj, int64_t nj,
                       int64_t* out_th);
int ab_sparse_indexing_grad(int64_t h, const int64_t* ii, int64_t ni, const int64_t* jj, int64_t nj,
                            const int64_t* ii0, const int64_t* jj0, const double* grad_vv0, int64_t nout,
                            double* grad_vv1 /*[T]*/);

/* stored entries of a CSR handle as triplets in row-major order (find(), src/sparse.jl:93-99) */
int ab_csr_export_coo(int64_t h, int64_t* ii, int64_t* jj, double* vv, int base);
/* out[t] = vv[t] * d[idx[t] - base]: diag_sparse_mat_mul (idx = rows) / sparse_diag_mat_mul
 * (idx = cols), deps/CustomOps/SparseMatMul/SparseMatMul.h:34-52 */
int ab_coo_scale(int64_t T, const int64_t* idx, const double* vv, const double* d, int64_t nd,
                 int base, double* out);
/* dense_to_sparse (src/sparse.jl:78-86): the nonzeros (a != 0) of a row-major [m,n] matrix as 1-based
 * triplets in row-major order -> triplet handle */
int ab_dense_to_coo(int64_t m, int64_t n, const double* a, int64_t* out_th);
/* vcat / hcat (src/sparse.jl:227-262 `_sparse_concate`; deps/CustomOps/SparseConcate/SparseConcate.h:11-27):
 * out = [T1; T2 shifted by m1 rows (vcat) or ncol1 columns (hcat)], input order kept */
int ab_coo_concat(int64_t n1, const int64_t* ii1, const int64_t* jj1, const double* vv1,
                  int64_t n2, const int64_t* ii2, const int64_t* jj2, const double* vv2,
                  int64_t m1, int64_t ncol1, int hcat, int64_t* oii, int64_t* ojj, double* ovv /*[n1+n2]*/);
/* Array(A) (src/sparse.jl:185-193; deps/CustomOps/SparseToDense/SparseToDense.h): row-major [m,n],
 * duplicates summed (at assembly); grad_v[t] = grad_out[i_t, j_t] per input triplet (0-based idx2) */
int ab_csr_to_dense(int64_t h, double* out /*[m*n]*/);
int ab_dense_grad_values(int64_t T, const int64_t* idx2, int64_t m, int64_t n, const double* grad_out,
                         double* grad_v /*[T]*/);
/* C = alpha*A + beta*B on the union pattern (A + B: src/sparse.jl:203-218 -> tf.sparse.add) */
int ab_csr_axpby(int64_t ha, double alpha, int64_t hb, double beta, int64_t* out_h);
/* C = A*B (sparse_sparse_mat_mul, SparseMatMul.h:11-32; src/sparse.jl:579-595).  transpose_out != 0
 * returns C' — whose row-major entries are C's column-major ones, the order the reference op emits.
 * Each C(i,j) is summed in k-ascending order (Eigen's column-major product order). */
int ab_spgemm(int64_t ha, int64_t hb, int transpose_out, int64_t* out_h);

/* ---- trisolve (src/sparse.jl:731-739; deps/CustomOps/TriSolve/TriSolve.h:1-37) — csrc/trisolve.cu -----
 * a[n-1] sub-diagonal (a[i-1] multiplies x[i-1] in row i), b[n] diagonal, c[n-1] super-diagonal, d[n] rhs.
 * Device form: parallel cyclic reduction (ceil(log2 n) streaming sweeps).  A zero pivot / non-finite
 * result returns AB_EBREAKDOWN.  Gradient: grad_d = solve of the transposed system with grad_x;
 * grad_a[i-1] = -grad_d[i] x[i-1], grad_b[i] = -grad_d[i] x[i], grad_c[i] = -grad_d[i] x[i+1]. */
int ab_tri_solve(int64_t n, const double* a, const double* b, const double* c, const double* d, double* x);
int ab_tri_solve_grad(int64_t n, const double* grad_x, const double* x, const double* a, const double* b,
                      const double* c, double* grad_a, double* grad_b, double* grad_c, double* grad_d);

/* ---- Newton-Raphson inner loop on the device — csrc/newton.cu ----------
 * src/optim.jl:462-592 (newton_raphson), :364-426 (backtracking).  The residual/Jacobian callback
 * stays in the host language; the linear solve, the update u - step*delta, ||r|| and the
 * line-search reductions run here on device-resident vectors.
 * ab_newton_step: delta = J \ r with any `\` method (J = handle h, LM shift already added by the
 * caller as J + mu*I); u_new = u - step*delta (may alias u); *rnorm = ||r||_2.  delta may be NULL. */
int ab_newton_step(int64_t h, const char* method, const double* r, const double* u, double step,
                   double* u_new, double* delta, double* rnorm, double tol, int maxit, int* lin_iters);
int ab_vec_axpby(int64_t n, double a, const double* x, double b, const double* y, double* out); /* a*x+b*y */
int ab_vec_dot(int64_t n, const double* x, const double* y, double* result);   /* fixed-order reduction */

/* ---- row-block distributed CSR + solve (a12) --------------------------
 * Replaces mpi_SparseTensor `\` (src/mpi.jl:420-508; MPITensor.h:103-152):
 * one process per GPU, rank r owns rows [ilower, iupper] (Hypre IJ layout,
 * inclusive), local CSR with GLOBAL column ids.  NCCL carries the CG inner
 * products; halo planes move over NVLink.  The host bootstraps NCCL by passing
 * the 128-byte ncclUniqueId made by ab_dist_unique_id on rank 0.           */
int ab_dist_unique_id(void* id128);
int ab_dist_init(int rank, int world, const void* id128);
int ab_dist_finalize(void);
int ab_dist_csr_create(int64_t h_local, int64_t ilower, int64_t iupper, int64_t N,
                       int64_t* out_dh);
/* rows owned, halo columns referenced, stored entries of this rank's block (any output may be NULL) */
int ab_dist_query(int64_t dh, int64_t* nloc, int64_t* nhalo, int64_t* nnz_local);
int ab_dist_spmv(int64_t dh, const double* x_local, double* y_local);
/* method (options.mpi.solver, src/options.jl:54-58): "CG" (caller asserts SPD); "BiCGSTAB" or the reference's
 * "GMRES" -> right-Jacobi BiCGSTAB (any regular matrix); "BoomerAMG" or "auto" -> CG when the matrix is
 * symmetric across ranks (bitwise check, cached per set of values) with a positive diagonal, BiCGSTAB
 * otherwise or when CG fails.  Every answer is vetted by the TRUE residual ||f - A u|| / ||f|| (one more
 * distributed SpMV) and refined by correction solves when the recurrence residual had drifted: AB_OK means
 * true relres <= 10 * tol, anything else returns AB_ENOCONV / AB_EBREAKDOWN (never silent).  tol <= 0 -> 1e-10
 * (the reference's Hypre tolerance, MPITensor.h:66,91); *relres receives the true relative residual.         */
int ab_dist_solve(int64_t dh, const char* method, const double* f_local, double* u_local,
                  double tol, int maxit, int* iters, double* relres);
int ab_dist_cg_fixed(int64_t dh, const double* f_local, double* u_local, int iters,
                     double* relres);
/* adjoint of the distributed solve (collective): x = A^{-T} grad_u — solved with A itself when the matrix is
 * symmetric across ranks, otherwise with the distributed transpose (built once per set of values) —, grad_f = x on the
 * local rows, grad_vals[e] = -x[row_e]*u[col_e] for every stored entry e of this rank's rows in the order
 * of the local CSR handle (u at off-rank columns arrives by one halo exchange).  Either output may be NULL. */
int ab_dist_solve_grad(int64_t dh, const double* grad_u_local, const double* u_local,
                       double* grad_vals_local /*[nnz_local]*/, double* grad_f_local /*[nloc]*/,
                       double tol, int maxit, int* iters, double* relres);
/* SparseTensor(sp::mpi_SparseTensor) — MPIGetMatrix (MPITensor.h:39-52, src/mpi.jl:502-504): this rank's rows
 * as COO; idx2 is the row-major [nnz_local,2] int64 index matrix (row relative to ilower, GLOBAL column),
 * entries in the order of the local CSR the handle was created from.  Either output may be NULL.            */
int ab_dist_export_coo(int64_t dh, int64_t* idx2 /*[nnz_local,2]*/, double* vals /*[nnz_local]*/);
/* adjoint(A::mpi_SparseTensor) — MPITensorTranspose.h:79-141, src/mpi.jl:544-563 (collective): every stored
 * entry travels to the rank that owns its column (one all-to-all-v of COO blocks); the new handle holds this
 * rank's rows of A' under the same row partition.                                                           */
int ab_dist_transpose(int64_t dh, int64_t* out_dh);
int ab_dist_destroy(int64_t dh);

/* ---- timers (reference: MPITensor_Solve_Timer_*, MPITensor.cpp:15-21) -- */
#define AB_TIMER_SOLVE    0
#define AB_TIMER_SPMM     1
#define AB_TIMER_ASSEMBLY 2
int    ab_timer_reset(void);
double ab_timer_get(int which);   /* accumulated device seconds */

/* ---- measurement hooks (bench.py only) --------------------------------
 * Launch the named kernel `reps` times back to back on the current stream and
 * return the average launch duration in ms from CUDA events on that stream. */
int ab_bench_spmv(int64_t h, const double* x, double* y, int reps, int with_dot, double* ms);
int ab_bench_cg_kernels(int64_t h, int reps, double* ms_spmv_dot, double* ms_update,
                        double* ms_direction);
int64_t ab_launch_count(void);   /* kernels launched by this library since ab_init */
/* test hook: y = (S A S) x, S = diag(A)^-1/2, through exactly the SpMV form (row-pattern / pattern-code /
 * offset-dictionary / int32) the CG solver uses for this handle, without the fused inner products */
int ab_spmv_solver_form(int64_t h, const double* x, double* y);
/* which SpMV form a handle currently uses: "noff" (offset-dictionary size, 0 = int32 colind),
 * "ncode" (pattern-code dictionary size of the solver values, 0 = none), "tile_rows",
 * "max_row_nnz", "scaled" (1 when CG runs on the symmetrically scaled copy) */
int ab_csr_get_info(int64_t h, const char* key, double* value);
/* test hook, HOST ONLY (no device needed): the work-item plan of the marching SpMV kernel for a matrix of m rows whose
 * stencil planes are M rows apart, on a device with `grid` SMs — csrc/zmarch_plan.h.  excl: one flag per 1024-row tile
 * that must stay out of every item (NULL: none).  fits_g_max: the widest position (tiles) the kernel can stage; fits_l2:
 * the operand is about L2-sized.  g_opt / len_opt / bal_opt: overrides (0 / 0 / -1 = automatic).  Outputs: out3 = {G,
 * balanced, segment length}, and up to `cap` items (first row, chain length, rows per position), the CTA offsets of the
 * balanced mode (grid + 1 entries at most) and the uncovered tiles; counts in n_items / n_cta / n_rest.  AB_EINVAL when
 * `cap` is too small. */
int ab_test_zmarch_plan(int64_t M, int64_t m, int grid, const uint8_t* excl, int64_t n_excl, int fits_g_max, int fits_l2,
                        int g_opt, int len_opt, int bal_opt, int* out3, int64_t cap, uint32_t* first, int32_t* count,
                        int32_t* rows, int32_t* cta_first, int32_t* rest, int64_t* n_items, int64_t* n_cta, int64_t* n_rest);

#ifdef __cplusplus
}
#endif
#endif /* ADCME_B200_H */
