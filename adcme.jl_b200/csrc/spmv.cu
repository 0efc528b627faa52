// spmv.cu — CSR SpMV y = A x in fp64, optionally fused with the dot products <y,w> and <y,y>.
//
// Replaces tf.sparse.sparse_dense_matmul with k == 1 as called from
// /root/reference/src/sparse.jl:227-241 (semantics: Y[i] += v * X[j] over stored entries).
//
// Two kernels, chosen per matrix from facts gathered at assembly time:
//
//  * spmv_tile_kernel<R,U> — for matrices whose R-row tiles hold <= TILE_NNZ entries (stencil /
//    FEM matrices; R = 256, 128 or 64 rows).  A CTA walks tiles grid-stride.  For each tile one
//    elected thread asks the TMA unit (cp.async.bulk, 1-D, mbarrier completion; SASS UBLKCP) to
//    stream the tile's contiguous colind[] and values[] slices into shared memory — fully
//    coalesced 16-byte-aligned bulk reads with no register staging — and then ONE THREAD PER ROW
//    walks its entries out of shared memory.  For banded matrices lane l of a warp handles row
//    r0+l, so at step j every lane touches the same diagonal and the x gathers of a warp are
//    consecutive addresses (2 x 128-byte lines per request) instead of the 7-8 scattered lines an
//    entry-per-lane mapping produces.  Summation inside a row is sequential in stored (column)
//    order, like the reference loop.
//    An optional tile list restricts a launch to a subset of tiles (distributed SpMV: interior
//    tiles run while the halo is in flight, boundary tiles afterwards, dots accumulated in order).
//
//  * spmv_vector_kernel<L> — general fallback (long / irregular rows): L lanes per row straight
//    from global memory, shuffle reduction.
//
// Algorithmic bytes per launch: 12*nnz + 4*(m+1) + 8*n + 8*m.
#include "common.cuh"

namespace ab {

constexpr int TILE_NNZ = 2048;   // staged entries per tile (incl. 16-byte alignment slack)

template <int R, int U>
__global__ void __launch_bounds__(R)
spmv_tile_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                 const double* __restrict__ values, const double* __restrict__ x, double* __restrict__ y,
                 int64_t m, unsigned ntiles, const int32_t* __restrict__ tile_list, int accumulate,
                 const double* __restrict__ dotvec, double* partials, uint32_t* ticket, double* dot_result,
                 const int* __restrict__ done) {
    __shared__ alignas(16) double  s_val[TILE_NNZ];
    __shared__ alignas(16) int32_t s_col[TILE_NNZ];
    __shared__ int32_t s_rp[R + 1];
    __shared__ alignas(8) uint64_t bar;
    __shared__ double s_red[32];
    if (done && *done) return;
    const int t = threadIdx.x;
    if (t == 0) {
        abd::mbar_init(&bar, 1);
        abd::mbar_fence_init();
    }
    double dot = 0.0, dot2 = 0.0;
    unsigned phase = 0;
    for (unsigned ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
        const unsigned tile = tile_list ? (unsigned)tile_list[ti] : ti;
        const int64_t r0 = (int64_t)tile * R;
        const int nr = (int)min((int64_t)R, m - r0);
        if (t < nr) s_rp[t] = rowptr[r0 + t];
        if (t == 0) s_rp[nr] = rowptr[r0 + nr];
        __syncthreads();   // also orders mbar_init before the first use
        const int a4 = s_rp[0] & ~3;
        if (t == 0) {
            const unsigned cnt = (unsigned)((s_rp[nr] - a4 + 3) & ~3);
            abd::mbar_expect_tx(&bar, cnt * 12u);
            if (cnt) {
                abd::bulk_g2s(s_col, colind + a4, cnt * 4u, &bar);
                abd::bulk_g2s(s_val, values + a4, cnt * 8u, &bar);
            }
        }
        abd::mbar_wait(&bar, phase);
        phase ^= 1u;
        if (t < nr) {
            double acc = 0.0;
            const int j0 = s_rp[t] - a4, j1 = s_rp[t + 1] - a4;
            for (int jb = j0; jb < j1; jb += U) {
                double av[U], xv[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int j = jb + u;
                    if (j < j1) {
                        av[u] = s_val[j];
                        xv[u] = __ldg(x + s_col[j]);
                    } else {
                        av[u] = 0.0;
                        xv[u] = 0.0;
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u) acc = fma(av[u], xv[u], acc);
            }
            y[r0 + t] = acc;
            if (dotvec) {
                dot  = fma(acc, dotvec[r0 + t], dot);
                dot2 = fma(acc, acc, dot2);
            }
        }
        __syncthreads();   // smem is rewritten by the next tile
    }
    if (dotvec) {
        double mine[2];
        mine[0] = abd::block_sum<R>(dot, s_red);
        mine[1] = abd::block_sum<R>(dot2, s_red);
        abd::grid_reduce<R, 2>(mine, partials, ticket, s_red, [=](const double(&tot)[2]) {
            dot_result[0] = accumulate ? dot_result[0] + tot[0] : tot[0];
            dot_result[1] = accumulate ? dot_result[1] + tot[1] : tot[1];
        });
    }
}

template <int L>
__global__ void __launch_bounds__(256)
spmv_vector_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colind,
                   const double* __restrict__ values, const double* __restrict__ x, double* __restrict__ y,
                   int64_t m, const double* __restrict__ dotvec, double* partials, uint32_t* ticket,
                   double* dot_result, const int* __restrict__ done) {
    __shared__ double s_red[32];
    if (done && *done) return;
    constexpr int RPB = 256 / L;
    const int sub = threadIdx.x % L;
    double dot = 0.0, dot2 = 0.0;
    for (int64_t rb = (int64_t)blockIdx.x * RPB; rb < m; rb += (int64_t)gridDim.x * RPB) {
        const int64_t row = rb + threadIdx.x / L;   // rows >= m still join the shuffles
        double acc = 0.0;
        if (row < m) {
            const int a = rowptr[row], b = rowptr[row + 1];
            for (int j = a + sub; j < b; j += L) acc = fma(values[j], __ldg(x + colind[j]), acc);
        }
#pragma unroll
        for (int o = L / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (row < m && sub == 0) {
            y[row] = acc;
            if (dotvec) {
                dot  = fma(acc, dotvec[row], dot);
                dot2 = fma(acc, acc, dot2);
            }
        }
    }
    if (dotvec) {
        double mine[2];
        mine[0] = abd::block_sum<256>(dot, s_red);
        mine[1] = abd::block_sum<256>(dot2, s_red);
        abd::grid_reduce<256, 2>(mine, partials, ticket, s_red, [=](const double(&tot)[2]) {
            dot_result[0] = tot[0];
            dot_result[1] = tot[1];
        });
    }
}

static int pick_vector_lanes(const Csr* c) {
    double avg = c->m ? (double)c->nnz / (double)c->m : 0.0;
    if (avg <= 3) return 2;
    if (avg <= 6) return 4;
    if (avg <= 12) return 8;
    if (avg <= 24) return 16;
    return 32;
}

// rows per tile the tile kernel would use for this matrix (0: not eligible)
int spmv_tile_rows(const Csr* c) {
    for (int l = 0; l < 3; ++l)
        if (c->max_tile_nnz[l] <= TILE_NNZ - 8) return 256 >> l;
    return 0;
}

template <int R, int U>
static int launch_tile(const Csr* c, const double* x, double* y, const double* dotvec, double* dot_result,
                       double* partials, uint32_t* ticket, const int* done_flag, int ctas_per_sm,
                       const int32_t* tile_list, int64_t nlist, int accumulate) {
    unsigned ntiles = tile_list ? (unsigned)nlist : (unsigned)((c->m + R - 1) / R);
    if (ntiles == 0) return AB_OK;
    unsigned cap = (unsigned)(sm_count() * ctas_per_sm);
    if (cap > (unsigned)MAX_REDUCE_GRID) cap = MAX_REDUCE_GRID;
    unsigned grid = ntiles < cap ? ntiles : cap;
    spmv_tile_kernel<R, U><<<grid, R, 0, stream()>>>(c->rowptr, c->colind, c->values, x, y, c->m, ntiles, tile_list,
                                                     accumulate, dotvec, partials, ticket, dot_result, done_flag);
    AB_LAUNCH_CHECK();
    return AB_OK;
}

// tile_list != nullptr: only those tiles (requires spmv_tile_rows(c) != 0); accumulate: add the
// fused dots onto dot_result instead of overwriting.
int launch_spmv_part(const Csr* c, const double* x, double* y, const double* dotvec, double* dot_result,
                     double* partials, uint32_t* ticket, const int* done_flag, const int32_t* tile_list,
                     int64_t nlist, int accumulate) {
    if (c->m == 0) return AB_OK;
    // spmv_variant: -1 auto, 0 vector kernel, 1 tile kernel U=8, 2 tile kernel U=4
    int variant = (int)opt("spmv_variant", -1);
    const int R = spmv_tile_rows(c);
    if (tile_list && R == 0) return fail(AB_EUNSUPPORTED, "tile lists need a tile-kernel matrix");
    if (variant < 0 || tile_list) variant = R ? (variant == 2 ? 2 : 1) : 0;
    if (variant >= 1 && R == 0) variant = 0;
    if (variant >= 1) {
        const int cps = (int)opt("spmv_ctas_per_sm", 8);
#define AB_T(RR, UU) launch_tile<RR, UU>(c, x, y, dotvec, dot_result, partials, ticket, done_flag, cps, tile_list, nlist, accumulate)
        if (R == 256) return variant == 2 ? AB_T(256, 4) : AB_T(256, 8);
        if (R == 128) return AB_T(128, 8);
        return AB_T(64, 8);
#undef AB_T
    }
    const int L = pick_vector_lanes(c);
    const int64_t rpb = 256 / L;
    int64_t nblk = (c->m + rpb - 1) / rpb;
    int64_t cap  = (int64_t)sm_count() * 8 * 4;
    if (cap > MAX_REDUCE_GRID) cap = MAX_REDUCE_GRID;
    unsigned grid = (unsigned)(nblk < cap ? nblk : cap);
#define AB_VEC(LL)                                                                                           \
    spmv_vector_kernel<LL><<<grid, 256, 0, stream()>>>(c->rowptr, c->colind, c->values, x, y, c->m, dotvec, \
                                                       partials, ticket, dot_result, done_flag)
    switch (L) {
        case 2: AB_VEC(2); break;
        case 4: AB_VEC(4); break;
        case 8: AB_VEC(8); break;
        case 16: AB_VEC(16); break;
        default: AB_VEC(32); break;
    }
#undef AB_VEC
    AB_LAUNCH_CHECK();
    return AB_OK;
}

int launch_spmv(const Csr* c, const double* x, double* y, const double* dotvec, double* dot_result,
                double* partials, uint32_t* ticket, const int* done_flag) {
    return launch_spmv_part(c, x, y, dotvec, dot_result, partials, ticket, done_flag, nullptr, 0, 0);
}

}  // namespace ab
