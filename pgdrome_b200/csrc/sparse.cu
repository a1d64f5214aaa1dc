// CSR kernels: SpMV, fused SpMV+dot, bilinear functional x^T A y (one template, three modes), Dirichlet.
// Three SpMV cores by size: k_spmv_bulk (n >= 32 768 rows: TMA ring, spmv_bulk.cuh), k_spmv_stream (n >= 8 192:
// register-staged row blocks, spmv_stream.cuh) and k_spmv (sub-warp per row, latency-bound small systems).
// Matrix values/indices are streamed once, x is gathered through the read-only path and stays L2-resident.
#include "common.cuh"
#include "spmv_bulk.cuh"
#include "spmv_stream.cuh"

// TMA-pipelined SpMV (spmv_bulk.cuh).  MODE as in k_spmv below.
#define SPMV_BULK_THREADS 512
template <int MODE>
__global__ void __launch_bounds__(SPMV_BULK_THREADS, BK_CTAS_PER_SM) k_spmv_bulk(const int32_t* __restrict__ rowptr,
                                                             const int32_t* __restrict__ colidx,
                                                             const double* __restrict__ vals, const double* __restrict__ x,
                                                             double* __restrict__ y, const double* __restrict__ w, int64_t n,
                                                             double* dot, double* part, unsigned int* counter) {
    extern __shared__ __align__(128) unsigned char bk_smem[];
    double acc = 0.0;
    auto epi = [&](int64_t row, double s) {
        if (MODE != 2) y[row] = s;
        if (MODE != 0) acc = fma(__ldg(&w[row]), s, acc);
    };
    bk_spmv_rows<BkGatherX, decltype(epi)&, SPMV_BULK_THREADS>(rowptr, colidx, vals, n, BkGatherX{x}, epi, bk_smem);
    if (MODE != 0) {
        acc = block_sum(acc);
        double v[1] = {acc};
        grid_sum_finish<1>(v, part, counter, dot, blockIdx.x, gridDim.x);
    }
}

// MODE as in k_spmv below.  Persistent grid, one ST_ROWS row block per CTA pass.
template <int MODE>
__global__ void __launch_bounds__(ST_THREADS) k_spmv_stream(const int32_t* __restrict__ rowptr,
                                                             const int32_t* __restrict__ colidx,
                                                             const double* __restrict__ vals, const double* __restrict__ x,
                                                             double* __restrict__ y, const double* __restrict__ w, int64_t n,
                                                             double* dot, double* part, unsigned int* counter) {
    __shared__ double s_prod[ST_TILE];
    __shared__ int s_rp[ST_ROWS + 1];
    const GatherX g{x};
    const int64_t nblk = (n + ST_ROWS - 1) / ST_ROWS;
    double acc = 0.0;
    for (int64_t rb = blockIdx.x; rb < nblk; rb += gridDim.x) {
        const int64_t r0 = rb * ST_ROWS;
        const int nr = (int)min((int64_t)ST_ROWS, n - r0);
        const double s = stream_rowblock(rowptr, colidx, vals, g, r0, nr, s_prod, s_rp);
        if ((int)threadIdx.x < nr) {
            if (MODE != 2) y[r0 + threadIdx.x] = s;
            if (MODE != 0) acc = fma(__ldg(&w[r0 + threadIdx.x]), s, acc);
        }
    }
    if (MODE != 0) {
        acc = block_sum(acc);
        double v[1] = {acc};
        grid_sum_finish<1>(v, part, counter, dot, blockIdx.x, gridDim.x);
    }
}

template <int LPR>
__device__ __forceinline__ double row_dot(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                          const double* __restrict__ vals, const double* __restrict__ x, int64_t row,
                                          int64_t n, int lane) {
    // every lane of the warp takes part in the shuffles; rows past the end get an empty range
    int k0 = 0, k1 = 0;
    if (row < n) {
        k0 = __ldg(&rowptr[row]);
        k1 = __ldg(&rowptr[row + 1]);
    }
    double s = 0.0;
    for (int k = k0 + lane; k < k1; k += LPR) s += ld_stream(&vals[k]) * __ldg(&x[ld_stream(&colidx[k])]);
#pragma unroll
    for (int o = LPR >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, LPR);
    return s;
}

// MODE 0: y = A x ; MODE 1: y = A x, dot = w^T y ; MODE 2: dot = w^T A x (no store)
template <int LPR, int MODE>
__global__ void __launch_bounds__(256) k_spmv(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                              const double* __restrict__ vals, const double* __restrict__ x,
                                              double* __restrict__ y, const double* __restrict__ w, int64_t n, double* dot,
                                              double* part, unsigned int* counter) {
    constexpr int RPB = 256 / LPR;  // rows per block per sweep
    const int lane = threadIdx.x % LPR;
    const int sub = threadIdx.x / LPR;
    double acc = 0.0;
    for (int64_t base = (int64_t)blockIdx.x * RPB; base < n; base += (int64_t)gridDim.x * RPB) {
        int64_t row = base + sub;
        double s = row_dot<LPR>(rowptr, colidx, vals, x, row, n, lane);
        if (row < n && lane == 0) {
            if (MODE != 2) y[row] = s;
            if (MODE != 0) acc += __ldg(&w[row]) * s;
        }
    }
    if (MODE != 0) {
        acc = block_sum(acc);
        double v[1] = {acc};
        grid_sum_finish<1>(v, part, counter, dot, blockIdx.x, gridDim.x);
    }
}

template <int MODE>
static int32_t launch_spmv(pgd_ctx* h, const int32_t* rp, const int32_t* ci, const double* va, const double* x, double* y,
                           const double* w, int64_t n, double* dot, int lpr, cudaStream_t st) {
    if (h->opt_stream >= 2 && n >= PGD_BULK_MIN_ROWS && (((uintptr_t)ci | (uintptr_t)va) & 15) == 0) {
        if (!(h->attr_mask & (1u << MODE))) {  // the opt-in is per device, i.e. per handle (not per process)
            PGD_CUDA(h, cudaFuncSetAttribute(k_spmv_bulk<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM_BYTES));
            h->attr_mask |= (1u << MODE);
        }
        k_spmv_bulk<MODE><<<BK_CTAS_PER_SM * h->sm_count, SPMV_BULK_THREADS, BK_SMEM_BYTES, st>>>(rp, ci, va, x, y, w, n, dot, h->partials,
                                                                              h->counters);
        PGD_LAUNCH_OK(h);
        return 0;
    }
    if (h->opt_stream && n >= PGD_STREAM_MIN_ROWS) {
        const int64_t nblk = (n + ST_ROWS - 1) / ST_ROWS;
        unsigned int blocks = (unsigned int)min(nblk, (int64_t)h->sm_count * 6);
        k_spmv_stream<MODE><<<blocks, ST_THREADS, 0, st>>>(rp, ci, va, x, y, w, n, dot, h->partials, h->counters);
        PGD_LAUNCH_OK(h);
        return 0;
    }
    if (lpr == 0) {
        // mean row length from the two ends of rowptr would need a D2H copy; use 16 (P1 tets ~15/row)
        lpr = 16;
    }
    int rpb = 256 / lpr;
    unsigned int blocks = pgd_blocks(n, rpb);
    unsigned int cap = (unsigned int)h->sm_count * 8 * ((MODE == 0) ? 64 : 1);
    if (MODE != 0 && cap > 4096) cap = 4096;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    double* part = h->partials;
    unsigned int* ctr = h->counters;
#define SPMV_CASE(L)                                                                                   \
    case L:                                                                                            \
        k_spmv<L, MODE><<<blocks, 256, 0, st>>>(rp, ci, va, x, y, w, n, dot, part, ctr);               \
        break;
    switch (lpr) {
        SPMV_CASE(2)
        SPMV_CASE(4)
        SPMV_CASE(8)
        SPMV_CASE(16)
        SPMV_CASE(32)
        default:
            snprintf(h->err, sizeof(h->err), "lanes_per_row must be 0,2,4,8,16,32");
            return -2;
    }
#undef SPMV_CASE
    PGD_LAUNCH_OK(h);
    return 0;
}

int32_t pgd_spmv_internal(pgd_ctx* h, const int32_t* rp, const int32_t* ci, const double* va, const double* x, double* y,
                          int64_t n, int lpr, cudaStream_t st) {
    return launch_spmv<0>(h, rp, ci, va, x, y, nullptr, n, nullptr, lpr, st);
}

extern "C" int32_t pgd_spmv(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                            const double* d_x, double* d_y, int64_t n_rows, int32_t lanes_per_row, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_rowptr && d_colidx && d_values && d_x && d_y && n_rows >= 0, "bad arguments");
    if (n_rows == 0) return 0;
    return launch_spmv<0>(h, d_rowptr, d_colidx, d_values, d_x, d_y, nullptr, n_rows, nullptr, lanes_per_row,
                          (cudaStream_t)stream);
}

extern "C" int32_t pgd_spmv_dot(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                                const double* d_x, double* d_y, const double* d_w, double* d_dot, int64_t n_rows,
                                int32_t lanes_per_row, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_rowptr && d_colidx && d_values && d_x && d_y && d_w && d_dot && n_rows > 0, "bad arguments");
    return launch_spmv<1>(h, d_rowptr, d_colidx, d_values, d_x, d_y, d_w, n_rows, d_dot, lanes_per_row,
                          (cudaStream_t)stream);
}

extern "C" int32_t pgd_bilinear(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                                const double* d_x, const double* d_y, int64_t n_rows, double* d_out, int32_t lanes_per_row,
                                void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_rowptr && d_colidx && d_values && d_x && d_y && d_out && n_rows > 0, "bad arguments");
    // x^T (A y): gather y, weight rows by x
    return launch_spmv<2>(h, d_rowptr, d_colidx, d_values, d_y, nullptr, d_x, n_rows, d_out, lanes_per_row,
                          (cudaStream_t)stream);
}

// ----------------------------------------------------------------------------- Dirichlet
// one warp per bc dof j: row j -> unit row; every (c, j) with c in pattern(row j) -> 0.  The right-hand-side lifting of
// non-zero values is NOT done here (it used to be an atomicAdd per touched row, i.e. order-dependent round-off): the
// entry point below forms it beforehand as one deterministic SpMV with the un-eliminated operator.
__global__ void __launch_bounds__(256) k_dirichlet(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                   double* __restrict__ vals, const int32_t* __restrict__ bc, int64_t n_bc) {
    int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (wid >= n_bc) return;
    int j = bc[wid];
    int k0 = rowptr[j], k1 = rowptr[j + 1];
    for (int k = k0 + lane; k < k1; k += 32) {
        int c = colidx[k];
        if (c == j) {
            vals[k] = 1.0;
        } else {
            vals[k] = 0.0;
            // find (c, j) in row c (columns ascending)
            int lo = rowptr[c], hi = rowptr[c + 1] - 1;
            while (lo <= hi) {
                int mid = (lo + hi) >> 1;
                int cc = colidx[mid];
                if (cc == j) {
                    vals[mid] = 0.0;
                    break;
                }
                if (cc < j) lo = mid + 1;
                else hi = mid - 1;
            }
        }
    }
}

__global__ void __launch_bounds__(256) k_axpy_neg(double* __restrict__ b, const double* __restrict__ y, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) b[i] -= y[i];
}

extern "C" int32_t pgd_apply_dirichlet(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, double* d_values,
                                       double* d_b, const int32_t* d_bc_dofs, const double* d_bc_vals, int64_t n_bc,
                                       int64_t n_rows, double* d_work, void* stream) {
    PGD_CHECK_HANDLE(h);
    if (n_bc <= 0) return 0;
    PGD_ARG(h, d_rowptr && d_colidx && d_bc_dofs, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (d_values && d_b && d_bc_vals) {
        // lifting b -= A g with g = the prescribed values on the bc dofs, 0 elsewhere: one SpMV with the operator as
        // it is BEFORE the elimination (fixed summation order per row => bitwise reproducible)
        PGD_ARG(h, n_rows > 0 && d_work, "non-zero Dirichlet values need n_rows and 2 * n_rows doubles of work space");
        double* g = d_work;
        double* y = d_work + n_rows;
        PGD_CUDA(h, cudaMemsetAsync(g, 0, sizeof(double) * n_rows, st));
        int32_t rc = pgd_set_entries(h, g, d_bc_dofs, d_bc_vals, n_bc, stream);
        if (rc) return rc;
        rc = pgd_spmv_internal(h, d_rowptr, d_colidx, d_values, g, y, n_rows, 0, st);
        if (rc) return rc;
        k_axpy_neg<<<pgd_blocks(n_rows, 256), 256, 0, st>>>(d_b, y, n_rows);
        PGD_LAUNCH_OK(h);
    }
    if (d_values) {
        k_dirichlet<<<pgd_blocks(n_bc * 32, 256), 256, 0, st>>>(d_rowptr, d_colidx, d_values, d_bc_dofs, n_bc);
        PGD_LAUNCH_OK(h);
    }
    if (d_b) return pgd_set_entries(h, d_b, d_bc_dofs, d_bc_vals, n_bc, stream);
    return 0;
}
