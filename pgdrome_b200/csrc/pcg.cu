// Jacobi / node-block-Jacobi preconditioned CG with all scalars and the convergence flag on the device
// (the host polls once per `check_every` iterations).  Three regimes, chosen per solve:
//   * the system fits on chip            -> pcg_resident.cu: ONE cooperative kernel per solve
//   * n >= PGD_BULK_MIN_ROWS (HBM-bound) -> 3 launches per iteration: k_spcg_direction (p = z + beta p),
//                                           k_spmv_bulk<1> (q = A p with p.q, TMA ring, 512-thread CTAs), k_pcg_update
//   * in between                         -> 2 launches: k_pcg_spmv[_stream|_bulk] (p_new = z + beta p_old fused into
//                                           the gather, q = A p_new, p.q) and k_pcg_update
//   k_pcg_update: alpha = rz/pq; x += alpha p; r -= alpha q; z = M^-1 r; rz' = r.z, rr = r.r; the block that
//                 finishes the reduction rotates rz, bumps the iteration counter and sets DONE (every other
//                 block has consumed the old scalars before it deposited its partials: no ordering hazard).
// Reductions are fixed-order two-stage sums (bitwise reproducible).  The sharded (multi-GPU) building blocks
// pgd_spcg_* at the end of this file reuse the same kernels with the reductions completed across ranks.
#include "common.cuh"
#include "spmv_bulk.cuh"
#include "spmv_stream.cuh"

#include "pcg_blocks.cuh"

// thread per node: Minv (BSxBS per node), r = b, z = Minv r, x = 0, p0 = p1 = 0; rz, bb
template <int BS>
__global__ void __launch_bounds__(256) k_pcg_init(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                  const double* __restrict__ vals, const double* __restrict__ b,
                                                  double* __restrict__ x, double* r, double* z, double* p0, double* p1,
                                                  double* minv, int64_t n_nodes, double rtol, double atol, double* sc,
                                                  int* fl, double* part, unsigned int* counter,
                                                  const double* __restrict__ ax0 = nullptr) {
    // ax0 != NULL: warm start, x holds the initial guess and ax0 = A x0  (r = b - A x0; bb stays ||b||^2)
    double rz = 0.0, bb = 0.0, rr = 0.0;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t nd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; nd < n_nodes; nd += stride) {
        double B[BS][BS], I[BS][BS], rb[BS];
#pragma unroll
        for (int i = 0; i < BS; ++i)
#pragma unroll
            for (int k = 0; k < BS; ++k) B[i][k] = csr_entry(rowptr, colidx, vals, (int)(nd * BS + i), (int)(nd * BS + k));
        invert_block<BS>(B, I);
#pragma unroll
        for (int i = 0; i < BS; ++i) {
            rb[i] = b[nd * BS + i];
            bb += rb[i] * rb[i];
            if (ax0) rb[i] -= ax0[nd * BS + i];
#pragma unroll
            for (int k = 0; k < BS; ++k) minv[(nd * BS + i) * BS + k] = I[i][k];
        }
#pragma unroll
        for (int i = 0; i < BS; ++i) {
            double zi = 0.0;
#pragma unroll
            for (int k = 0; k < BS; ++k) zi += I[i][k] * rb[k];
            int64_t d = nd * BS + i;
            r[d] = rb[i];
            z[d] = zi;
            if (!ax0) x[d] = 0.0;
            p0[d] = 0.0;
            p1[d] = 0.0;
            rz += rb[i] * zi;
            rr += rb[i] * rb[i];
        }
    }
    rz = block_sum(rz);
    bb = block_sum(bb);
    rr = block_sum(rr);
    double v[3] = {rz, bb, rr};
    grid_sum_finish<3>(v, part, counter, sc + S_TMP, blockIdx.x, gridDim.x);  // k_pcg_init_fin publishes the scalars
}

// single-thread epilogue of init (keeps k_pcg_init simple and race-free)
__global__ void k_pcg_init_fin(double* sc, int* fl, double rtol, double atol) {
    double rz = sc[S_TMP], bb = sc[S_TMP + 1], rr = sc[S_TMP + 2];  // rr = ||b - A x0||^2 (= bb for x0 = 0)
    sc[S_RZ_OLD] = 1.0;
    sc[S_RZ_NEW] = rz;
    sc[S_PQ] = 1.0;
    sc[S_RR] = rr;
    sc[S_BB] = bb;
    double t2 = rtol * rtol * bb;
    if (atol * atol > t2) t2 = atol * atol;
    sc[S_TOL2] = t2;
    fl[F_ITER] = 0;
    fl[F_BAD] = 0;
    fl[F_DONE] = (rr <= t2 || bb == 0.0) ? 1 : 0;
}

template <int LPR>
__global__ void __launch_bounds__(256) k_pcg_spmv(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                  const double* __restrict__ vals, const double* __restrict__ z,
                                                  double* pa, double* pb, double* __restrict__ q, int64_t n,
                                                  const double* sc, const int* fl, double* sc_out, double* part,
                                                  unsigned int* counter) {
    if (fl[F_DONE]) return;
    const int it = fl[F_ITER];
    const double beta = (it == 0) ? 0.0 : sc[S_RZ_NEW] / sc[S_RZ_OLD];
    const double* __restrict__ p_old = (it & 1) ? pb : pa;
    double* __restrict__ p_new = (it & 1) ? pa : pb;
    constexpr int RPB = 256 / LPR;
    const int lane = threadIdx.x % LPR;
    const int sub = threadIdx.x / LPR;
    double acc = 0.0;
    for (int64_t base = (int64_t)blockIdx.x * RPB; base < n; base += (int64_t)gridDim.x * RPB) {
        int64_t row = base + sub;
        int k0 = 0, k1 = 0;
        if (row < n) {
            k0 = __ldg(&rowptr[row]);
            k1 = __ldg(&rowptr[row + 1]);
        }
        double s = 0.0;
        for (int k = k0 + lane; k < k1; k += LPR) {
            int c = ld_stream(&colidx[k]);
            double pj = fma(beta, p_old[c], z[c]);
            s = fma(ld_stream(&vals[k]), pj, s);
        }
#pragma unroll
        for (int o = LPR >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, LPR);
        if (row < n && lane == 0) {
            double pi = fma(beta, p_old[row], z[row]);
            p_new[row] = pi;
            q[row] = s;
            acc = fma(pi, s, acc);
        }
    }
    acc = block_sum(acc);
    double v[1] = {acc};
    grid_sum_finish<1>(v, part, counter, sc_out + S_PQ, blockIdx.x, gridDim.x);
}

template <int BS>
__global__ void __launch_bounds__(256) k_pcg_update(double* __restrict__ x, double* __restrict__ r, double* __restrict__ z,
                                                    const double* pa, const double* pb, const double* __restrict__ q,
                                                    const double* __restrict__ minv, int64_t n_nodes, double* sc, int* fl,
                                                    double* part, unsigned int* counter) {
    if (fl[F_DONE]) return;
    const int it = fl[F_ITER];
    const double* __restrict__ p = (it & 1) ? pa : pb;  // the p_new written by k_pcg_spmv
    const double rz_cur = sc[S_RZ_NEW];
    const double alpha = rz_cur / sc[S_PQ];
    double rz = 0.0, rr = 0.0;
    pcg_update_rows<BS>(x, r, z, p, q, minv, n_nodes, alpha, rz, rr);
    rz = block_sum(rz);
    rr = block_sum(rr);
    double v[2] = {rz, rr};
    const bool last = grid_sum_finish<2>(v, part, counter, sc + S_TMP, blockIdx.x, gridDim.x);
    if (last && threadIdx.x == 0) {  // scalar rotation
        const double rz_next = sc[S_TMP], rr_new = sc[S_TMP + 1];
        sc[S_RZ_OLD] = rz_cur;
        sc[S_RZ_NEW] = rz_next;
        sc[S_RR] = rr_new;
        fl[F_ITER] = it + 1;
        if (!(rr_new > sc[S_TOL2])) fl[F_DONE] = 1;  // also stops on NaN
        if (!(rr_new == rr_new) || !(rz_next == rz_next)) fl[F_BAD] = 1;
    }
}

// TMA-pipelined variant of k_pcg_spmv (spmv_bulk.cuh)
__global__ void __launch_bounds__(BK_THREADS, BK_CTAS_PER_SM) k_pcg_spmv_bulk(const int32_t* __restrict__ rowptr,
                                                                 const int32_t* __restrict__ colidx,
                                                                 const double* __restrict__ vals,
                                                                 const double* __restrict__ z, double* pa, double* pb,
                                                                 double* __restrict__ q, int64_t n, const double* sc,
                                                                 const int* fl, double* sc_out, double* part,
                                                                 unsigned int* counter) {
    extern __shared__ __align__(128) unsigned char bk_smem[];
    if (fl[F_DONE]) return;
    const int it = fl[F_ITER];
    const double beta = (it == 0) ? 0.0 : sc[S_RZ_NEW] / sc[S_RZ_OLD];
    const double* __restrict__ p_old = (it & 1) ? pb : pa;
    double* __restrict__ p_new = (it & 1) ? pa : pb;
    double acc = 0.0;
    bk_spmv_rows(rowptr, colidx, vals, n, BkGatherZP{z, p_old, beta},
                 [&](int64_t row, double s) {
                     const double pi = fma(beta, p_old[row], z[row]);
                     p_new[row] = pi;
                     q[row] = s;
                     acc = fma(pi, s, acc);
                 },
                 bk_smem);
    acc = block_sum(acc);
    double v[1] = {acc};
    grid_sum_finish<1>(v, part, counter, sc_out + S_PQ, blockIdx.x, gridDim.x);
}

// streaming variant of k_pcg_spmv for systems that do not fit on chip (spmv_stream.cuh)
__global__ void __launch_bounds__(ST_THREADS) k_pcg_spmv_stream(const int32_t* __restrict__ rowptr,
                                                                 const int32_t* __restrict__ colidx,
                                                                 const double* __restrict__ vals,
                                                                 const double* __restrict__ z, double* pa, double* pb,
                                                                 double* __restrict__ q, int64_t n, const double* sc,
                                                                 const int* fl, double* sc_out, double* part,
                                                                 unsigned int* counter) {
    __shared__ double s_prod[ST_TILE];
    __shared__ int s_rp[ST_ROWS + 1];
    if (fl[F_DONE]) return;
    const int it = fl[F_ITER];
    const double beta = (it == 0) ? 0.0 : sc[S_RZ_NEW] / sc[S_RZ_OLD];
    const double* __restrict__ p_old = (it & 1) ? pb : pa;
    double* __restrict__ p_new = (it & 1) ? pa : pb;
    const GatherZP g{z, p_old, beta};
    const int64_t nblk = (n + ST_ROWS - 1) / ST_ROWS;
    double acc = 0.0;
    for (int64_t rb = blockIdx.x; rb < nblk; rb += gridDim.x) {
        const int64_t r0 = rb * ST_ROWS;
        const int nr = (int)min((int64_t)ST_ROWS, n - r0);
        const double s = stream_rowblock(rowptr, colidx, vals, g, r0, nr, s_prod, s_rp);
        if ((int)threadIdx.x < nr) {
            const int64_t row = r0 + threadIdx.x;
            const double pi = fma(beta, p_old[row], z[row]);
            p_new[row] = pi;
            q[row] = s;
            acc = fma(pi, s, acc);
        }
    }
    acc = block_sum(acc);
    double v[1] = {acc};
    grid_sum_finish<1>(v, part, counter, sc_out + S_PQ, blockIdx.x, gridDim.x);
}

int32_t pgd_spmv_internal(pgd_ctx* h, const int32_t* rp, const int32_t* ci, const double* va, const double* x, double* y,
                          int64_t n, int lpr, cudaStream_t st);
__global__ void __launch_bounds__(256) k_spcg_direction(const double* __restrict__ z, double* __restrict__ p, int64_t n, const double* sc,
                                 const int* fl);

template <int BS>
static int32_t run_pcg(pgd_ctx* h, const int32_t* rp, const int32_t* ci, const double* va, const double* b, double* x,
                       int64_t n, double rtol, double atol, int maxit, int check_every, int lpr, double* work,
                       int32_t* h_iters, double* h_relres, cudaStream_t st, bool warm) {
    const int64_t n_nodes = n / BS;
    const int64_t ns = (n + 1) & ~(int64_t)1;  // even stride: the work arrays stay 16-byte aligned (128-bit vector path)
    double* r = work;
    double* z = r + ns;
    double* p0 = z + ns;
    double* p1 = p0 + ns;
    double* q = p1 + ns;
    double* minv = q + ns;  // n*BS
    double* sc = h->scalars;
    int* fl = h->flags;
    unsigned int vb = pgd_blocks(n_nodes, 256);
    unsigned int capv = (unsigned int)h->sm_count * 8;
    if (vb > capv) vb = capv;
    if (lpr == 0) lpr = 16;
    unsigned int sb = pgd_blocks(n, 256 / lpr);
    unsigned int caps = (unsigned int)h->sm_count * 8;
    if (sb > caps) sb = caps;
    if (warm) {  // q = A x0 (q is free until the first iteration)
        int32_t rc = pgd_spmv_internal(h, rp, ci, va, x, q, n, lpr, st);
        if (rc) return rc;
    }
    k_pcg_init<BS><<<vb, 256, 0, st>>>(rp, ci, va, b, x, r, z, p0, p1, minv, n_nodes, rtol, atol, sc, fl, h->partials,
                                        h->counters, warm ? q : nullptr);
    PGD_LAUNCH_OK(h);
    k_pcg_init_fin<<<1, 1, 0, st>>>(sc, fl, rtol, atol);
    PGD_LAUNCH_OK(h);
    int hf[4] = {0, 0, 0, 0};
    double hs[8];
    int launched = 0;
    if (check_every < 1) check_every = 1;
    PGD_CUDA(h, cudaEventRecord(h->ev0, st));
    while (true) {
        PGD_CUDA(h, pgd_fetch(h, hf, fl, sizeof(int) * 4, nullptr, nullptr, 0, st));
        if (hf[F_DONE] || launched >= maxit) break;
        int todo = maxit - launched;
        if (todo > check_every) todo = check_every;
        const bool stream = h->opt_stream && n >= PGD_STREAM_MIN_ROWS;
        const unsigned int stb = (unsigned int)min((n + ST_ROWS - 1) / ST_ROWS, (int64_t)h->sm_count * 6);
        const bool bulk = h->opt_stream >= 2 && n >= PGD_BULK_MIN_ROWS && (((uintptr_t)ci | (uintptr_t)va) & 15) == 0;
        if (bulk)
            PGD_CUDA(h, cudaFuncSetAttribute(k_pcg_spmv_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM_BYTES));
        for (int i = 0; i < todo; ++i) {
            if (bulk && h->opt_pcg3) {
                // 3-kernel iteration for the HBM-bound regime: p = z + beta p materialised (vectors are L2-resident),
                // then the single-gather 512-thread SpMV+dot, then the update.  138 -> 12x us per iteration on the
                // 128^3 mesh against the 2-kernel form whose SpMV gathers z and p (90 registers, 256-thread CTAs).
                unsigned int db = pgd_blocks(n, 512);
                if (db > (unsigned int)h->sm_count * 16) db = (unsigned int)h->sm_count * 16;
                k_spcg_direction<<<db, 256, 0, st>>>(z, p0, n, sc, fl);
                int32_t rc = pgd_spmv_dot(h, rp, ci, va, p0, q, p0, sc + S_PQ, n, 0, (void*)st);
                if (rc) return rc;
                k_pcg_update<BS><<<vb, 256, 0, st>>>(x, r, z, p0, p0, q, minv, n_nodes, sc, fl, h->partials, h->counters);
                h->n_launches += 1;
                continue;
            }
            if (bulk) {
                k_pcg_spmv_bulk<<<BK_CTAS_PER_SM * h->sm_count, BK_THREADS, BK_SMEM_BYTES, st>>>(rp, ci, va, z, p0, p1, q, n, sc, fl, sc,
                                                                                    h->partials, h->counters);
            } else if (stream) {
                k_pcg_spmv_stream<<<stb, ST_THREADS, 0, st>>>(rp, ci, va, z, p0, p1, q, n, sc, fl, sc, h->partials,
                                                               h->counters);
            } else
            switch (lpr) {
#define PCG_CASE(L)                                                                                                     \
    case L:                                                                                                             \
        k_pcg_spmv<L><<<sb, 256, 0, st>>>(rp, ci, va, z, p0, p1, q, n, sc, fl, sc, h->partials, h->counters);           \
        break;
                PCG_CASE(2)
                PCG_CASE(4)
                PCG_CASE(8)
                PCG_CASE(16)
                PCG_CASE(32)
#undef PCG_CASE
                default:
                    snprintf(h->err, sizeof(h->err), "lanes_per_row must be 0,2,4,8,16,32");
                    return -2;
            }
            k_pcg_update<BS><<<vb, 256, 0, st>>>(x, r, z, p0, p1, q, minv, n_nodes, sc, fl, h->partials, h->counters);
        }
        PGD_LAUNCH_OK(h);
        h->n_launches += 2 * (int64_t)todo - 1;  // (the 3-kernel path counted its third launch above)
        launched += todo;
    }
    PGD_CUDA(h, cudaEventRecord(h->ev1, st));
    PGD_CUDA(h, pgd_fetch(h, hs, sc, sizeof(double) * 8, nullptr, nullptr, 0, st));
    {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->pcg_ms += ms;
        h->pcg_solves += 1;
        h->pcg_iters += hf[F_ITER];
    }
    if (h_iters) *h_iters = hf[F_ITER];
    if (h_relres) *h_relres = (hs[S_BB] > 0.0) ? sqrt(hs[S_RR] / hs[S_BB]) : 0.0;
    if (!(hs[S_BB] > 0.0) && warm)  // b = 0: the solution is 0 whatever the initial guess was (as in the SM-resident solver)
        PGD_CUDA(h, cudaMemsetAsync(x, 0, sizeof(double) * n, st));
    if (hf[F_BAD]) {
        snprintf(h->err, sizeof(h->err), "pgd_pcg_sync: NaN encountered (matrix not SPD?)");
        return -3;
    }
    return 0;
}

int32_t pgd_pcg_resident(pgd_ctx* h, const int32_t* rp, const int32_t* ci, const double* va, const double* b, double* x,
                         int64_t n, int64_t nnz_hint, double rtol, double atol, int maxit, int block, double* work,
                         int32_t* h_iters, double* h_relres, cudaStream_t st, int warm, int defer);
int32_t pgd_pcg_finish_impl(pgd_ctx* h, int32_t* h_iters, double* h_relres);

static int32_t pcg_entry(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                         const double* d_b, double* d_x, int64_t n, double rtol, double atol, int32_t maxit,
                         int32_t check_every, int32_t block, int32_t lanes_per_row, double* d_work, int32_t* h_iters,
                         double* h_relres, void* stream, bool warm) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_rowptr && d_colidx && d_values && d_b && d_x && d_work && n > 0, "bad arguments");
    PGD_ARG(h, block >= 1 && block <= 3 && n % block == 0, "block must be 1..3 and divide n");
    cudaStream_t st = (cudaStream_t)stream;
    {  // a solve started earlier and never collected: finish it first (its counters are kept, its status is dropped)
        int32_t rc0 = pgd_pcg_finish_impl(h, nullptr, nullptr);
        if (rc0 < 0) return rc0;
    }
    if (h->opt_resident && n <= ((int64_t)1 << 22)) {
        if (h->nnz_key != (const void*)d_rowptr || h->nnz_key_n != n) {
            int32_t last = 0;
            PGD_CUDA(h, pgd_fetch(h, &last, d_rowptr + n, sizeof(int32_t), nullptr, nullptr, 0, st));
            h->nnz_key = (const void*)d_rowptr;
            h->nnz_key_n = n;
            h->nnz_val = last;
        }
        int32_t rc = pgd_pcg_resident(h, d_rowptr, d_colidx, d_values, d_b, d_x, n, h->nnz_val, rtol, atol, maxit, block,
                                      d_work, h_iters, h_relres, st, warm ? 1 : 0, 0);
        if (rc != 1) return rc;
    }
    if (h->opt_persist >= 2 && n >= PGD_BULK_MIN_ROWS && (((uintptr_t)d_colidx | (uintptr_t)d_values | (uintptr_t)d_work) & 15) == 0)
        // pgd_set_option("persist", 2): the whole solve in one persistent cooperative kernel (pcg_persist.cu) also for
        // plain CSR on one GPU.  Not the default: measured 0.235 ms against 0.215 ms per iteration for the three-launch
        // form on the 4 M-dof scalar operator (the persistent kernel must gather p with coherent loads and pays three
        // grid-wide waits; it wins with the node-block walk and on sharded systems, where the caller selects it).
        // d_work of (5 + block) n + 8 doubles covers its layout for n_local = n
        return pgd_pcg_persist_sync(h, d_rowptr, d_colidx, d_values, d_b, d_x, n, n, block, rtol, atol, maxit, warm ? 1 : 0,
                                    d_work, nullptr, nullptr, nullptr, nullptr, nullptr, 0, h_iters, h_relres, stream);
    if (block == 1)
        return run_pcg<1>(h, d_rowptr, d_colidx, d_values, d_b, d_x, n, rtol, atol, maxit, check_every, lanes_per_row, d_work,
                          h_iters, h_relres, st, warm);
    if (block == 2)
        return run_pcg<2>(h, d_rowptr, d_colidx, d_values, d_b, d_x, n, rtol, atol, maxit, check_every, lanes_per_row, d_work,
                          h_iters, h_relres, st, warm);
    return run_pcg<3>(h, d_rowptr, d_colidx, d_values, d_b, d_x, n, rtol, atol, maxit, check_every, lanes_per_row, d_work,
                      h_iters, h_relres, st, warm);
}

extern "C" int32_t pgd_pcg_sync(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                                const double* d_b, double* d_x, int64_t n, double rtol, double atol, int32_t maxit,
                                int32_t check_every, int32_t block, int32_t lanes_per_row, double* d_work,
                                int32_t* h_iters, double* h_relres, void* stream) {
    return pcg_entry(h, d_rowptr, d_colidx, d_values, d_b, d_x, n, rtol, atol, maxit, check_every, block, lanes_per_row, d_work,
                     h_iters, h_relres, stream, false);
}

/* warm start: d_x holds the initial guess x0 on entry (same stopping rule, relative to ||b||) */
extern "C" int32_t pgd_pcg_x0_sync(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                                   const double* d_b, double* d_x, int64_t n, double rtol, double atol, int32_t maxit,
                                   int32_t check_every, int32_t block, int32_t lanes_per_row, double* d_work,
                                   int32_t* h_iters, double* h_relres, void* stream) {
    return pcg_entry(h, d_rowptr, d_colidx, d_values, d_b, d_x, n, rtol, atol, maxit, check_every, block, lanes_per_row, d_work,
                     h_iters, h_relres, stream, true);
}

// ================================================================================================
// Sharded PCG building blocks (spatial mesh partitioned by rows over the GPUs of one box).
//
// Each rank owns a contiguous block of rows; its local vectors are laid out [owned | ghost] and the
// local CSR uses that numbering.  One iteration on every rank is
//     pgd_spcg_direction   p = z + beta p                         (owned entries; beta from device scalars)
//     <halo exchange of p  -- NCCL, issued by the host on the same stream>
//     pgd_spcg_matvec      q = A_loc p,  local p.q   -> sc[S_PQ]  (the TMA-pipelined SpMV+dot kernel)
//     <allreduce sc[S_PQ]>
//     pgd_spcg_update      x += a p; r -= a q; z = M^-1 r; local r.z, r.r -> sc[S_TMP], sc[S_TMP+1]
//     <allreduce sc[S_TMP..S_TMP+1]>
//     pgd_spcg_rotate      scalar rotation, iteration counter, DONE flag (identical on every rank)
// All calls are asynchronous launches without host synchronisation or allocation, so the host can
// capture an iteration (kernels + NCCL collectives) in a CUDA graph.  d_sc: >= 16 doubles, d_fl: >= 4
// ints, both caller-owned (the collectives operate on slices of d_sc).
// work layout: r[ns] z[ns] q[ns] minv[even(no*block)] p[nl]   (no = owned rows, ns = no rounded up to even,
// nl = owned + ghost)
// ================================================================================================
__global__ void __launch_bounds__(256) k_spcg_direction(const double* __restrict__ z, double* __restrict__ p, int64_t n,
                                                        const double* sc, const int* fl) {
    if (fl[F_DONE]) return;
    const double beta = (fl[F_ITER] == 0) ? 0.0 : sc[S_RZ_NEW] / sc[S_RZ_OLD];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(p)) & 15) == 0) {
        const int64_t n2 = n >> 1;
        const double2* z2 = reinterpret_cast<const double2*>(z);
        double2* p2 = reinterpret_cast<double2*>(p);
#pragma unroll 2
        for (int64_t i = gtid; i < n2; i += stride) {
            const double2 zv = z2[i];
            double2 pv = p2[i];
            pv.x = fma(beta, pv.x, zv.x);
            pv.y = fma(beta, pv.y, zv.y);
            p2[i] = pv;
        }
        if ((n & 1) && gtid == 0) p[n - 1] = fma(beta, p[n - 1], z[n - 1]);
        return;
    }
    for (int64_t i = gtid; i < n; i += stride) p[i] = fma(beta, p[i], z[i]);
}

// x += alpha p ; r -= alpha q ; z = M^-1 r ; local sums (no rotation: the host all-reduces first)
template <int BS>
__global__ void __launch_bounds__(256) k_spcg_update(double* __restrict__ x, double* __restrict__ r, double* __restrict__ z,
                                                     const double* __restrict__ p, const double* __restrict__ q,
                                                     const double* __restrict__ minv, int64_t n_nodes, double* sc,
                                                     const int* fl, double* part, unsigned int* counter) {
    if (fl[F_DONE]) return;
    const double alpha = sc[S_RZ_NEW] / sc[S_PQ];
    double rz = 0.0, rr = 0.0;
    pcg_update_rows<BS>(x, r, z, p, q, minv, n_nodes, alpha, rz, rr);
    rz = block_sum(rz);
    rr = block_sum(rr);
    double v[2] = {rz, rr};
    grid_sum_finish<2>(v, part, counter, sc + S_TMP, blockIdx.x, gridDim.x);
}

__global__ void k_spcg_rotate(double* sc, int* fl) {
    if (fl[F_DONE]) return;
    const double rz_next = sc[S_TMP], rr = sc[S_TMP + 1];
    sc[S_RZ_OLD] = sc[S_RZ_NEW];
    sc[S_RZ_NEW] = rz_next;
    sc[S_RR] = rr;
    fl[F_ITER] = fl[F_ITER] + 1;
    if (!(rr > sc[S_TOL2])) fl[F_DONE] = 1;
    if (!(rr == rr) || !(rz_next == rz_next)) fl[F_BAD] = 1;
}

struct SpcgWork {
    double *r, *z, *q, *minv, *p;
};
static SpcgWork spcg_work(pgd_ctx* h, double* work, int64_t no, int block) {
    SpcgWork w;
    const int64_t ns = (no + 1) & ~(int64_t)1;  // even strides keep every array 16-byte aligned (128-bit vector paths)
    w.r = work;
    w.z = w.r + ns;
    w.q = w.z + ns;
    w.minv = w.q + ns;
    w.p = h->p_override ? h->p_override : w.minv + ((no * block + 1) & ~(int64_t)1);
    return w;
}

extern "C" int32_t pgd_spcg_init(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                                 const double* d_b, double* d_x, int64_t n_owned, int64_t n_local, int32_t block,
                                 double* d_work, double* d_sc, int32_t* d_fl, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_rowptr && d_colidx && d_values && d_b && d_x && d_work && d_sc && d_fl, "null pointer");
    PGD_ARG(h, n_owned >= 0 && n_local >= n_owned && block >= 1 && block <= 3 && n_owned % block == 0, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    SpcgWork w = spcg_work(h, d_work, n_owned, block);
    PGD_CUDA(h, cudaMemsetAsync(w.p, 0, sizeof(double) * n_local, st));
    PGD_CUDA(h, cudaMemsetAsync(d_sc, 0, sizeof(double) * 16, st));
    PGD_CUDA(h, cudaMemsetAsync(d_fl, 0, sizeof(int32_t) * 4, st));
    const int64_t n_nodes = n_owned / block;
    if (n_nodes == 0) return 0;
    unsigned int vb = pgd_blocks(n_nodes, 256);
    unsigned int capv = (unsigned int)h->sm_count * 8;
    if (vb > capv) vb = capv;
    // k_pcg_init writes p0/p1 = 0 for owned entries (here both alias w.p) and the local (rz, bb) to sc[S_TMP..]
    if (block == 1) k_pcg_init<1><<<vb, 256, 0, st>>>(d_rowptr, d_colidx, d_values, d_b, d_x, w.r, w.z, w.p, w.p, w.minv, n_nodes, 0.0, 0.0, d_sc, d_fl, h->partials, h->counters);
    else if (block == 2) k_pcg_init<2><<<vb, 256, 0, st>>>(d_rowptr, d_colidx, d_values, d_b, d_x, w.r, w.z, w.p, w.p, w.minv, n_nodes, 0.0, 0.0, d_sc, d_fl, h->partials, h->counters);
    else k_pcg_init<3><<<vb, 256, 0, st>>>(d_rowptr, d_colidx, d_values, d_b, d_x, w.r, w.z, w.p, w.p, w.minv, n_nodes, 0.0, 0.0, d_sc, d_fl, h->partials, h->counters);
    PGD_LAUNCH_OK(h);
    return 0;
}

/* after the all-reduce of d_sc[S_TMP..S_TMP+1] = (r.z, b.b) */
extern "C" int32_t pgd_spcg_init_fin(pgd_handle_t h, double* d_sc, int32_t* d_fl, double rtol, double atol, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_sc && d_fl, "null pointer");
    k_pcg_init_fin<<<1, 1, 0, (cudaStream_t)stream>>>(d_sc, d_fl, rtol, atol);
    PGD_LAUNCH_OK(h);
    return 0;
}

extern "C" int32_t pgd_spcg_direction(pgd_handle_t h, double* d_work, int64_t n_owned, int32_t block, const double* d_sc,
                                      const int32_t* d_fl, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_work && d_sc && d_fl && n_owned >= 0, "bad arguments");
    if (n_owned == 0) return 0;
    SpcgWork w = spcg_work(h, d_work, n_owned, block);
    unsigned int vb = pgd_blocks(n_owned, 256);
    unsigned int capv = (unsigned int)h->sm_count * 16;
    if (vb > capv) vb = capv;
    k_spcg_direction<<<vb, 256, 0, (cudaStream_t)stream>>>(w.z, w.p, n_owned, d_sc, d_fl);
    PGD_LAUNCH_OK(h);
    return 0;
}

extern "C" int32_t pgd_spcg_matvec(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                                   double* d_work, int64_t n_owned, int32_t block, double* d_sc, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_rowptr && d_colidx && d_values && d_work && d_sc && n_owned >= 0, "bad arguments");
    SpcgWork w = spcg_work(h, d_work, n_owned, block);
    if (n_owned == 0) return (int32_t)cudaMemsetAsync(d_sc + S_PQ, 0, sizeof(double), (cudaStream_t)stream);
    // q = A_loc p (p includes the ghost entries), sc[S_PQ] = local p.q     [not skipped when DONE: harmless]
    return pgd_spmv_dot(h, d_rowptr, d_colidx, d_values, w.p, w.q, w.p, d_sc + S_PQ, n_owned, 0, stream);
}

extern "C" int32_t pgd_spcg_update(pgd_handle_t h, double* d_x, double* d_work, int64_t n_owned, int32_t block, double* d_sc,
                                   const int32_t* d_fl, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_x && d_work && d_sc && d_fl && n_owned >= 0 && block >= 1 && block <= 3, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    SpcgWork w = spcg_work(h, d_work, n_owned, block);
    const int64_t n_nodes = n_owned / block;
    if (n_nodes == 0) return (int32_t)cudaMemsetAsync(d_sc + S_TMP, 0, 2 * sizeof(double), st);
    unsigned int vb = pgd_blocks(n_nodes, 256);
    unsigned int capv = (unsigned int)h->sm_count * 8;
    if (vb > capv) vb = capv;
    if (block == 1) k_spcg_update<1><<<vb, 256, 0, st>>>(d_x, w.r, w.z, w.p, w.q, w.minv, n_nodes, d_sc, d_fl, h->partials, h->counters);
    else if (block == 2) k_spcg_update<2><<<vb, 256, 0, st>>>(d_x, w.r, w.z, w.p, w.q, w.minv, n_nodes, d_sc, d_fl, h->partials, h->counters);
    else k_spcg_update<3><<<vb, 256, 0, st>>>(d_x, w.r, w.z, w.p, w.q, w.minv, n_nodes, d_sc, d_fl, h->partials, h->counters);
    PGD_LAUNCH_OK(h);
    return 0;
}

extern "C" int32_t pgd_spcg_rotate(pgd_handle_t h, double* d_sc, int32_t* d_fl, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_sc && d_fl, "null pointer");
    k_spcg_rotate<<<1, 1, 0, (cudaStream_t)stream>>>(d_sc, d_fl);
    PGD_LAUNCH_OK(h);
    return 0;
}

/* Started / finished solve (see include/pgd_b200.h). */
extern "C" int32_t pgd_pcg_start(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                                 const double* d_b, double* d_x, int64_t n, double rtol, double atol, int32_t maxit,
                                 int32_t block, double* d_work, int32_t warm, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_rowptr && d_colidx && d_values && d_b && d_x && d_work && n > 0, "bad arguments");
    PGD_ARG(h, block >= 1 && block <= 3 && n % block == 0, "block must be 1..3 and divide n");
    if (!h->opt_resident || h->fit_key != (const void*)d_rowptr || h->fit_n != n || h->fit_block != block ||
        h->nnz_key != (const void*)d_rowptr || h->nnz_key_n != n)
        return 1;  // not known to fit the SM-resident solver: use pgd_pcg_sync / pgd_pcg_x0_sync
    return pgd_pcg_resident(h, d_rowptr, d_colidx, d_values, d_b, d_x, n, h->nnz_val, rtol, atol, maxit, block, d_work, nullptr,
                            nullptr, (cudaStream_t)stream, warm ? 1 : 0, 1);
}

extern "C" int32_t pgd_pcg_finish(pgd_handle_t h, int32_t* h_iters, double* h_relres) {
    PGD_CHECK_HANDLE(h);
    int32_t rc = pgd_pcg_finish_impl(h, h_iters, h_relres);
    if (rc == 1) {
        snprintf(h->err, sizeof(h->err), "pgd_pcg_finish: the started system no longer fits the SM-resident solver");
        return -5;
    }
    return rc;
}
