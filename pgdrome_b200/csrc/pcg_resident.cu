// SM-resident Jacobi / node-block-Jacobi PCG: ONE cooperative kernel per solve.
//
// For the systems of BASELINE configs[0]/[1] (and every 1-D/2-D dimension of the other configs) the
// CSR matrix is a few MB: per-iteration cost is launch latency and reduction tails, not bandwidth.
// Here the matrix never leaves the chip: CTA b (one per SM, 148 on B200) copies its contiguous,
// nnz-balanced row slice of (values, columns) into its shared memory ONCE (up to ~200 KB per SM, i.e.
// ~2.5 M nonzeros chip-wide) together with its slices of x, r, p, q, M^-1, and then iterates
//
//     S2: q_i = sum_j A_ij (z_j + beta p_j)      z, p_old over the slice's column window: cp.async.cg started right
//                                                  after the previous barrier, overlapping the reduction round trip
//         p_i = z_i + beta p_i ; pq = sum p_i q_i                         -> grid barrier (reduces pq)
//     S3: x += alpha p ; r -= alpha q ; z = M^-1 r ; rz, rr                -> grid barrier (reduces rz, rr)
//
// with two grid-wide barriers per iteration (monotonic arrival counter in L2 + per-CTA partials that
// every CTA sums in the same fixed order => bitwise identical scalars in all CTAs, deterministic, and
// a uniform convergence exit without any host round trip).  Launched with
// cudaLaunchCooperativeKernel so that co-residency of all CTAs is guaranteed by the driver.
#include <cooperative_groups.h>

#include "common.cuh"

// 16-byte asynchronous global -> shared copy that bypasses L1 (.cg): the window of p / z is written by other CTAs
__device__ __forceinline__ void cp_async_cg16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

#ifndef RES_THREADS
#define RES_THREADS 512
#endif

struct ResArgs {
    const int32_t* rowptr;
    const int32_t* colidx;
    const double* vals;
    const double* b;
    double* x;
    int64_t n;
    double rtol, atol;
    int maxit;
    double* zg;        // [n] z, globally visible
    double* pg0;       // [n] p ping
    double* pg1;       // [n] p pong
    double* part;      // [2][2][G] partial sums
    unsigned int* bar; // arrival counter (0 at launch)
    double* out_sc;    // [0]=rr [1]=bb
    int* out_fl;       // [0]=iters [1]=status (0 ok, 1 slice does not fit, 2 NaN)
    int cap_nnz, cap_rows, cap_win;
    int warm;          // 1: x holds the initial guess x0 (r = b - A x0)
};
// zg / pg0 / pg1 are spaced by an even number of doubles so that 16-byte copies of a window stay aligned

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// all threads call; returns after every CTA of the grid has arrived (and their prior global writes
// are visible).  `target` is this CTA's private running arrival target.
// LEAD = false: the caller has already executed a __syncthreads after the CTA's last global write (block_sum does).
template <bool LEAD = true>
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int& target, unsigned int G) {
    if (LEAD) __syncthreads();
    target += G;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        while (ld_acquire(bar) < target) {
        }
    }
    __syncthreads();
}

// Sum over the block of two values at once (same tree per value as block_sum); results valid in thread 0.
__device__ __forceinline__ void block_sum2(double& a, double& b) {
    __shared__ double sh2[2][32];
    __syncthreads();
    a = warp_sum(a);
    b = warp_sum(b);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) {
        sh2[0][w] = a;
        sh2[1][w] = b;
    }
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    a = (threadIdx.x < nw) ? sh2[0][threadIdx.x] : 0.0;
    b = (threadIdx.x < nw) ? sh2[1][threadIdx.x] : 0.0;
    if (w == 0) {
        a = warp_sum(a);
        b = warp_sum(b);
    }
}

// two values per CTA stored side by side (one 16-byte load per CTA fetches both): same per-value order as sum_partials
__device__ __forceinline__ void sum_partials_pair(const double2* part, unsigned int G, double* s_out) {
    if (threadIdx.x < 32) {
        double s0 = 0.0, s1 = 0.0;
        for (unsigned int i = threadIdx.x; i < G; i += 32) {
            const double2 v = __ldcg(&part[i]);
            s0 += v.x;
            s1 += v.y;
        }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        if (threadIdx.x == 0) {
            s_out[0] = s0;
            s_out[1] = s1;
        }
    }
    __syncthreads();
}

// fixed-order sum of NV x G partials by warp 0, broadcast through shared memory
template <int NV>
__device__ __forceinline__ void sum_partials(const double* part, unsigned int G, double* s_out) {
    if (threadIdx.x < 32) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            double s = 0.0;
            for (unsigned int i = threadIdx.x; i < G; i += 32) s += __ldcg(&part[(size_t)v * G + i]);
            s = warp_sum(s);
            if (threadIdx.x == 0) s_out[v] = s;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ int64_t slice_boundary(const int32_t* rowptr, int64_t n, int64_t nnz, unsigned int b,
                                                  unsigned int G, int bs) {
    if (b == 0) return 0;
    if (b >= G) return n;
    int64_t tgt = (int64_t)((double)nnz * (double)b / (double)G);
    int64_t lo = 0, hi = n;  // first row r with rowptr[r] >= tgt
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (rowptr[mid] >= tgt) hi = mid;
        else lo = mid + 1;
    }
    lo = (lo / bs) * bs;
    return lo;
}

template <int BS>
__global__ void __launch_bounds__(512, 1) k_pcg_resident(ResArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_red[4];
    __shared__ int s_bad;
    const unsigned int G = gridDim.x, bid = blockIdx.x;
    const int tid = threadIdx.x, nt = blockDim.x;
    unsigned int target = 0;

    const int64_t nnz_total = a.rowptr[a.n];
    const int64_t r0 = slice_boundary(a.rowptr, a.n, nnz_total, bid, G, BS);
    const int64_t r1 = slice_boundary(a.rowptr, a.n, nnz_total, bid + 1, G, BS);
    const int rows = (int)(r1 - r0);
    const int k0 = a.rowptr[r0];
    const int nnz = a.rowptr[r1] - k0;

    double* s_vals = reinterpret_cast<double*>(smem_raw);
    double* s_x = s_vals + a.cap_nnz;
    double* s_r = s_x + a.cap_rows;
    double* s_p = s_r + a.cap_rows;
    double* s_q = s_p + a.cap_rows;
    double* s_z = s_q + a.cap_rows;
    double* s_minv = s_z + a.cap_rows;  // [cap_rows * BS]
    double* s_zw = s_minv + (size_t)a.cap_rows * BS;  // [cap_win] window of z over the slice's column range
    double* s_pw = s_zw + a.cap_win;                   // [cap_win] window of p_old
    int* s_cols = reinterpret_cast<int*>(s_pw + a.cap_win);
    int* s_rp = s_cols + a.cap_nnz;  // [cap_rows + 1], relative to k0

    if (tid == 0) s_bad = (nnz > a.cap_nnz || rows > a.cap_rows) ? 1 : 0;
    __syncthreads();
    if (s_bad && tid == 0) a.out_fl[1] = 1;
    grid_barrier(a.bar, target, G);
    if (__ldcg(&a.out_fl[1]) != 0) return;  // uniform: some slice does not fit -> host falls back

    // ---- stage the slice
    for (int k = tid; k < nnz; k += nt) {
        s_vals[k] = __ldcs(&a.vals[k0 + k]);
        s_cols[k] = __ldcs(&a.colidx[k0 + k]);
    }
    for (int i = tid; i <= rows; i += nt) s_rp[i] = a.rowptr[r0 + i] - k0;
    __syncthreads();
    // column range of the slice: if it fits, every iteration first loads p_new = z + beta p_old over
    // [cmin, cmax] into shared memory (asynchronous 16-byte copies) and the SpMV then runs entirely out of shared memory
    int cmin = 0x7fffffff, cmax = -1;
    for (int k = tid; k < nnz; k += nt) {
        cmin = min(cmin, s_cols[k]);
        cmax = max(cmax, s_cols[k]);
    }
    {
        __shared__ int s_mm[2];
        if (tid == 0) {
            s_mm[0] = (int)r0;
            s_mm[1] = (int)r1 - 1;
        }
        __syncthreads();
        atomicMin(&s_mm[0], cmin);
        atomicMax(&s_mm[1], cmax);
        __syncthreads();
        cmin = s_mm[0];
        cmax = s_mm[1];
    }
    cmin &= ~1;                                   // even start and even length: 16-byte copies
    const int wlen = ((cmax + 2) & ~1) - cmin;
    const bool windowed = (rows > 0) && (wlen <= a.cap_win);
    // after a grid barrier: start copying the window of z and of the given p into shared memory (no registers, L1
    // bypassed); it is waited for at the top of the next S2, so the round trip overlaps the reduction of the partials
    auto prefetch_window = [&](const double* psrc) {
        if (!windowed) return;
        for (int j = 2 * tid; j < wlen; j += 2 * nt) {
            cp_async_cg16(&s_zw[j], &a.zg[cmin + j]);
            cp_async_cg16(&s_pw[j], &psrc[cmin + j]);
        }
    };
    if (windowed)
        for (int k = tid; k < nnz; k += nt) s_cols[k] -= cmin;
    __syncthreads();
    // ---- M^-1 (node blocks), r = b, z = M^-1 r, x = 0, p = 0
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
    for (int nd = tid; nd < rows / BS; nd += nt) {
        double B[BS][BS], I[BS][BS], rb[BS];
#pragma unroll
        for (int i = 0; i < BS; ++i) {
#pragma unroll
            for (int k = 0; k < BS; ++k) B[i][k] = 0.0;
            const int row = nd * BS + i;
            for (int k = s_rp[row]; k < s_rp[row + 1]; ++k) {
                int c = s_cols[k] + (windowed ? cmin : 0) - (int)(r0 + (int64_t)nd * BS);
                if (c >= 0 && c < BS) {
#pragma unroll
                    for (int kk = 0; kk < BS; ++kk)
                        if (c == kk) B[i][kk] = s_vals[k];
                }
            }
            rb[i] = a.b[r0 + row];
            if (a.warm) {  // r = b - A x0: x0 is a read-only input until the final store, no barrier needed
                double ax = 0.0;
                for (int k = s_rp[row]; k < s_rp[row + 1]; ++k)
                    ax = fma(s_vals[k], __ldcg(&a.x[s_cols[k] + (windowed ? cmin : 0)]), ax);
                acc1 += rb[i] * rb[i];
                rb[i] -= ax;
            }
        }
        if constexpr (BS == 1) {
            I[0][0] = 1.0 / B[0][0];
        } else if constexpr (BS == 2) {
            double id = 1.0 / (B[0][0] * B[1][1] - B[0][1] * B[1][0]);
            I[0][0] = B[1][1] * id;
            I[0][1] = -B[0][1] * id;
            I[1][0] = -B[1][0] * id;
            I[1][1] = B[0][0] * id;
        } else {
            double c00 = B[1][1] * B[2][2] - B[1][2] * B[2][1];
            double c01 = B[1][2] * B[2][0] - B[1][0] * B[2][2];
            double c02 = B[1][0] * B[2][1] - B[1][1] * B[2][0];
            double id = 1.0 / (B[0][0] * c00 + B[0][1] * c01 + B[0][2] * c02);
            I[0][0] = c00 * id;
            I[1][0] = c01 * id;
            I[2][0] = c02 * id;
            I[0][1] = (B[0][2] * B[2][1] - B[0][1] * B[2][2]) * id;
            I[1][1] = (B[0][0] * B[2][2] - B[0][2] * B[2][0]) * id;
            I[2][1] = (B[0][1] * B[2][0] - B[0][0] * B[2][1]) * id;
            I[0][2] = (B[0][1] * B[1][2] - B[0][2] * B[1][1]) * id;
            I[1][2] = (B[0][2] * B[1][0] - B[0][0] * B[1][2]) * id;
            I[2][2] = (B[0][0] * B[1][1] - B[0][1] * B[1][0]) * id;
        }
#pragma unroll
        for (int i = 0; i < BS; ++i) {
            const int row = nd * BS + i;
            double zi = 0.0;
#pragma unroll
            for (int k = 0; k < BS; ++k) {
                s_minv[row * BS + k] = I[i][k];
                zi += I[i][k] * rb[k];
            }
            s_x[row] = a.warm ? a.x[r0 + row] : 0.0;
            s_r[row] = rb[i];
            s_p[row] = 0.0;
            s_z[row] = zi;
            a.zg[r0 + row] = zi;
            a.pg0[r0 + row] = 0.0;
            acc0 += rb[i] * zi;
            if (!a.warm) acc1 += rb[i] * rb[i];
            acc2 += rb[i] * rb[i];
        }
    }
    acc0 = block_sum(acc0);
    acc1 = block_sum(acc1);
    acc2 = block_sum(acc2);
    int slot = 0;
    if (tid == 0) {
        a.part[(size_t)(slot * 2 + 0) * G + bid] = acc0;
        a.part[(size_t)(slot * 2 + 1) * G + bid] = acc1;
        a.part[(size_t)4 * G + bid] = acc2;
    }
    grid_barrier(a.bar, target, G);
    prefetch_window(a.pg0);
    sum_partials<2>(a.part + (size_t)slot * 2 * G, G, s_red);
    const double rz0 = s_red[0], bb0 = s_red[1];
    __syncthreads();
    sum_partials<1>(a.part + (size_t)4 * G, G, s_red);
    const double rr0 = s_red[0];
    __syncthreads();
    slot ^= 1;
    double rz = rz0;
    const double bb = bb0;
    double rr = rr0;  // ||b - A x0||^2 (= bb for a cold start)
    double tol2 = a.rtol * a.rtol * bb;
    if (a.atol * a.atol > tol2) tol2 = a.atol * a.atol;
    int it = 0;
    int status = 0;
    double rz_old = 1.0;
    if (rr > tol2 && bb > 0.0) {
        while (it < a.maxit) {
            const double beta = (it == 0) ? 0.0 : rz / rz_old;
            const double* __restrict__ pold = (it & 1) ? a.pg1 : a.pg0;
            double* __restrict__ pnew = (it & 1) ? a.pg0 : a.pg1;
            // ---- S2
            double pq = 0.0;
            if (windowed) {
                cp_async_commit_wait_all();  // the window copies started after the previous barrier
                __syncthreads();
                for (int i = tid; i < rows; i += nt) {
                    const int ka = s_rp[i], kb = s_rp[i + 1];
                    double s = 0.0;
                    for (int k = ka; k < kb; ++k) {
                        const int c = s_cols[k];
                        s = fma(s_vals[k], fma(beta, s_pw[c], s_zw[c]), s);
                    }
                    const double pi = fma(beta, s_p[i], s_z[i]);
                    s_p[i] = pi;
                    pnew[r0 + i] = pi;
                    s_q[i] = s;
                    pq = fma(pi, s, pq);
                }
            } else {
                for (int i = tid; i < rows; i += nt) {
                    const int ka = s_rp[i], kb = s_rp[i + 1];
                    double s = 0.0;
                    for (int k = ka; k < kb; ++k) {
                        const int c = s_cols[k];
                        const double pj = fma(beta, __ldcg(&pold[c]), __ldcg(&a.zg[c]));
                        s = fma(s_vals[k], pj, s);
                    }
                    const double pi = fma(beta, s_p[i], s_z[i]);
                    s_p[i] = pi;
                    pnew[r0 + i] = pi;
                    s_q[i] = s;
                    pq = fma(pi, s, pq);
                }
            }
            pq = block_sum(pq);
            if (tid == 0) a.part[(size_t)(slot * 2 + 0) * G + bid] = pq;
            grid_barrier<false>(a.bar, target, G);
            sum_partials<1>(a.part + (size_t)slot * 2 * G, G, s_red);
            slot ^= 1;
            const double alpha = rz / s_red[0];
            // ---- S3
            for (int i = tid; i < rows; i += nt) {
                s_x[i] = fma(alpha, s_p[i], s_x[i]);
                s_r[i] = fma(-alpha, s_q[i], s_r[i]);
            }
            if (BS > 1) __syncthreads();
            double a0 = 0.0, a1 = 0.0;
            for (int i = tid; i < rows; i += nt) {
                const int nd0 = (i / BS) * BS;
                double zi = 0.0;
#pragma unroll
                for (int k = 0; k < BS; ++k) zi = fma(s_minv[i * BS + k], s_r[nd0 + k], zi);
                const double ri = s_r[i];
                s_z[i] = zi;
                a.zg[r0 + i] = zi;
                a0 = fma(ri, zi, a0);
                a1 = fma(ri, ri, a1);
            }
            block_sum2(a0, a1);
            double2* pair = reinterpret_cast<double2*>(a.part + (size_t)6 * G) + (size_t)slot * G;  // [2 slots][G] pairs
            if (tid == 0) pair[bid] = make_double2(a0, a1);
            grid_barrier<false>(a.bar, target, G);
            prefetch_window(pnew);  // p of this iteration = p_old of the next
            sum_partials_pair(pair, G, s_red);
            slot ^= 1;
            rz_old = rz;
            rz = s_red[0];
            rr = s_red[1];
            ++it;
            if (!(rr == rr) || !(rz == rz)) {
                status = 2;
                break;
            }
            if (!(rr > tol2)) break;
        }
    }
    cp_async_commit_wait_all();  // a window copy may still be in flight when the loop exits
    for (int i = tid; i < rows; i += nt) a.x[r0 + i] = (bb > 0.0) ? s_x[i] : 0.0;
    if (bid == 0 && tid == 0) {
        a.out_sc[0] = rr;
        a.out_sc[1] = bb;
        a.out_fl[0] = it;
        a.out_fl[1] = status;
    }
}

int32_t pgd_pcg_finish_impl(pgd_ctx* h, int32_t* h_iters, double* h_relres);

// returns 0 ok, 1 = not resident (caller falls back), <0 / >1 errors as usual.  defer != 0: enqueue only, the
// results are collected by pgd_pcg_finish_impl (the caller knows from an earlier solve that the system fits)
int32_t pgd_pcg_resident(pgd_ctx* h, const int32_t* rp, const int32_t* ci, const double* va, const double* b, double* x,
                         int64_t n, int64_t nnz_hint, double rtol, double atol, int maxit, int block, double* work,
                         int32_t* h_iters, double* h_relres, cudaStream_t st, int warm, int defer) {
    {  // one solve in flight per handle: its results sit in the page-locked block until they are collected
        int32_t rc0 = pgd_pcg_finish_impl(h, nullptr, nullptr);
        if (rc0 < 0) return rc0;
    }
    const int G = h->sm_count;
    const size_t smem_max = 200 * 1024;
    // capacity estimate from the mean slice (+25 % slack for uneven rows); the kernel re-checks exactly
    int cap_rows = (int)((n + G - 1) / G);
    cap_rows = (cap_rows + cap_rows / 4 + 8 * block + 1) & ~1;  // even: the windows behind the row arrays stay 16-byte aligned
    if ((reinterpret_cast<uintptr_t>(work) & 15) != 0) return 1;  // the window copies are 16-byte transfers
    int64_t cap_nnz64 = (nnz_hint + G - 1) / G;
    cap_nnz64 = cap_nnz64 + cap_nnz64 / 4 + 64;
    if (cap_nnz64 > (1 << 22)) return 1;
    int cap_nnz = (int)((cap_nnz64 + 1) & ~(int64_t)1);
    size_t bytes = sizeof(double) * ((size_t)cap_nnz + (size_t)cap_rows * (5 + block)) +
                   sizeof(int) * ((size_t)cap_nnz + cap_rows + 1);
    if (bytes > smem_max) return 1;
    // whatever shared memory is left becomes the p-window (column range of the slice)
    int cap_win = (int)((smem_max - bytes) / (2 * sizeof(double)));
    if (cap_win > 8192) cap_win = 8192;
    cap_win &= ~1;
    bytes += 2 * sizeof(double) * (size_t)cap_win;
    ResArgs a;
    a.rowptr = rp;
    a.colidx = ci;
    a.vals = va;
    a.b = b;
    a.x = x;
    a.n = n;
    a.rtol = rtol;
    a.atol = atol;
    a.maxit = maxit;
    const int64_t ns = n + (n & 1);  // even spacing (work holds (5 + block) n + 8 doubles)
    a.zg = work;
    a.pg0 = work + ns;
    a.pg1 = work + 2 * ns;
    a.part = h->partials;
    a.bar = h->counters + (PGD_MAX_COUNTERS - 1);
    a.out_sc = h->scalars + 32;
    a.out_fl = h->flags + 8;
    a.cap_nnz = cap_nnz;
    a.cap_rows = cap_rows;
    a.cap_win = cap_win;
    a.warm = warm;
    PGD_CUDA(h, cudaMemsetAsync(a.bar, 0, sizeof(unsigned int), st));
    PGD_CUDA(h, cudaMemsetAsync(a.out_fl, 0, 2 * sizeof(int), st));
    void* kargs[] = {&a};
    const void* fn = (block == 1) ? (const void*)k_pcg_resident<1>
                                  : (block == 2 ? (const void*)k_pcg_resident<2> : (const void*)k_pcg_resident<3>);
    PGD_CUDA(h, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    PGD_CUDA(h, cudaEventRecord(h->ev0, st));
    PGD_CUDA(h, cudaLaunchCooperativeKernel(fn, dim3(G), dim3(RES_THREADS), kargs, bytes, st));
    h->n_launches += 1;
    PGD_CUDA(h, cudaEventRecord(h->ev1, st));
    // results -> the handle's page-locked block; read by pgd_pcg_finish_impl after the stream has drained
    char* pin = static_cast<char*>(h->pinned) + 256;
    PGD_CUDA(h, cudaMemcpyAsync(pin, a.out_fl, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    PGD_CUDA(h, cudaMemcpyAsync(pin + 64, a.out_sc, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
    PGD_CUDA(h, cudaEventRecord(h->ev_done, st));
    h->res_pending = 1;
    h->res_kind = 1;
    h->res_stream = (void*)st;
    if (defer) {
        if (h_iters) *h_iters = -1;
        return 0;
    }
    int32_t rc = pgd_pcg_finish_impl(h, h_iters, h_relres);
    if (rc == 0) {  // this system fits the SMs: later solves on it may be started without waiting
        h->fit_key = (const void*)rp;
        h->fit_n = n;
        h->fit_block = block;
    }
    return rc;
}

// waits for a started solve; 0 ok, 1 = the slices did not fit (nothing was solved), -3 NaN; *h_iters = -1 if none was pending
int32_t pgd_pcg_finish_impl(pgd_ctx* h, int32_t* h_iters, double* h_relres) {
    if (!h->res_pending) {
        if (h_iters) *h_iters = -1;
        return 0;
    }
    h->res_pending = 0;
    const int kind = h->res_kind;
    h->res_kind = 0;
    PGD_CUDA(h, cudaEventSynchronize(h->ev_done));
    const char* pin = static_cast<const char*>(h->pinned) + 256;
    int hf[2];
    double hs[2];
    memcpy(hf, pin, sizeof(hf));
    memcpy(hs, pin + 64, sizeof(hs));
    if (kind != 2 && hf[1] == 1) return 1;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->pcg_ms += ms;
    h->pcg_solves += 1;
    h->pcg_iters += hf[0];
    if (kind != 2) h->pcg_resident_solves += 1;
    if (h_iters) *h_iters = hf[0];
    if (h_relres) *h_relres = (hs[1] > 0.0) ? sqrt(hs[0] / hs[1]) : 0.0;
    if (kind == 2 && hf[1] == 3) {  // persistent streaming kernel (pgd_pcg_persist_start, single GPU): a CTA did not arrive
        snprintf(h->err, sizeof(h->err), "pgd_pcg_persist_start: a CTA did not arrive within %d ms", h->opt_spin_ms);
        return -6;
    }
    if (hf[1] == 2) {
        snprintf(h->err, sizeof(h->err), "pgd_pcg: NaN encountered (matrix not SPD?)");
        return -3;
    }
    return 0;
}
