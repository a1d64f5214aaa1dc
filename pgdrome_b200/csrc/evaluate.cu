// PGD.evaluate on the device (pgdrome/model.py:724-860):
//   k_eval_weights : W[k,c] = prod_i phi_{i,k}(p[c,i])   1-D Lagrange P1/P2 free dimensions
//   k_eval_gemv    : u[n]   = sum_k X[k,n] w[k]          single parameter point (HBM-bound)
//   k_eval_gemm    : U[c,n] = sum_k W[k,c] X[k,n]        FP64 tensor cores (DMMA m8n8k4), the only
//                    dense contraction of the path; K = number of modes (tens), output-write heavy.
#include "common.cuh"

#define EV_MAX_FREE 8
struct EvalDims {
    const double* xs[EV_MAX_FREE];
    const int32_t* cd[EV_MAX_FREE];
    const double* Phi[EV_MAX_FREE];
    int nc[EV_MAX_FREE];
    int deg[EV_MAX_FREE];
    int64_t ld[EV_MAX_FREE];
    int n_free;
};

__global__ void __launch_bounds__(128) k_eval_weights(EvalDims D, int R, const double* __restrict__ pts, int64_t C,
                                                      double* __restrict__ W, int* flag) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    int dof[EV_MAX_FREE][3];
    double shp[EV_MAX_FREE][3];
    for (int i = 0; i < D.n_free; ++i) {
        const double* xs = D.xs[i];
        const int nc = D.nc[i];
        double p = pts[c * D.n_free + i];
        double lo = xs[0], hi = xs[nc];
        double tol = 1e-12 * fmax(fabs(hi - lo), 1e-300);
        if (!(p >= lo - tol && p <= hi + tol)) {
            atomicCAS(flag, 0, 1 + i);
            p = fmin(fmax(p, lo), hi);
        }
        int a = 0, b = nc;  // find e with xs[e] <= p < xs[e+1]
        while (b - a > 1) {
            int m = (a + b) >> 1;
            if (p >= xs[m]) a = m;
            else b = m;
        }
        double xi = (p - xs[a]) / (xs[a + 1] - xs[a]);
        const int nl = D.deg[i] + 1;
        for (int l = 0; l < nl; ++l) dof[i][l] = D.cd[i][a * nl + l];
        if (D.deg[i] == 1) {
            shp[i][0] = 1.0 - xi;
            shp[i][1] = xi;
            shp[i][2] = 0.0;
            dof[i][2] = dof[i][0];
        } else {
            shp[i][0] = (1.0 - xi) * (1.0 - 2.0 * xi);
            shp[i][1] = xi * (2.0 * xi - 1.0);
            shp[i][2] = 4.0 * xi * (1.0 - xi);
        }
    }
    for (int k = 0; k < R; ++k) {
        double w = 1.0;
        for (int i = 0; i < D.n_free; ++i) {
            const double* ph = D.Phi[i] + (size_t)k * D.ld[i];
            double v = shp[i][0] * __ldg(&ph[dof[i][0]]) + shp[i][1] * __ldg(&ph[dof[i][1]]);
            if (D.deg[i] == 2) v += shp[i][2] * __ldg(&ph[dof[i][2]]);
            w *= v;
        }
        W[(size_t)k * C + c] = w;
    }
}

extern "C" int32_t pgd_eval_weights(pgd_handle_t h, int32_t n_free, const double* const* h_xs, const int32_t* const* h_cd,
                                    const double* const* h_Phi, const int32_t* h_nc, const int32_t* h_deg,
                                    const int64_t* h_ld, int32_t R, const double* d_points, int64_t C, double* d_W,
                                    int32_t* d_flag, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, n_free >= 0 && n_free <= EV_MAX_FREE, "n_free must be <= 8");
    PGD_ARG(h, d_W && d_flag && R > 0 && C >= 0 && (n_free == 0 || d_points), "bad arguments");
    if (C == 0) return 0;
    EvalDims D;
    memset(&D, 0, sizeof(D));
    D.n_free = n_free;
    for (int i = 0; i < n_free; ++i) {
        PGD_ARG(h, h_deg[i] == 1 || h_deg[i] == 2, "free-dimension degree must be 1 or 2");
        PGD_ARG(h, h_nc[i] >= 1, "free dimension needs at least one cell");
        D.xs[i] = h_xs[i];
        D.cd[i] = h_cd[i];
        D.Phi[i] = h_Phi[i];
        D.nc[i] = h_nc[i];
        D.deg[i] = h_deg[i];
        D.ld[i] = h_ld[i];
    }
    k_eval_weights<<<pgd_blocks(C, 128), 128, 0, (cudaStream_t)stream>>>(D, R, d_points, C, d_W, d_flag);
    PGD_LAUNCH_OK(h);
    return 0;
}

// ----------------------------------------------------------------------------- GEMV
__global__ void __launch_bounds__(256) k_eval_gemv(const double* __restrict__ X, int64_t ldx, int R,
                                                   const double* __restrict__ w, int64_t N, double* __restrict__ u) {
    extern __shared__ double sw[];
    for (int k = threadIdx.x; k < R; k += blockDim.x) sw[k] = w[k];
    __syncthreads();
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < N; n += stride) {
        double s = 0.0;
#pragma unroll 8
        for (int k = 0; k < R; ++k) s = fma(__ldcs(&X[(size_t)k * ldx + n]), sw[k], s);
        u[n] = s;
    }
}

extern "C" int32_t pgd_eval_gemv(pgd_handle_t h, const double* d_X, int64_t ldx, int32_t R, const double* d_w, int64_t N,
                                 double* d_u, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_X && d_w && d_u && R > 0 && R <= 4096 && N >= 0 && ldx >= N, "bad arguments");
    if (N == 0) return 0;
    unsigned int blocks = pgd_blocks(N, 256);
    unsigned int cap = (unsigned int)h->sm_count * 16;
    if (blocks > cap) blocks = cap;
    k_eval_gemv<<<blocks, 256, sizeof(double) * R, (cudaStream_t)stream>>>(d_X, ldx, R, d_w, N, d_u);
    PGD_LAUNCH_OK(h);
    return 0;
}

// ----------------------------------------------------------------------------- DMMA GEMM
// CTA tile 128(c) x 128(n), K chunk <= 64 staged in shared memory (rows padded to 136 doubles so
// that the 4 k-rows x 8 m-columns of a fragment load fall in distinct banks).  8 warps as 4(M) x
// 2(N); warp tile 32 x 64 = 4 x 8 DMMA m8n8k4 tiles (64 accumulator doubles per lane).
#define GM_TM 128
#define GM_TN 128
#define GM_KC 64
#define GM_LD 136

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 1) k_eval_gemm(const double* __restrict__ W, int64_t ldw, const double* __restrict__ X,
                                                      int64_t ldx, int R, int64_t C, int64_t N, double* __restrict__ U,
                                                      int64_t ldu, int vec_ok) {
    extern __shared__ double sm[];
    double* As = sm;                    // [GM_KC][GM_LD]  As[k][m] = W[k, c0+m]
    double* Bs = sm + GM_KC * GM_LD;    // [GM_KC][GM_LD]  Bs[k][n] = X[k, n0+n]
    const int64_t n_tiles_n = (N + GM_TN - 1) / GM_TN;
    const int64_t tile = blockIdx.x;
    const int64_t c0 = (tile / n_tiles_n) * GM_TM;
    const int64_t n0 = (tile % n_tiles_n) * GM_TN;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 64;
    const int fr = lane >> 2, fk = lane & 3;
    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int k0 = 0; k0 < R; k0 += GM_KC) {
        const int kc = min(GM_KC, R - k0);
        const int kc4 = (kc + 3) & ~3;
        __syncthreads();
        for (int idx = threadIdx.x; idx < kc4 * GM_TM; idx += 256) {
            int k = idx / GM_TM, m = idx - k * GM_TM;
            double a = 0.0, b = 0.0;
            if (k < kc) {
                if (c0 + m < C) a = __ldg(&W[(size_t)(k0 + k) * ldw + c0 + m]);
                if (n0 + m < N) b = __ldg(&X[(size_t)(k0 + k) * ldx + n0 + m]);
            }
            As[k * GM_LD + m] = a;
            Bs[k * GM_LD + m] = b;
        }
        __syncthreads();
        for (int kk = 0; kk < kc4; kk += 4) {
            double af[4], bf[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = As[(kk + fk) * GM_LD + wm + i * 8 + fr];
#pragma unroll
            for (int j = 0; j < 8; ++j) bf[j] = Bs[(kk + fk) * GM_LD + wn + j * 8 + fr];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
    }
    // epilogue: lane holds rows fr, cols 2*fk, 2*fk+1 of every 8x8 tile
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t c = c0 + wm + i * 8 + fr;
        if (c >= C) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int64_t n = n0 + wn + j * 8 + 2 * fk;
            double* dst = U + (size_t)c * ldu + n;
            if (vec_ok && n + 1 < N) {
                __stcs(reinterpret_cast<double2*>(dst), make_double2(acc[i][j][0], acc[i][j][1]));
            } else {
                if (n < N) __stcs(dst, acc[i][j][0]);
                if (n + 1 < N) __stcs(dst + 1, acc[i][j][1]);
            }
        }
    }
}


// ---- v2: 2 CTAs per SM, A tile resident, B tiles double-buffered with cp.async -----------------
// CTA = one 128-row c-tile x a run of GM2_RUN consecutive 64-column n-tiles.  W's tile (A) is staged
// once; X's tiles (B) stream through a 2-deep cp.async ring so that the loads of tile t+1 overlap the
// DMMAs and the streaming stores of tile t, and the second resident CTA fills the remaining bubbles
// (warp tile 32 x 32 = 64 accumulator registers => 2 CTAs of 256 threads per SM).  Leading dimensions
// 132 / 68 make every fragment load bank-conflict free (4 k-rows x 4 m-columns per half-warp hit 16
// distinct bank pairs).  Requires 16-byte aligned W/X/U and even ldw/ldx/ldu (launcher checks).
#define GM2_TM 128
#define GM2_TN 64
#define GM2_KMAX 64
#define GM2_LDA 132
#define GM2_LDB 68
#define GM2_RUN 16

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__global__ void __launch_bounds__(256, 2) k_eval_gemm2(const double* __restrict__ W, int64_t ldw, const double* __restrict__ X,
                                                       int64_t ldx, int R, int64_t C, int64_t N, double* __restrict__ U,
                                                       int64_t ldu, int64_t n_tiles, int64_t n_groups) {
    extern __shared__ __align__(16) double sm2[];
    const int kc4 = (R + 3) & ~3;
    double* As = sm2;                       // [kc4][GM2_LDA]
    double* Bs = sm2 + kc4 * GM2_LDA;       // [2][kc4][GM2_LDB]
    const int64_t ct = blockIdx.x / n_groups, ng = blockIdx.x % n_groups;
    const int64_t c0 = ct * GM2_TM;
    const int64_t t0 = ng * GM2_RUN;
    const int nt = (int)min((int64_t)GM2_RUN, n_tiles - t0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = (warp & 3) * 32, wn = (warp >> 2) * 32;
    const int fr = lane >> 2, fk = lane & 3;
    // zero the K padding rows once (never touched by cp.async)
    for (int idx = tid; idx < (kc4 - R) * GM2_LDA; idx += 256) As[R * GM2_LDA + idx] = 0.0;
    for (int b = 0; b < 2; ++b)
        for (int idx = tid; idx < (kc4 - R) * GM2_LDB; idx += 256) Bs[(b * kc4 + R) * GM2_LDB + idx] = 0.0;
    // A tile: R rows x 64 chunks of 16 B
    for (int idx = tid; idx < R * (GM2_TM / 2); idx += 256) {
        const int k = idx / (GM2_TM / 2), ch = idx - k * (GM2_TM / 2);
        const int64_t c = c0 + 2 * ch;
        const int64_t left = (C - c) * 8;
        const int bytes = left >= 16 ? 16 : (left > 0 ? (int)left : 0);
        cp_async16(&As[k * GM2_LDA + 2 * ch], &W[(size_t)k * ldw + (bytes ? c : 0)], bytes);
    }
    auto load_B = [&](int t, int buf) {
        const int64_t n0 = (t0 + t) * GM2_TN;
        double* dst = Bs + (size_t)buf * kc4 * GM2_LDB;
        for (int idx = tid; idx < R * (GM2_TN / 2); idx += 256) {
            const int k = idx / (GM2_TN / 2), ch = idx - k * (GM2_TN / 2);
            const int64_t n = n0 + 2 * ch;
            const int64_t left = (N - n) * 8;
            const int bytes = left >= 16 ? 16 : (left > 0 ? (int)left : 0);
            cp_async16(&dst[k * GM2_LDB + 2 * ch], &X[(size_t)k * ldx + (bytes ? n : 0)], bytes);
        }
    };
    load_B(0, 0);
    cp_async_commit();
    for (int t = 0; t < nt; ++t) {
        if (t + 1 < nt) load_B(t + 1, (t + 1) & 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const double* Bt = Bs + (size_t)(t & 1) * kc4 * GM2_LDB;
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        for (int kk = 0; kk < kc4; kk += 4) {
            double af[4], bf[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) af[i] = As[(kk + fk) * GM2_LDA + wm + i * 8 + fr];
#pragma unroll
            for (int j = 0; j < 4; ++j) bf[j] = Bt[(kk + fk) * GM2_LDB + wn + j * 8 + fr];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
        const int64_t n0 = (t0 + t) * GM2_TN;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int64_t c = c0 + wm + i * 8 + fr;
            if (c >= C) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t n = n0 + wn + j * 8 + 2 * fk;
                double* dst = U + (size_t)c * ldu + n;
                if (n + 1 < N) __stcs(reinterpret_cast<double2*>(dst), make_double2(acc[i][j][0], acc[i][j][1]));
                else if (n < N) __stcs(dst, acc[i][j][0]);
            }
        }
        __syncthreads();  // every warp is done with Bs[t & 1] before iteration t + 1 prefetches into it
    }
}

extern "C" int32_t pgd_eval_gemm_f64(pgd_handle_t h, const double* d_W, int64_t ldw, const double* d_X, int64_t ldx,
                                     int32_t R, int64_t C, int64_t N, double* d_U, int64_t ldu, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, R > 0 && C >= 0 && N >= 0 && ldw >= C && ldx >= N && ldu >= N, "bad arguments");
    if (C == 0 || N == 0) return 0;  // empty sweep: nothing to do (pointers of empty tensors may be NULL)
    PGD_ARG(h, d_W && d_X && d_U, "null pointer");
    const bool aligned = ((ldu | ldx | ldw) % 2 == 0) &&
                         (((reinterpret_cast<uintptr_t>(d_U) | reinterpret_cast<uintptr_t>(d_X) | reinterpret_cast<uintptr_t>(d_W)) % 16) == 0);
    if (aligned && R <= GM2_KMAX) {
        const int kc4 = (R + 3) & ~3;
        const int64_t n_tiles = (N + GM2_TN - 1) / GM2_TN, c_tiles = (C + GM2_TM - 1) / GM2_TM;
        const int64_t n_groups = (n_tiles + GM2_RUN - 1) / GM2_RUN;
        PGD_ARG(h, c_tiles * n_groups < ((int64_t)1 << 31), "too many tiles");
        const size_t smem2 = sizeof(double) * (size_t)kc4 * (GM2_LDA + 2 * GM2_LDB);
        PGD_CUDA(h, cudaFuncSetAttribute(k_eval_gemm2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        k_eval_gemm2<<<(unsigned int)(c_tiles * n_groups), 256, smem2, (cudaStream_t)stream>>>(d_W, ldw, d_X, ldx, R, C, N, d_U,
                                                                                            ldu, n_tiles, n_groups);
        PGD_LAUNCH_OK(h);
        return 0;
    }
    int64_t tiles = ((C + GM_TM - 1) / GM_TM) * ((N + GM_TN - 1) / GM_TN);
    PGD_ARG(h, tiles < ((int64_t)1 << 31), "too many tiles");
    size_t smem = sizeof(double) * 2 * GM_KC * GM_LD;
    PGD_CUDA(h, cudaFuncSetAttribute(k_eval_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int vec_ok = ((ldu % 2) == 0) && ((reinterpret_cast<uintptr_t>(d_U) % 16) == 0);
    k_eval_gemm<<<(unsigned int)tiles, 256, smem, (cudaStream_t)stream>>>(d_W, ldw, d_X, ldx, R, C, N, d_U, ldu, vec_ok);
    PGD_LAUNCH_OK(h);
    return 0;
}

// ----------------------------------------------------------------------------- row-wise reductions of a sweep U [C, N]
// (PGDErrorComputation.evaluate_error / evaluate_min, _max, ... over a whole batch of parameter points: model.py:955-1086,
// 1785-1825 loop over the samples on the host, one evaluate + one NumPy reduction each).  One CTA per row, fixed-order
// block reduction => deterministic.  stats[c] = {min, max, min|.|, max|.|, sum u^2, sum (u - f)^2, sum f^2}; the last two
// need the reference rows F [C, N] (may be NULL).
#define ROWSTAT_N 7
__global__ void __launch_bounds__(256) k_row_stats(const double* __restrict__ U, int64_t ldu, const double* __restrict__ F,
                                                   int64_t ldf, int64_t N, double* __restrict__ out) {
    const int64_t c = blockIdx.x;
    const double* u = U + c * ldu;
    const double* f = F ? F + c * ldf : nullptr;
    double mn = INFINITY, mx = -INFINITY, mna = INFINITY, mxa = 0.0, s2 = 0.0, d2 = 0.0, f2 = 0.0;
    for (int64_t i = threadIdx.x; i < N; i += blockDim.x) {
        const double v = u[i], av = fabs(v);
        mn = fmin(mn, v);
        mx = fmax(mx, v);
        mna = fmin(mna, av);
        mxa = fmax(mxa, av);
        s2 = fma(v, v, s2);
        if (f) {
            const double w = f[i], d = v - w;
            d2 = fma(d, d, d2);
            f2 = fma(w, w, f2);
        }
    }
    __shared__ double sh[ROWSTAT_N][256];
    sh[0][threadIdx.x] = mn;
    sh[1][threadIdx.x] = mx;
    sh[2][threadIdx.x] = mna;
    sh[3][threadIdx.x] = mxa;
    sh[4][threadIdx.x] = s2;
    sh[5][threadIdx.x] = d2;
    sh[6][threadIdx.x] = f2;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            sh[0][threadIdx.x] = fmin(sh[0][threadIdx.x], sh[0][threadIdx.x + o]);
            sh[1][threadIdx.x] = fmax(sh[1][threadIdx.x], sh[1][threadIdx.x + o]);
            sh[2][threadIdx.x] = fmin(sh[2][threadIdx.x], sh[2][threadIdx.x + o]);
            sh[3][threadIdx.x] = fmax(sh[3][threadIdx.x], sh[3][threadIdx.x + o]);
            sh[4][threadIdx.x] += sh[4][threadIdx.x + o];
            sh[5][threadIdx.x] += sh[5][threadIdx.x + o];
            sh[6][threadIdx.x] += sh[6][threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x < ROWSTAT_N) out[c * ROWSTAT_N + threadIdx.x] = sh[threadIdx.x][0];
}

extern "C" int32_t pgd_row_stats(pgd_handle_t h, const double* d_U, int64_t ldu, const double* d_F, int64_t ldf, int64_t C,
                                 int64_t N, double* d_stats, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, C >= 0 && N >= 0 && ldu >= N && (!d_F || ldf >= N) && C < ((int64_t)1 << 31), "bad arguments");
    if (C == 0) return 0;
    PGD_ARG(h, d_U && d_stats, "null pointer");
    k_row_stats<<<(unsigned int)C, 256, 0, (cudaStream_t)stream>>>(d_U, ldu, d_F, ldf, N, d_stats);
    PGD_LAUNCH_OK(h);
    return 0;
}
