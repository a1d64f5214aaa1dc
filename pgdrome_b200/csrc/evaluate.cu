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

extern "C" int32_t pgd_eval_gemm_f64(pgd_handle_t h, const double* d_W, int64_t ldw, const double* d_X, int64_t ldx,
                                     int32_t R, int64_t C, int64_t N, double* d_U, int64_t ldu, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_W && d_X && d_U && R > 0 && C >= 0 && N >= 0 && ldw >= C && ldx >= N && ldu >= N, "bad arguments");
    if (C == 0 || N == 0) return 0;
    int64_t tiles = ((C + GM_TM - 1) / GM_TM) * ((N + GM_TN - 1) / GM_TN);
    PGD_ARG(h, tiles < ((int64_t)1 << 31), "too many tiles");
    size_t smem = sizeof(double) * 2 * GM_KC * GM_LD;
    PGD_CUDA(h, cudaFuncSetAttribute(k_eval_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int vec_ok = ((ldu % 2) == 0) && ((reinterpret_cast<uintptr_t>(d_U) % 16) == 0);
    k_eval_gemm<<<(unsigned int)tiles, 256, smem, (cudaStream_t)stream>>>(d_W, ldw, d_X, ldx, R, C, N, d_U, ldu, vec_ok);
    PGD_LAUNCH_OK(h);
    return 0;
}
