// Streaming CSR SpMV core for systems that do not fit on chip (the HBM-bound regime).
//
// The sub-warp-per-row kernel (sparse.cu) keeps one nonzero per lane in flight behind a
// rowptr -> colidx -> x[col] dependency chain and measured 25 % of DRAM peak on a 128^3 P1 mesh
// (profiles/r01_v0_*): latency-bound.  Here a CTA owns ST_ROWS consecutive rows, i.e. ONE contiguous
// range [rowptr[r0], rowptr[r0+ST_ROWS]) of the value / column arrays.  All 256 threads stream
// that range with independent, fully coalesced loads (4 nonzeros per thread in flight before the
// first use), multiply by the gathered vector entry and park the products in shared memory; then
// thread t sums the products of row r0+t in ascending-k order (fixed order => bitwise
// reproducible).  Rows of any length are handled by tiling the range through the ST_TILE buffer.
#pragma once
#include "common.cuh"

#define ST_ROWS 256
#define ST_TILE 4096
#define ST_THREADS 256

// Gather functor: value of the multiplied vector at column c.
struct GatherX {
    const double* __restrict__ x;
    __device__ __forceinline__ double operator()(int c) const { return __ldg(&x[c]); }
};
// PCG: p_new[c] = z[c] + beta * p_old[c] formed on the fly (p_new is only written for owned rows)
struct GatherZP {
    const double* __restrict__ z;
    const double* __restrict__ p;
    double beta;
    __device__ __forceinline__ double operator()(int c) const { return fma(beta, __ldg(&p[c]), __ldg(&z[c])); }
};

// Computes s = (A v)[r0 + tid] for tid < nr, for the row block starting at r0.  s_prod: ST_TILE
// doubles, s_rp: ST_ROWS + 1 ints of shared memory.  All threads of the CTA must call.
template <class Gather>
__device__ __forceinline__ double stream_rowblock(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                  const double* __restrict__ vals, const Gather& g, int64_t r0, int nr,
                                                  double* s_prod, int* s_rp) {
    const int tid = threadIdx.x;
    __syncthreads();  // previous row block is done with s_rp / s_prod
    for (int i = tid; i <= nr; i += ST_THREADS) s_rp[i] = __ldg(&rowptr[r0 + i]);
    __syncthreads();
    const int k0 = s_rp[0], k1 = s_rp[nr];
    int a = 0, b = 0;
    if (tid < nr) {
        a = s_rp[tid];
        b = s_rp[tid + 1];
    }
    double s = 0.0;
    for (int kt = k0; kt < k1; kt += ST_TILE) {
        const int kend = min(kt + ST_TILE, k1);
        for (int k = kt + tid; k < kend; k += 4 * ST_THREADS) {
            int c[4];
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int kk = k + u * ST_THREADS;
                c[u] = (kk < kend) ? ld_stream(&colidx[kk]) : -1;
                v[u] = (kk < kend) ? ld_stream(&vals[kk]) : 0.0;
            }
            double xv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) xv[u] = (c[u] >= 0) ? g(c[u]) : 0.0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int kk = k + u * ST_THREADS;
                if (kk < kend) s_prod[kk - kt] = v[u] * xv[u];
            }
        }
        __syncthreads();
        const int lo = max(a, kt), hi = min(b, kend);
        for (int k = lo; k < hi; ++k) s += s_prod[k - kt];
        if (kend < k1) __syncthreads();
    }
    return s;
}
