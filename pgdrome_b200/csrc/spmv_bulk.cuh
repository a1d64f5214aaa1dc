// TMA-pipelined CSR SpMV core for systems that do not fit on chip (the HBM-bound regime).
//
// A persistent CTA walks row blocks of RB = 512 / LPR consecutive rows.  The (column, value)
// entries of a row block are ONE contiguous range of the CSR arrays, so they are brought into
// shared memory with two bulk asynchronous copies (cp.async.bulk global -> shared, completion on
// an mbarrier: the TMA engine, no registers, no per-thread address arithmetic) into a two-stage
// ring: while the threads work on tile t, the copies of tile t+1 (same row block, or the next one
// of this CTA) are already in flight.  The compute step is row-major on purpose: LPR (1, 2, 4 ...)
// lanes own one row and walk its entries out of shared memory, so the 32 lanes of a warp gather
// x[col] for the SAME entry position of 32/LPR CONSECUTIVE rows -- on a mesh-ordered matrix these are
// neighbouring columns, i.e. 2-4 cache lines per warp-wide gather instead of ~15 when a warp walks
// consecutive nonzeros (ncu on the first version of this kernel: L1TEX wavefronts, not DRAM, were
// the top unit, profiles/README.md).  Row sums are formed in a fixed order (LPR lanes, ascending k,
// xor-shuffle tree), so results are bitwise reproducible.  Any row length is handled: a row block
// whose range exceeds the stage capacity is simply cut into several tiles.
//
// Alignment: bulk copies need 16-byte aligned addresses and sizes.  A tile [kt, kend) is widened
// to [kt & ~3, roundup4(kend)) clipped to nnz & ~3; the (at most 3) entries of the matrix tail that
// the clip drops are fetched with ordinary loads.  colidx / vals base pointers must be 16-byte
// aligned (checked by the launcher, which otherwise uses the register-staged kernel).
#pragma once
#include "common.cuh"

#ifndef BK_THREADS
#define BK_THREADS 256
#endif
#ifndef BK_CAP
#define BK_CAP 4480                     // entries per stage
#endif
#ifndef BK_CTAS_PER_SM
#define BK_CTAS_PER_SM 2                // resident CTAs per SM (shared memory: 2 stages x BK_BUF x 12 B each)
#endif
#define BK_BUF (BK_CAP + 8)             // + alignment slack on both ends
#define BK_SMEM_BYTES (2 * BK_BUF * 12 + 2 * 260 * 4 + 64)

__device__ __forceinline__ uint32_t bk_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bk_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bk_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bk_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bk_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bk_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     bk_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(bk_smem_u32(bar))
                 : "memory");
}
// Wait for the phase with the given parity.  The poll is bounded: a copy that never completes (a bug, never seen in
// a correct run) traps after ~4 s instead of hanging the GPU.
__device__ __forceinline__ void bk_mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = bk_smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (spins > (1u << 22)) __trap();  // each failed try_wait already waits a hardware-defined interval (~1 us)
    }
}

struct BkGatherX {
    const double* __restrict__ x;
    __device__ __forceinline__ double operator()(int c) const { return __ldg(&x[c]); }
};
struct BkGatherZP {  // PCG: p_new[c] = z[c] + beta p_old[c], formed on the fly
    const double* __restrict__ z;
    const double* __restrict__ p;
    double beta;
    __device__ __forceinline__ double operator()(int c) const { return fma(beta, __ldg(&p[c]), __ldg(&z[c])); }
};

// lanes per row from the mean row length (device side: nnz = rowptr[n] is read by the kernel):
// the largest row block RB = BK_THREADS / LPR whose expected entry count fits one stage
template <int THREADS>
__device__ __forceinline__ int bk_pick_lpr(int64_t n, int64_t nnz) {
    const double avg = (double)nnz / (double)(n > 0 ? n : 1);
    const double fill = 0.93 * BK_CAP;
    int lpr = THREADS / 256 > 1 ? THREADS / 256 : 1;  // at most 256 rows per block (s_rp holds 260 entries)
    while (lpr < 16 && avg * (THREADS / lpr) > fill) lpr *= 2;
    return lpr;
}

// Calls epi(row, s) exactly once for every row owned by this CTA (s = (A v)[row], valid in the
// row's lane 0; epi is invoked by that lane only).  All BK_THREADS threads must call.
// THREADS = CTA size: 256 (default; register-heavy callers such as the PCG gather keep 2 CTAs per SM) or 512 (plain
// SpMV / functionals: two lanes per row halve the serial gather chain, measured 80 % vs 74 % of the HBM peak)
template <class Gather, class Epi, int THREADS = BK_THREADS>
__device__ __forceinline__ void bk_spmv_rows(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                             const double* __restrict__ vals, int64_t n, const Gather& g, Epi&& epi,
                                             unsigned char* smem, uint32_t* t_state = nullptr) {
    double* const s_vals0 = reinterpret_cast<double*>(smem);             // [2][BK_BUF]
    int* const s_cols0 = reinterpret_cast<int*>(s_vals0 + 2 * BK_BUF);   // [2][BK_BUF]
    int* const s_rp0 = s_cols0 + 2 * BK_BUF;                             // [2][260]
    uint64_t* const bars = reinterpret_cast<uint64_t*>(s_rp0 + 2 * 260);  // 8-byte aligned: sizes above are multiples of 8
#define s_vals(i) (s_vals0 + (i) * BK_BUF)
#define s_cols(i) (s_cols0 + (i) * BK_BUF)
#define s_rp(i) (s_rp0 + (i) * 260)

    const int tid = threadIdx.x;
    const int64_t nnz = __ldg(&rowptr[n]);
    const int nnz_al = (int)(nnz & ~(int64_t)3);
    const int LPR = bk_pick_lpr<THREADS>(n, nnz);
    const int RB = THREADS / LPR;
    const int64_t nblk = (n + RB - 1) / RB;
    int64_t rb = blockIdx.x;
    if (rb >= nblk) return;
    const int myrow = tid / LPR, lane = tid % LPR;

    auto issue = [&](uint32_t t, int kt, int kend) {  // thread 0 only
        const int st = t & 1;
        const int ka = kt & ~3;
        int kb = min((kend + 3) & ~3, nnz_al);
        if (kb < ka) kb = ka;
        const uint32_t cnt = (uint32_t)(kb - ka);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bk_mbar_expect_tx(&bars[st], cnt * 12u);
        if (cnt) {
            bk_bulk_g2s(s_cols(st), colidx + ka, cnt * 4u, &bars[st]);
            bk_bulk_g2s(s_vals(st), vals + ka, cnt * 8u, &bars[st]);
        }
    };

    int64_t r0 = rb * RB;
    int nr = (int)min((int64_t)RB, n - r0);
    for (int i = tid; i <= nr; i += THREADS) s_rp(0)[i] = __ldg(&rowptr[r0 + i]);
    // t_state (persistent callers): the mbarriers were initialised once by bk_init_barriers and the running tile
    // counter (stage = t & 1, phase parity = (t >> 1) & 1) carries over from call to call
    if (tid == 0 && !t_state) {
        bk_mbar_init(&bars[0], 1);
        bk_mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int pb = 0;
    int k0 = s_rp(0)[0], k1 = s_rp(0)[nr];
    uint32_t t = t_state ? *t_state : 0u;
    if (tid == 0) issue(t, k0, min(k0 + BK_CAP, k1));

    while (true) {
        // prefetch the NEXT row block's rowptr slice into registers (RB + 1 <= BK_THREADS + 1 entries)
        const int64_t rbn = rb + gridDim.x;
        const bool has_next = rbn < nblk;
        const int64_t r0n = rbn * RB;
        const int nrn = has_next ? (int)min((int64_t)RB, n - r0n) : 0;
        int rp_reg = 0, rp_end = 0;
        if (has_next && tid < nrn) rp_reg = __ldg(&rowptr[r0n + tid]);
        if (has_next && tid == 0) rp_end = __ldg(&rowptr[r0n + nrn]);

        int a = 0, b = 0;
        if (myrow < nr) {
            a = s_rp(pb)[myrow];
            b = s_rp(pb)[myrow + 1];
        }
        double s = 0.0;
        int kt = k0;
        while (true) {
            const int kend = min(kt + BK_CAP, k1);
            const bool next_same = kend < k1;
            if (!next_same && has_next) {
                if (tid < nrn) s_rp(pb ^ 1)[tid] = rp_reg;
                if (tid == 0) s_rp(pb ^ 1)[nrn] = rp_end;
            }
            __syncthreads();  // (a) stage (t+1)&1 is free again; s_rp[pb^1] is visible
            if (tid == 0) {
                if (next_same) issue(t + 1, kend, min(kend + BK_CAP, k1));
                else if (has_next) {
                    const int nk0 = s_rp(pb ^ 1)[0], nk1 = s_rp(pb ^ 1)[nrn];
                    issue(t + 1, nk0, min(nk0 + BK_CAP, nk1));
                }
            }
            const int st = t & 1;
            const int ka = kt & ~3;
            bk_mbar_wait(&bars[st], (t >> 1) & 1);
            double* sv = s_vals(st);
            const int* sc = s_cols(st);
            if (kend > nnz_al) {  // matrix tail not covered by the 16-byte granular copies (uniform branch)
                const int kq = max(kt, nnz_al);
                if (tid < kend - kq) {
                    sv[kq + tid - ka] = __ldg(&vals[kq + tid]);
                    const_cast<int*>(sc)[kq + tid - ka] = __ldg(&colidx[kq + tid]);
                }
                __syncthreads();
            }
            // row-major compute straight out of the stage buffers: 4 independent gathers in flight per lane
            const int lo = max(a, kt) - ka, hi = min(b, kend) - ka;
            int k = lo + lane;
            for (; k + 3 * LPR < hi; k += 4 * LPR) {
                const int c0 = sc[k], c1 = sc[k + LPR], c2 = sc[k + 2 * LPR], c3 = sc[k + 3 * LPR];
                const double x0 = g(c0), x1 = g(c1), x2 = g(c2), x3 = g(c3);
                s = fma(sv[k], x0, s);
                s = fma(sv[k + LPR], x1, s);
                s = fma(sv[k + 2 * LPR], x2, s);
                s = fma(sv[k + 3 * LPR], x3, s);
            }
            for (; k < hi; k += LPR) s = fma(sv[k], g(sc[k]), s);
            ++t;
            if (!next_same) break;
            kt = kend;
        }
        for (int o = LPR >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (myrow < nr && lane == 0) epi(r0 + myrow, s);
        if (!has_next) break;
        pb ^= 1;
        rb = rbn;
        r0 = r0n;
        nr = nrn;
        k0 = s_rp(pb)[0];
        k1 = s_rp(pb)[nr];
    }
    if (t_state) *t_state = t;
}

// one-time initialisation of the two stage barriers for persistent callers of bk_spmv_rows (thread 0; follow with
// __syncthreads before the first call)
__device__ __forceinline__ void bk_init_barriers(unsigned char* smem) {
    double* const s_vals0 = reinterpret_cast<double*>(smem);
    int* const s_cols0 = reinterpret_cast<int*>(s_vals0 + 2 * BK_BUF);
    uint64_t* const bars = reinterpret_cast<uint64_t*>(s_cols0 + 2 * BK_BUF + 2 * 260);
    bk_mbar_init(&bars[0], 1);
    bk_mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
#undef s_vals
#undef s_cols
#undef s_rp
