// Node-block SpMV core for vector P1/P2 operators (bs = 2, 3): the matrix stays in its CSR value order, but the
// kernel walks it by BS x BS NODE BLOCKS and only needs one column index per block.
//
// On a node-blocked space the pattern is a union of full BS x BS blocks, so the CSR rows 3nd, 3nd+1, 3nd+2 of node
// nd are three contiguous segments of equal length 3*nb whose columns are (3c, 3c+1, 3c+2) for the same nb
// neighbour nodes c.  Streaming `values` (8 B/nnz) plus the compact block-column list `bcol` (4 B per block, i.e.
// 4/9 B per nonzero for bs = 3) moves 8.44 B/nnz instead of the 12 B/nnz of plain CSR, and a lane gathers the
// BS entries of x of a block once instead of BS times (x gathers, not DRAM, are the top unit of the CSR kernel on
// 45-entry rows: ncu profiles/r02_pcg_bs3_v0).
//
// Structure as in spmv_bulk.cuh: a persistent CTA walks tiles of NBR consecutive block rows; the tile's value
// range and block-column range are two contiguous arrays brought in by cp.async.bulk (TMA) into a two-stage
// mbarrier ring.  LPR lanes own one block row, lane j takes the blocks j, j + LPR, ...; values of block j are at
// seg_i + BS*j + k in the three row segments (stride BS doubles between lanes => conflict-free shared-memory
// reads), the BS row sums are completed by a fixed xor-shuffle tree (bitwise reproducible) and handed to
// epi(row, sum) by the block row's lane 0.  NBR * BS * BS * max_blocks_per_row must fit one stage (the host
// plan picks NBR / LPR from the longest block row and falls back to the CSR kernel otherwise).
#pragma once
#include "spmv_bulk.cuh"

#define BB_BCAP (BK_CAP / 4 + 8)               // block-column entries per stage (bs = 2 is the densest case)
#define BB_MAXNBR 64
#define BB_RP (BB_MAXNBR * 3 + 4)
#define BB_SMEM_BYTES (2 * BK_BUF * 8 + 2 * BB_BCAP * 4 + 2 * BB_RP * 4 + 64)

struct BsrPlan {
    const int32_t* bcol;   // [nnz / (BS*BS)] block column (node) of every block, block rows in order
    int nbr;               // block rows per tile
    int lpr;               // lanes per block row (power of two, nbr * lpr == CTA size)
};

__device__ __forceinline__ void bb_init_barriers(unsigned char* smem) {
    double* const s_vals0 = reinterpret_cast<double*>(smem);
    int* const s_bc0 = reinterpret_cast<int*>(s_vals0 + 2 * BK_BUF);
    uint64_t* const bars = reinterpret_cast<uint64_t*>(s_bc0 + 2 * BB_BCAP + 2 * BB_RP);
    bk_mbar_init(&bars[0], 1);
    bk_mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// Calls epi(row, s) once for every row of this CTA's block rows (by the block row's lane 0).  All THREADS threads
// must call.  n = number of scalar rows (multiple of BS).  t_state as in bk_spmv_rows (NULL: barriers initialised here).
template <int BS, class Gather, class Epi, int THREADS>
__device__ __forceinline__ void bb_spmv_rows(const int32_t* __restrict__ rowptr, const double* __restrict__ vals,
                                             const BsrPlan& plan, int64_t n, const Gather& g, Epi&& epi, unsigned char* smem,
                                             uint32_t* t_state = nullptr) {
    double* const s_vals0 = reinterpret_cast<double*>(smem);                 // [2][BK_BUF]
    int* const s_bc0 = reinterpret_cast<int*>(s_vals0 + 2 * BK_BUF);         // [2][BB_BCAP]
    int* const s_rp0 = s_bc0 + 2 * BB_BCAP;                                  // [2][BB_RP]
    uint64_t* const bars = reinterpret_cast<uint64_t*>(s_rp0 + 2 * BB_RP);   // 8-byte aligned (all sizes above are multiples of 8)
#define sb_vals(i) (s_vals0 + (i) * BK_BUF)
#define sb_bc(i) (s_bc0 + (i) * BB_BCAP)
#define sb_rp(i) (s_rp0 + (i) * BB_RP)
    constexpr int B2 = BS * BS;
    const int tid = threadIdx.x;
    const int64_t nbrows = n / BS;
    const int NBR = plan.nbr, LPR = plan.lpr;
    const int64_t nnz = __ldg(&rowptr[n]);
    const int nnz_al = (int)(nnz & ~(int64_t)3);
    const int nblk_al = (int)((nnz / B2) & ~(int64_t)3);
    const int64_t ntile = (nbrows + NBR - 1) / NBR;
    int64_t tb = blockIdx.x;
    if (tb >= ntile) return;
    const int myrow = tid / LPR, lane = tid % LPR;

    auto issue = [&](uint32_t t, int kt, int kend) {  // thread 0: entries [kt, kend) of vals, blocks [kt/B2, kend/B2) of bcol
        const int st = t & 1;
        const int ka = kt & ~3;
        int kb = min((kend + 3) & ~3, nnz_al);
        if (kb < ka) kb = ka;
        const int ba = (kt / B2) & ~3;
        int bb = min((kend / B2 + 3) & ~3, nblk_al);
        if (bb < ba) bb = ba;
        const uint32_t cnt = (uint32_t)(kb - ka), bcnt = (uint32_t)(bb - ba);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bk_mbar_expect_tx(&bars[st], cnt * 8u + bcnt * 4u);
        if (cnt) bk_bulk_g2s(sb_vals(st), vals + ka, cnt * 8u, &bars[st]);
        if (bcnt) bk_bulk_g2s(sb_bc(st), plan.bcol + ba, bcnt * 4u, &bars[st]);
    };

    int64_t nd0 = tb * NBR;
    int nr = (int)min((int64_t)NBR, nbrows - nd0);
    for (int i = tid; i <= nr * BS; i += THREADS) sb_rp(0)[i] = __ldg(&rowptr[nd0 * BS + i]);
    if (tid == 0 && !t_state) {
        bk_mbar_init(&bars[0], 1);
        bk_mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int pb = 0;
    uint32_t t = t_state ? *t_state : 0u;
    if (tid == 0) issue(t, sb_rp(0)[0], sb_rp(0)[nr * BS]);

    while (true) {
        const int64_t tbn = tb + gridDim.x;
        const bool has_next = tbn < ntile;
        const int64_t nd0n = tbn * NBR;
        const int nrn = has_next ? (int)min((int64_t)NBR, nbrows - nd0n) : 0;
        // this tile's row offsets go to registers BEFORE the barrier below: the buffer sb_rp(pb) is overwritten by
        // the fastest threads at the top of the next iteration
        const int kt = sb_rp(pb)[0], kend = sb_rp(pb)[nr * BS];
        int seg[BS + 1];
#pragma unroll
        for (int i = 0; i <= BS; ++i) seg[i] = (myrow < nr) ? sb_rp(pb)[myrow * BS + i] : 0;
        if (has_next)
            for (int i = tid; i <= nrn * BS; i += THREADS) sb_rp(pb ^ 1)[i] = __ldg(&rowptr[nd0n * BS + i]);
        __syncthreads();  // stage (t+1)&1 is free again (every thread finished tile t-1); sb_rp[pb^1] is visible
        if (tid == 0 && has_next) issue(t + 1, sb_rp(pb ^ 1)[0], sb_rp(pb ^ 1)[nrn * BS]);
        const int st = t & 1;
        const int ka = kt & ~3, ba = (kt / B2) & ~3;
        bk_mbar_wait(&bars[st], (t >> 1) & 1);
        double* sv = sb_vals(st);
        int* sc = sb_bc(st);
        if (kend > nnz_al || kend / B2 > nblk_al) {  // matrix tail not covered by the 16-byte granular copies (uniform)
            const int kq = max(kt, nnz_al);
            if (tid < kend - kq) sv[kq + tid - ka] = __ldg(&vals[kq + tid]);
            const int bq = max(kt / B2, nblk_al);
            if (tid < kend / B2 - bq) sc[bq + tid - ba] = __ldg(&plan.bcol[bq + tid]);
            __syncthreads();
        }
        double y[BS];
#pragma unroll
        for (int i = 0; i < BS; ++i) y[i] = 0.0;
        if (myrow < nr) {
            const int nb = (seg[1] - seg[0]) / BS;
            const int bbase = seg[0] / B2 - ba;
            for (int j = lane; j < nb; j += LPR) {
                const int c = sc[bbase + j] * BS;
                double xv[BS];
#pragma unroll
                for (int k = 0; k < BS; ++k) xv[k] = g(c + k);
#pragma unroll
                for (int i = 0; i < BS; ++i) {
                    const double* v = sv + (seg[i] - ka) + BS * j;
#pragma unroll
                    for (int k = 0; k < BS; ++k) y[i] = fma(v[k], xv[k], y[i]);
                }
            }
        }
        for (int o = LPR >> 1; o > 0; o >>= 1) {
#pragma unroll
            for (int i = 0; i < BS; ++i) y[i] += __shfl_xor_sync(0xffffffffu, y[i], o);
        }
        if (myrow < nr && lane == 0) {
#pragma unroll
            for (int i = 0; i < BS; ++i) epi((nd0 + myrow) * BS + i, y[i]);
        }
        ++t;
        if (!has_next) break;
        pb ^= 1;
        tb = tbn;
        nd0 = nd0n;
        nr = nrn;
    }
    if (t_state) *t_state = t;
#undef sb_vals
#undef sb_bc
#undef sb_rp
}

// ---------------------------------------------------------------------------------------------------------------
// Direct node-block walk: no shared memory, no CTA barrier.  LPR lanes own one block row; its BS row segments are
// contiguous runs of L = BS * (blocks in the row) doubles, so lane l reads entries l, l + LPR, ... of EVERY segment
// (fully coalesced, each sector requested once) and all BS segments share one gathered x entry per position:
// column(idx) = BS * bcol[idx / BS] + idx % BS.  Up to 3 positions per lane are issued together (3 * BS value loads +
// 3 gathers in flight per thread, the only dependent chain is bcol -> x); the row offsets of the next block row are
// fetched one step ahead.  Block rows are dealt to the CTAs in contiguous chunks (DRAM page locality, and the load
// balance is one block row per LPR lanes instead of one tile per CTA).  Row sums by the same fixed xor-shuffle tree.
template <int BS, class Gather, class Epi, int THREADS>
__device__ __forceinline__ void bb_spmv_direct(const int32_t* __restrict__ rowptr, const double* __restrict__ vals,
                                               const BsrPlan& plan, int64_t n, const Gather& g, Epi&& epi) {
    constexpr int B2 = BS * BS;
    constexpr int M = 3;
    const int LPR = plan.lpr;
    const int sub = threadIdx.x / LPR, lane = threadIdx.x % LPR, nsub = THREADS / LPR;
    const int64_t nbrows = n / BS;
    const int64_t per = (nbrows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per;
    const int64_t r1 = min(nbrows, r0 + per);
    int segn[BS + 1];
    {
        const int64_t r = r0 + sub;
#pragma unroll
        for (int i = 0; i <= BS; ++i) segn[i] = (r < r1) ? __ldg(&rowptr[r * BS + i]) : 0;
    }
    for (int64_t base = r0; base < r1; base += nsub) {  // uniform trip count in the CTA: full-mask shuffles are safe
        const int64_t r = base + sub;
        int seg[BS + 1];
#pragma unroll
        for (int i = 0; i <= BS; ++i) seg[i] = segn[i];
        {
            const int64_t rn = r + nsub;
#pragma unroll
            for (int i = 0; i <= BS; ++i) segn[i] = (rn < r1) ? __ldg(&rowptr[rn * BS + i]) : 0;
        }
        double y[BS];
#pragma unroll
        for (int i = 0; i < BS; ++i) y[i] = 0.0;
        if (r < r1) {
            const int L = seg[1] - seg[0];
            const int bbase = seg[0] / B2;
            int cc[M];
            double vv[M][BS];
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const int idx = lane + m * LPR;
                const bool ok = idx < L;
                cc[m] = ok ? __ldg(&plan.bcol[bbase + idx / BS]) * BS + idx % BS : -1;
#pragma unroll
                for (int i = 0; i < BS; ++i) vv[m][i] = ok ? __ldcs(vals + seg[i] + idx) : 0.0;
            }
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const double xv = cc[m] >= 0 ? g(cc[m]) : 0.0;
#pragma unroll
                for (int i = 0; i < BS; ++i) y[i] = fma(vv[m][i], xv, y[i]);
            }
            for (int idx = lane + M * LPR; idx < L; idx += LPR) {  // block rows longer than 3 * LPR / BS blocks
                const double xv = g(__ldg(&plan.bcol[bbase + idx / BS]) * BS + idx % BS);
#pragma unroll
                for (int i = 0; i < BS; ++i) y[i] = fma(__ldcs(vals + seg[i] + idx), xv, y[i]);
            }
        }
        for (int o = LPR >> 1; o > 0; o >>= 1) {
#pragma unroll
            for (int i = 0; i < BS; ++i) y[i] += __shfl_xor_sync(0xffffffffu, y[i], o);
        }
        if (r < r1 && lane == 0) {
#pragma unroll
            for (int i = 0; i < BS; ++i) epi(r * BS + i, y[i]);
        }
    }
}
