// Internal helpers shared by the libpgdb200 translation units (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/pgd_b200.h"

#define PGD_MAX_PARTIALS (1 << 20)
#define PGD_MAX_COUNTERS 4096
#define PGD_STREAM_MIN_ROWS 8192
#define PGD_BULK_MIN_ROWS 32768

struct pgd_ctx {
    int device;
    int sm_count;
    char err[512];
    double* partials;        // [PGD_MAX_PARTIALS] scratch for two-stage reductions
    unsigned int* counters;  // [PGD_MAX_COUNTERS] "last block done" tickets, always left at 0
    double* scalars;         // [64] PCG scalars
    int* flags;              // [16] PCG flags / iteration counter
    // pending sparsity pattern (between pattern_build_sync and pattern_export)
    int32_t* pat_rowptr;
    int32_t* pat_colidx;
    int64_t* pat_gptr;
    int32_t* pat_gidx;
    int64_t pat_nnz, pat_ndofs, pat_ncontrib;
    // statistics (pgd_get_stats): kernel launches, PCG solves / iterations / device time
    int64_t n_launches, pcg_solves, pcg_iters, pcg_resident_solves;
    double pcg_ms;
    void* pinned;            // 512 B of page-locked host memory: [0,256) staging for the small device->host reads of the *_sync
                             // calls, [256,512) results of a started (pgd_pcg_start) solve
    int res_pending;         // a solve started with pgd_pcg_start / pgd_pcg_persist_start has not been finished yet
    int res_kind;            // ... 1 = SM-resident kernel, 2 = persistent streaming kernel (status codes differ)
    void* res_stream;        // ... on this stream
    const void* fit_key;     // rowptr / n / block of the last system the SM-resident PCG solved (=> it fits: eligible for start)
    int64_t fit_n;
    int fit_block;
    void* arena;             // grow-only device scratch of the set-up calls (pattern / vecmap builds), see pgd_arena
    size_t arena_cap;
    int pat_in_arena;        // the pending pattern lives in the arena (nothing to free)
    int opt_resident;        // pgd_set_option("pcg_resident"): 1 = use the SM-resident PCG when the system fits
    int opt_stream;          // pgd_set_option("spmv_stream"): 0 = sub-warp-per-row kernels only; 1 = register-staged row-block
                             // streaming for n >= PGD_STREAM_MIN_ROWS; 2 (default) = additionally the TMA-pipelined kernel
                             // for n >= PGD_BULK_MIN_ROWS
    const void* nnz_key;     // cache of rowptr[n] (one 4-byte D2H per new matrix)
    int64_t nnz_key_n, nnz_val;
    cudaEvent_t ev0, ev1;
    cudaEvent_t ev_done;     // recorded behind the result copies of a started solve: pgd_pcg_finish waits for it only, not for
                             // whatever else has been enqueued on the stream since
    void* comm;              // ncclComm_t of the sharded solves (pgd_comm_init), NULL = single rank
    int comm_rank, comm_world;
    // NVLink peer window of the sharded solves (pgd_peer_window_*): this rank's cudaMalloc'ed window, the
    // IPC-mapped windows of the other ranks, and the layout (identical on every rank)
    void* win_local;
    void* win_peer[16];
    int64_t win_pcap;        // doubles reserved for p = [owned | ghost]
    int win_world, win_rank;
    unsigned long long win_ar_seq, win_halo_seq;
    double* p_override;      // when set, the sharded-PCG blocks keep p here (inside the window) instead of in d_work
    int opt_p2p;             // pgd_set_option("p2p"): 1 (default) = use the peer window when it exists, 0 = NCCL
    void* cap_stream;        // private stream used to capture the peer-window iteration into a CUDA graph
    int opt_pcg3;            // pgd_set_option("pcg3"): 1 (default) = 3-kernel PCG iteration in the HBM-bound regime
    int opt_fused;           // pgd_set_option("fused"): 1 = peer-window iteration with the collectives fused into the SpMV /
                             // update kernels (4 launches); 0 (default) = separate push / wait / all-reduce kernels (8
                             // launches, CUDA-graph replayed) -- measured 7 % faster at 2 GPUs: the fused SpMV has to gather p
                             // with coherent loads instead of ld.global.nc
    int opt_graph;           // pgd_set_option("graph"): 1 (default) = replay the peer-window iteration from a CUDA graph
    int opt_persist;         // pgd_set_option("persist"): 1 (default) = HBM-bound solves run in the persistent cooperative kernel
    int opt_bsr;             // pgd_set_option("bsr"): node-block walk of vector operators inside that kernel: 2 (default) = direct (no shared memory), 1 = tiles through the TMA ring, 0 = off
    int opt_ll;              // pgd_set_option("ll"): 1 (default) = LL mailboxes / halo in the persistent sharded kernel (data + sequence per 8-byte word), 0 = data, fence, flag
    int opt_spin_ms;         // pgd_set_option("spin_ms"): budget of every in-kernel wait (default 20 000 ms)
    int opt_prof;            // pgd_set_option("prof"): 1 = the persistent kernel accumulates per-phase times (pgd_get_phase_ns)
    int opt_single_reduction;  // pgd_set_option("single_reduction"): 0 never, 1 sharded solves, 2 always (pcg_persist.cu)
    unsigned int attr_mask;  // which kernels already carry their dynamic shared-memory opt-in on this device
    void* mailbox;           // single-GPU stand-in for the peer window's mailboxes (PwLayout{0}: slots + flags, 2 KB)
};

void pgd_free_pattern(pgd_ctx* h);

#define PGD_CHECK_HANDLE(h) \
    if (!(h)) return -1;

#define PGD_ARG(h, cond, msg)                                      \
    do {                                                           \
        if (!(cond)) {                                             \
            snprintf((h)->err, sizeof((h)->err), "%s: %s", __func__, msg); \
            return -2;                                             \
        }                                                          \
    } while (0)

#define PGD_CUDA(h, expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            snprintf((h)->err, sizeof((h)->err), "%s: %s -> %s", __func__, #expr, cudaGetErrorString(_e)); \
            return (int32_t)_e;                                                             \
        }                                                                                   \
    } while (0)

#define PGD_LAUNCH_OK(h)                 \
    do {                                 \
        (h)->n_launches += 1;            \
        PGD_CUDA(h, cudaGetLastError()); \
    } while (0)

// Device -> host read of up to two small items (<= 128 B each) followed by a stream synchronise, staged through the
// handle's page-locked buffer: a pageable destination sends cudaMemcpyAsync down the driver's blocking staging path,
// which showed up as sporadic 100+ ms stalls of the *_sync entry points on a busy host.
static inline cudaError_t pgd_fetch(pgd_ctx* h, void* dst0, const void* src0, size_t n0, void* dst1, const void* src1,
                                    size_t n1, cudaStream_t st) {
    char* pin = static_cast<char*>(h->pinned);
    cudaError_t e = cudaMemcpyAsync(pin, src0, n0, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && n1) e = cudaMemcpyAsync(pin + 128, src1, n1, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    memcpy(dst0, pin, n0);
    if (n1) memcpy(dst1, pin + 128, n1);
    return cudaSuccess;
}

// Scratch for the set-up calls.  Requests up to PGD_ARENA_KEEP bytes are served from one grow-only block that stays
// with the handle (no cudaMalloc / cudaFree per call: both are blocking driver calls); larger ones (the 4 M-dof
// patterns need GBs for a moment) are allocated for the call and must be returned with pgd_arena_release.
#define PGD_ARENA_KEEP ((size_t)512 << 20)
static inline cudaError_t pgd_arena(pgd_ctx* h, size_t bytes, void** out, bool* temporary) {
    *temporary = false;
    if (bytes > PGD_ARENA_KEEP) {
        *temporary = true;
        return cudaMalloc(out, bytes);
    }
    if (bytes > h->arena_cap) {
        if (h->arena) cudaFree(h->arena);
        h->arena = nullptr;
        h->arena_cap = 0;
        size_t cap = bytes + bytes / 4;
        if (cap > PGD_ARENA_KEEP) cap = PGD_ARENA_KEEP;
        cudaError_t e = cudaMalloc(&h->arena, cap);
        if (e != cudaSuccess) return e;
        h->arena_cap = cap;
    }
    *out = h->arena;
    return cudaSuccess;
}

static inline int pgd_set_device(pgd_ctx* h) { return (int)cudaSetDevice(h->device); }

static inline unsigned int pgd_blocks(int64_t n, int per_block) { return (unsigned int)((n + per_block - 1) / per_block); }

// ----------------------------------------------------------------------------- device helpers
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the block; result valid in thread 0. blockDim.x multiple of 32, <= 1024.
__device__ __forceinline__ double block_sum(double v) {
    __shared__ double sh[32];
    __syncthreads();  // protect sh across consecutive calls
    v = warp_sum(v);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
    if (w == 0) v = warp_sum(v);
    return v;
}

// Deterministic grid reduction of NV values per block: every block deposits its partials, the
// last block to arrive sums them in a fixed order and stores out[0..NV). `counter` returns to 0.
// part layout: part[v * gridsize + block]. Works for 1-D grids (gridsize = gridDim.x) or a row
// of a 2-D grid when the caller passes its own base pointers.
// Returns true (to every thread) in the block that performed the final sum.
template <int NV>
__device__ __forceinline__ bool grid_sum_finish(const double (&v)[NV], double* part, unsigned int* counter,
                                                double* out, unsigned int bid, unsigned int gridsize) {
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) part[(size_t)k * gridsize + bid] = v[k];
        __threadfence();
        unsigned int t = atomicAdd(counter, 1u);
        s_last = (t == gridsize - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double s = 0.0;
            for (unsigned int i = threadIdx.x; i < gridsize; i += blockDim.x) s += __ldcg(&part[(size_t)k * gridsize + i]);
            s = block_sum(s);
            if (threadIdx.x == 0) out[k] = s;
        }
        if (threadIdx.x == 0) *counter = 0u;
    }
    return s_last;
}

// streaming loads for data that is touched once per kernel (matrix values / column indices)
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }
__device__ __forceinline__ int ld_stream(const int* p) { return __ldcs(p); }
