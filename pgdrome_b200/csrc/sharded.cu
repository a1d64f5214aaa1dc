// Sharded Jacobi-PCG solve over the GPUs of one box: the whole iteration loop in one host call.
//
// One process per GPU.  Every rank owns a contiguous block of rows (local CSR, vectors laid out
// [owned | ghost], see pgd_b200.h "sharded PCG building blocks").  Per iteration the host enqueues, on
// ONE stream and without ever synchronising inside a batch of `check_every` iterations:
//     k_spcg_direction   p = z + beta p
//     k_halo_pack        send buffer = p[send_idx]
//     ncclGroup{ ncclSend / ncclRecv per neighbour }      ghosts land directly in p's tail
//     k_spmv_* (+dot)    q = A_loc p, local p.q
//     ncclAllReduce      p.q                               (1 double, in place in the device scalars)
//     k_spcg_update      x, r, z, local r.z and r.r
//     ncclAllReduce      r.z, r.r                          (2 doubles)
//     k_spcg_rotate      scalars, iteration counter, convergence flag (identical on every rank)
// NCCL is bound at run time (dlopen of libnccl.so.2: inside a PyTorch process that is the copy torch
// already loaded), so the library itself has no link-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"
#include "pcg_blocks.cuh"
#include "spmv_bulk.cuh"

struct NcclApi {
    void* lib;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
};
static NcclApi g_nccl = {};

static const char* nccl_load() {
    if (g_nccl.lib) return nullptr;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return "libnccl.so.2 not found";
#define NCCL_SYM(field, name)                                   \
    *(void**)(&g_nccl.field) = dlsym(lib, name);                \
    if (!g_nccl.field) return "missing NCCL symbol " name;
    NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    NCCL_SYM(CommInitRank, "ncclCommInitRank")
    NCCL_SYM(CommDestroy, "ncclCommDestroy")
    NCCL_SYM(AllReduce, "ncclAllReduce")
    NCCL_SYM(Send, "ncclSend")
    NCCL_SYM(Recv, "ncclRecv")
    NCCL_SYM(GroupStart, "ncclGroupStart")
    NCCL_SYM(GroupEnd, "ncclGroupEnd")
    NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef NCCL_SYM
    g_nccl.lib = lib;
    return nullptr;
}

#define PGD_NCCL(h, expr)                                                                               \
    do {                                                                                                \
        ncclResult_t _r = (expr);                                                                       \
        if (_r != ncclSuccess) {                                                                        \
            snprintf((h)->err, sizeof((h)->err), "%s: %s -> %s", __func__, #expr, g_nccl.GetErrorString(_r)); \
            return -5;                                                                                  \
        }                                                                                               \
    } while (0)

extern "C" int32_t pgd_comm_unique_id(void* h_id128) {
    if (!h_id128) return -2;
    if (nccl_load()) return -5;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return -5;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(h_id128, &id, 128);
    return 0;
}

extern "C" int32_t pgd_comm_init(pgd_handle_t h, const void* h_id128, int32_t rank, int32_t world) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, h_id128 && world >= 1 && rank >= 0 && rank < world, "bad arguments");
    const char* e = nccl_load();
    if (e) {
        snprintf(h->err, sizeof(h->err), "pgd_comm_init: %s", e);
        return -5;
    }
    if (h->comm) {
        g_nccl.CommDestroy((ncclComm_t)h->comm);
        h->comm = nullptr;
    }
    pgd_set_device(h);
    ncclUniqueId id;
    memcpy(&id, h_id128, 128);
    ncclComm_t comm;
    PGD_NCCL(h, g_nccl.CommInitRank(&comm, world, id, rank));
    h->comm = comm;
    h->comm_rank = rank;
    h->comm_world = world;
    return 0;
}

extern "C" int32_t pgd_comm_destroy(pgd_handle_t h) {
    PGD_CHECK_HANDLE(h);
    if (h->comm && g_nccl.lib) g_nccl.CommDestroy((ncclComm_t)h->comm);
    h->comm = nullptr;
    h->comm_world = 0;
    return 0;
}

__global__ void __launch_bounds__(256) k_halo_pack(const double* __restrict__ v, const int64_t* __restrict__ idx, int64_t n,
                                                   double* __restrict__ buf) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) buf[i] = v[idx[i]];
}

// ================================================================================================
// NVLink peer window: halo exchange and dot-product all-reduce WITHOUT NCCL in the iteration.
//
// Every rank cudaMalloc's one window and maps the windows of the other ranks of the box with CUDA IPC
// (NVSwitch: every peer at full NVLink bandwidth).  Layout, identical on every rank (win_pcap doubles
// for p, then the small mailboxes):
//     p[win_pcap]                      this rank's direction vector [owned | ghost]; neighbours STORE their
//                                      boundary values straight into the ghost tail (k_halo_push)
//     ar_slot[2][16][4] doubles        all-reduce mailboxes, double-buffered by sequence parity
//     ar_flag[2][16], halo_flag[16]    u64 sequence numbers written with st.release.sys by the sender
// k_halo_push     after p = z + beta p: gathers the boundary entries and stores them into the owners' ghost
//                 slots over NVLink; the last CTA publishes halo_flag[me] = seq on every destination.
// k_halo_wait     one warp spins (ld.acquire.sys) until every source has published seq.
// k_allreduce_p2p one-shot all-reduce of <= 4 doubles: store my partials into everybody's mailbox, flag,
//                 wait for everybody's flag, sum in RANK ORDER (=> bitwise identical on all ranks).
// All spins carry a clock64() budget; on expiry they raise the NaN/bad flag instead of hanging the GPU.
// ================================================================================================
#include "peer_window.cuh"

// wall-clock budget of every spin of this file (pgd_set_option "spin_ms"), refreshed at the start of each solve
__device__ unsigned long long g_pw_budget_ns = 20000000000ULL;
__device__ __forceinline__ unsigned long long pw_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(256) k_halo_push(const double* __restrict__ p, const int64_t* __restrict__ send_idx,
                                                   int64_t n_send, PwPeers peers, PwHalo hp, PwLayout lay, int me, int world,
                                                   const unsigned long long* seq_ctr, unsigned int* counter, const int* fl) {
    if (fl[0]) return;
    const unsigned long long seq = *seq_ctr + 1ULL;  // advanced by k_halo_wait (sequence numbers live on the device so
                                                     // that the iteration can be replayed from a CUDA graph)
    __shared__ bool s_last;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_send; s += stride) {
        int r = 0;
        while (s >= hp.seg_start[r + 1]) ++r;
        double* dst = reinterpret_cast<double*>(peers.base[r]) + hp.dst_off[r] + (s - hp.seg_start[r]);
        *dst = p[send_idx[s]];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(counter, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if ((int)threadIdx.x < world && (int)threadIdx.x != me && hp.seg_start[threadIdx.x + 1] > hp.seg_start[threadIdx.x]) {
            unsigned long long* f = reinterpret_cast<unsigned long long*>(peers.base[threadIdx.x] + lay.haloflag_off()) + me;
            st_release_sys(f, seq);
        }
        if (threadIdx.x == 0) *counter = 0u;
    }
}

__global__ void k_halo_wait(unsigned char* mine, PwHalo hp, PwLayout lay, int world, unsigned long long* seq_ctr, int* fl) {
    if (fl[0]) return;
    const unsigned long long seq = *seq_ctr + 1ULL;
    const int r = threadIdx.x;
    if (r < world && hp.recv_from[r]) {
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(mine + lay.haloflag_off()) + r;
        const unsigned long long t0 = pw_now();
        while (ld_acquire_sys(f) < seq) {
            if (pw_now() - t0 > g_pw_budget_ns) {
                fl[2] = 2;  // peer did not arrive: report instead of hanging
                fl[0] = 1;
                break;
            }
        }
    }
    __syncwarp();
    if (r == 0) *seq_ctr = seq;
}

// One-shot mailbox all-reduce executed by ONE CTA (threads 0..world-1 talk to one peer each):
// vals: nv (<= 4) doubles, reduced in place; every rank sums the mailboxes in rank order.
__device__ __forceinline__ void ar_p2p_block(double* vals, int nv, const PwPeers& peers, const PwLayout& lay, int me,
                                             int world, unsigned long long* seq_ctr, int* fl) {
    __shared__ int s_bad;
    const unsigned long long seq = *seq_ctr + 1ULL;
    const int par = (int)(seq & 1ULL);
    const int r = threadIdx.x;
    if (r == 0) s_bad = 0;
    __syncthreads();
    if (r < world) {
        double* slot = reinterpret_cast<double*>(peers.base[r] + lay.slot_off()) + ((size_t)par * PW_MAXR + me) * PW_AR_VALS;
        for (int k = 0; k < nv; ++k) slot[k] = *((volatile double*)&vals[k]);
        __threadfence_system();
        st_release_sys(reinterpret_cast<unsigned long long*>(peers.base[r] + lay.arflag_off()) + par * PW_MAXR + me, seq);
        const unsigned long long* f =
            reinterpret_cast<const unsigned long long*>(peers.base[me] + lay.arflag_off()) + par * PW_MAXR + r;
        const unsigned long long t0 = pw_now();
        while (ld_acquire_sys(f) < seq) {
            if (pw_now() - t0 > g_pw_budget_ns) {
                s_bad = 1;
                break;
            }
        }
    }
    __syncthreads();
    if (r == 0) {
        if (s_bad) {
            fl[2] = 2;
            fl[0] = 1;
        } else {
            const double* mine = reinterpret_cast<const double*>(peers.base[me] + lay.slot_off()) + (size_t)par * PW_MAXR * PW_AR_VALS;
            for (int k = 0; k < nv; ++k) {
                double sum = 0.0;
                for (int q = 0; q < world; ++q) sum += *((volatile const double*)&mine[q * PW_AR_VALS + k]);
                vals[k] = sum;
            }
        }
        *seq_ctr = seq;
    }
    __syncthreads();
}

__global__ void k_allreduce_p2p(double* vals, int nv, PwPeers peers, PwLayout lay, int me, int world,
                                unsigned long long* seq_ctr, int* fl, int skip_when_done) {
    if (skip_when_done && fl[0]) return;
    ar_p2p_block(vals, nv, peers, lay, me, world, seq_ctr, fl);
}

// ------------------------------------------------------------------------------------------------
// Fused iteration of the peer-window path: 4 launches, the collectives inside the compute kernels.
//   k_halo_push_pre        boundary entries of p = z + beta p_old stored into the neighbours' ghost slots over
//                          NVLink (+ flag), then the plain direction kernel updates p in place
//   k_spcg_matvec_p2p      every CTA first acquires the neighbours' halo flags, then the TMA-pipelined
//                          SpMV q = A_loc p with the local p.q; the CTA that finishes the grid reduction
//                          runs the mailbox all-reduce of p.q with the other GPUs
//   k_spcg_update_p2p      x, r, z update with local r.z, r.r; the finishing CTA all-reduces both and rotates
//                          the CG scalars / iteration counter / convergence flag
// ------------------------------------------------------------------------------------------------

// boundary entries of the NEW direction p = z + beta p_old, formed from z and p_old (the same fma the direction
// kernel evaluates right afterwards => bitwise identical) and stored into the neighbours' ghost slots
__global__ void __launch_bounds__(256) k_halo_push_pre(const double* __restrict__ z, const double* __restrict__ p_old,
                                                       const double* sc, const int* fl,
                                                       const int64_t* __restrict__ send_idx, int64_t n_send, PwPeers peers,
                                                       PwHalo hp, PwLayout lay, int me, int world,
                                                       const unsigned long long* halo_ctr, unsigned int* counter) {
    if (fl[F_DONE]) return;
    __shared__ bool s_last;
    const double beta = (fl[F_ITER] == 0) ? 0.0 : sc[S_RZ_NEW] / sc[S_RZ_OLD];
    const unsigned long long seq = *halo_ctr + 1ULL;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_send; s += stride) {
        int r = 0;
        while (s >= hp.seg_start[r + 1]) ++r;
        const int64_t i = send_idx[s];
        double* dst = reinterpret_cast<double*>(peers.base[r]) + hp.dst_off[r] + (s - hp.seg_start[r]);
        *dst = fma(beta, p_old[i], z[i]);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(counter, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if ((int)threadIdx.x < world && (int)threadIdx.x != me && hp.seg_start[threadIdx.x + 1] > hp.seg_start[threadIdx.x]) {
        unsigned long long* f = reinterpret_cast<unsigned long long*>(peers.base[threadIdx.x] + lay.haloflag_off()) + me;
        st_release_sys(f, seq);
    }
    if (threadIdx.x == 0) *counter = 0u;
}

__global__ void __launch_bounds__(BK_THREADS, BK_CTAS_PER_SM) k_spcg_matvec_p2p(const int32_t* __restrict__ rowptr,
                                                                   const int32_t* __restrict__ colidx,
                                                                   const double* __restrict__ vals, const double* p,
                                                                   double* __restrict__ q, int64_t n, double* sc, int* fl,
                                                                   double* part, unsigned int* counter, PwPeers peers,
                                                                   PwHalo hp, PwLayout lay, int me, int world,
                                                                   unsigned long long* halo_ctr, unsigned long long* ar_ctr) {
    extern __shared__ __align__(128) unsigned char bk_smem[];
    if (fl[F_DONE]) return;
    // ---- acquire the halo of p (the neighbours stored it into this rank's ghost slots)
    const unsigned long long hseq = *halo_ctr + 1ULL;
    if ((int)threadIdx.x < world && hp.recv_from[threadIdx.x]) {
        const unsigned long long* f =
            reinterpret_cast<const unsigned long long*>(peers.base[me] + lay.haloflag_off()) + threadIdx.x;
        const unsigned long long t0 = pw_now();
        while (ld_acquire_sys(f) < hseq) {
            if (pw_now() - t0 > g_pw_budget_ns) {
                fl[F_BAD] = 2;
                break;
            }
        }
    }
    __syncthreads();
    // ---- q = A_loc p (p is read through L2: its ghost tail was written by peers), local p.q
    double acc = 0.0;
    struct GatherCG {
        const double* x;
        // plain (L1-cached, not .nc) loads: peers write p's ghost tail while this kernel is resident, but no line
        // of p is touched by this SM before the acquire above
        __device__ __forceinline__ double operator()(int c) const { return x[c]; }
    };
    bk_spmv_rows(rowptr, colidx, vals, n, GatherCG{p},
                 [&](int64_t row, double s) {
                     q[row] = s;
                     acc = fma(p[row], s, acc);
                 },
                 bk_smem);
    acc = block_sum(acc);
    double v[1] = {acc};
    const bool last = grid_sum_finish<1>(v, part, counter, sc + S_PQ, blockIdx.x, gridDim.x);
    if (!last) return;
    __syncthreads();
    if (threadIdx.x == 0) *halo_ctr = hseq;  // every CTA has passed its wait (it arrived at the reduction counter)
    ar_p2p_block(sc + S_PQ, 1, peers, lay, me, world, ar_ctr, fl);
    if (threadIdx.x == 0 && fl[F_BAD] == 2) fl[F_DONE] = 1;
}

template <int BS>
__global__ void __launch_bounds__(256) k_spcg_update_p2p(double* __restrict__ x, double* __restrict__ r, double* __restrict__ z,
                                                         const double* p, const double* __restrict__ q,
                                                         const double* __restrict__ minv, int64_t n_nodes, double* sc, int* fl,
                                                         double* part, unsigned int* counter, PwPeers peers, PwLayout lay,
                                                         int me, int world, unsigned long long* ar_ctr) {
    if (fl[F_DONE]) return;
    const double alpha = sc[S_RZ_NEW] / sc[S_PQ];
    const double rz_cur = sc[S_RZ_NEW];
    const int it = fl[F_ITER];
    double rz = 0.0, rr = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t nd = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; nd < n_nodes; nd += stride) {
        double rn[BS];
#pragma unroll
        for (int i = 0; i < BS; ++i) {
            const int64_t d = nd * BS + i;
            x[d] = fma(alpha, p[d], x[d]);
            rn[i] = fma(-alpha, q[d], r[d]);
            r[d] = rn[i];
            rr = fma(rn[i], rn[i], rr);
        }
#pragma unroll
        for (int i = 0; i < BS; ++i) {
            double zi = 0.0;
#pragma unroll
            for (int k = 0; k < BS; ++k) zi = fma(__ldg(&minv[(nd * BS + i) * BS + k]), rn[k], zi);
            z[nd * BS + i] = zi;
            rz = fma(rn[i], zi, rz);
        }
    }
    rz = block_sum(rz);
    rr = block_sum(rr);
    double v[2] = {rz, rr};
    const bool last = grid_sum_finish<2>(v, part, counter, sc + S_TMP, blockIdx.x, gridDim.x);
    if (!last) return;
    __syncthreads();
    ar_p2p_block(sc + S_TMP, 2, peers, lay, me, world, ar_ctr, fl);
    if (threadIdx.x == 0 && !fl[F_DONE]) {  // scalar rotation (identical on every rank)
        const double rz_next = sc[S_TMP], rr_new = sc[S_TMP + 1];
        sc[S_RZ_OLD] = rz_cur;
        sc[S_RZ_NEW] = rz_next;
        sc[S_RR] = rr_new;
        fl[F_ITER] = it + 1;
        if (!(rr_new > sc[S_TOL2])) fl[F_DONE] = 1;
        if (!(rr_new == rr_new) || !(rz_next == rz_next)) fl[F_BAD] = 1;
    }
}

extern "C" int32_t pgd_peer_window_destroy(pgd_handle_t h);

extern "C" int32_t pgd_peer_window_create(pgd_handle_t h, int64_t p_capacity, void* h_ipc64) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, p_capacity > 0 && h_ipc64, "bad arguments");
    pgd_set_device(h);
    if (h->win_local) pgd_peer_window_destroy(h);  // closes the peer mappings of the old window before it is freed (the
                                                   // caller puts a barrier across the ranks in front: peers still map it)
    PwLayout lay{p_capacity};
    void* p = nullptr;
    PGD_CUDA(h, cudaMalloc(&p, lay.bytes()));
    PGD_CUDA(h, cudaMemset(p, 0, lay.bytes()));
    cudaIpcMemHandle_t ih;
    PGD_CUDA(h, cudaIpcGetMemHandle(&ih, p));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(h_ipc64, &ih, 64);
    h->win_local = p;
    h->win_pcap = p_capacity;
    h->win_world = 0;
    return 0;
}

extern "C" int32_t pgd_peer_window_open(pgd_handle_t h, int32_t rank, int32_t world, const void* h_all_ipc) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, h->win_local && h_all_ipc && world >= 1 && world <= PW_MAXR && rank >= 0 && rank < world, "bad arguments");
    pgd_set_device(h);
    for (int r = 0; r < h->win_world; ++r)  // re-open: drop the old mappings first
        if (r != h->win_rank && h->win_peer[r]) cudaIpcCloseMemHandle(h->win_peer[r]);
    h->win_world = 0;
    {  // sequence numbers restart at 0 below: so must the flags, the mailboxes and the LL halo words of this rank's window
        PwLayout lay0{h->win_pcap};
        PGD_CUDA(h, cudaMemset(h->win_local, 0, lay0.bytes()));
    }
    for (int r = 0; r < world; ++r) {
        if (r == rank) {
            h->win_peer[r] = h->win_local;
            continue;
        }
        cudaIpcMemHandle_t ih;
        memcpy(&ih, (const char*)h_all_ipc + 64 * r, 64);
        void* q = nullptr;
        PGD_CUDA(h, cudaIpcOpenMemHandle(&q, ih, cudaIpcMemLazyEnablePeerAccess));
        h->win_peer[r] = q;
    }
    h->win_world = world;
    h->win_rank = rank;
    h->win_ar_seq = 0;
    h->win_halo_seq = 0;
    PGD_CUDA(h, cudaMemset(h->scalars + 40, 0, 2 * sizeof(double)));  // device-side sequence counters (halo, all-reduce)
    PGD_CUDA(h, cudaDeviceSynchronize());
    return 0;
}

extern "C" int32_t pgd_peer_window_destroy(pgd_handle_t h) {
    PGD_CHECK_HANDLE(h);
    pgd_set_device(h);
    for (int r = 0; r < h->win_world; ++r)
        if (r != h->win_rank && h->win_peer[r]) cudaIpcCloseMemHandle(h->win_peer[r]);
    if (h->win_local) cudaFree(h->win_local);
    h->win_local = nullptr;
    h->win_world = 0;
    for (int r = 0; r < PW_MAXR; ++r) h->win_peer[r] = nullptr;
    h->p_override = nullptr;
    return 0;
}

// building blocks defined in pcg.cu
extern "C" int32_t pgd_spcg_init(pgd_handle_t, const int32_t*, const int32_t*, const double*, const double*, double*, int64_t,
                                 int64_t, int32_t, double*, double*, int32_t*, void*);

extern "C" int32_t pgd_spcg_solve_sync(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                                       const double* d_b, double* d_x, int64_t n_owned, int64_t n_local, int32_t block,
                                       const int64_t* d_send_idx, const int64_t* h_send_counts, const int64_t* h_recv_counts,
                                       double rtol, double atol, int32_t maxit, int32_t check_every, double* d_work,
                                       int32_t* h_iters, double* h_relres, void* stream, const int64_t* h_peer_ghost_base) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_rowptr && d_colidx && d_values && d_b && d_x && d_work, "null pointer");
    PGD_ARG(h, n_owned >= 0 && n_local >= n_owned && block >= 1 && block <= 3, "bad sizes");
    const bool p2p = h->opt_p2p && h->win_local && h->win_world > 1 && h_peer_ghost_base && n_local <= h->win_pcap;
    const int world = p2p ? h->win_world : (h->comm ? h->comm_world : 1);
    PGD_ARG(h, world == 1 || (h_send_counts && h_recv_counts), "split sizes required with more than one rank");
    cudaStream_t st = (cudaStream_t)stream;
    ncclComm_t comm = (ncclComm_t)h->comm;
    int64_t n_send = 0;
    if (world > 1)
        for (int r = 0; r < world; ++r) n_send += h_send_counts[r];
    PGD_ARG(h, n_send == 0 || d_send_idx, "send index list required");
    // work: r z q minv p[n_local] sendbuf[n_send]   (p2p: p lives in the peer window instead); even strides
    const int64_t ns = (n_owned + 1) & ~(int64_t)1;
    const int64_t p_off = 3 * ns + ((n_owned * block + 1) & ~(int64_t)1);
    h->p_override = p2p ? reinterpret_cast<double*>(h->win_local) : nullptr;
    double* p = p2p ? h->p_override : d_work + p_off;
    double* sendbuf = d_work + p_off + ((n_local + 1) & ~(int64_t)1);
    double* sc = h->scalars;
    int* fl = h->flags;
    PwPeers peers;
    PwHalo hp;
    PwLayout lay{h->win_pcap};
    const int me = p2p ? h->win_rank : 0;
    unsigned long long* seq_halo = reinterpret_cast<unsigned long long*>(h->scalars + 40);
    unsigned long long* seq_ar = reinterpret_cast<unsigned long long*>(h->scalars + 41);
    if (p2p) {
        for (int r = 0; r < PW_MAXR; ++r) peers.base[r] = (unsigned char*)(r < world ? h->win_peer[r] : nullptr);
        hp.seg_start[0] = 0;
        for (int r = 0; r < PW_MAXR; ++r) {
            hp.seg_start[r + 1] = hp.seg_start[r] + (r < world ? h_send_counts[r] : 0);
            hp.dst_off[r] = r < world ? h_peer_ghost_base[r] : 0;
            hp.recv_from[r] = (r < world && h_recv_counts[r] > 0) ? 1 : 0;
        }
    }
    auto allreduce = [&](double* vals, int nv, int skip_when_done, cudaStream_t st) -> int32_t {
        if (world <= 1) return 0;
        if (p2p) {
            k_allreduce_p2p<<<1, 32, 0, st>>>(vals, nv, peers, lay, me, world, seq_ar, fl, skip_when_done);
            h->n_launches += 1;
            return (int32_t)cudaGetLastError();
        }
        PGD_NCCL(h, g_nccl.AllReduce(vals, vals, (size_t)nv, ncclDouble, ncclSum, comm, st));
        return 0;
    };
    if (p2p) {
        const unsigned long long budget = (unsigned long long)h->opt_spin_ms * 1000000ULL;
        PGD_CUDA(h, cudaMemcpyToSymbolAsync(g_pw_budget_ns, &budget, sizeof(budget), 0, cudaMemcpyHostToDevice, st));
    }
    int32_t rc = pgd_spcg_init(h, d_rowptr, d_colidx, d_values, d_b, d_x, n_owned, n_local, block, d_work, sc, fl, stream);
    if (rc) return rc;
    if ((rc = allreduce(sc + 8, 3, 0, st))) return rc;
    rc = pgd_spcg_init_fin(h, sc, fl, rtol, atol, stream);
    if (rc) return rc;
    if (check_every < 1) check_every = 1;
    int hf[4] = {0, 0, 0, 0};
    int launched = 0;
    PGD_CUDA(h, cudaEventRecord(h->ev0, st));
    const bool fused = p2p && h->opt_fused;
    if (fused) PGD_CUDA(h, cudaFuncSetAttribute(k_spcg_matvec_p2p, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM_BYTES));
    auto enqueue_iteration = [&](cudaStream_t st) -> int32_t {
        void* stream = (void*)st;
        int32_t rc2;
        if (fused) {
            double* w_r = d_work;
            double* w_z = w_r + ns;
            double* w_q = w_z + ns;
            double* w_minv = w_q + ns;
            if (n_send) {
                unsigned int pb = pgd_blocks(n_send, 256);
                if (pb > 64) pb = 64;
                k_halo_push_pre<<<pb, 256, 0, st>>>(w_z, p, sc, fl, d_send_idx, n_send, peers, hp, lay, me, world, seq_halo,
                                                    h->counters + (PGD_MAX_COUNTERS - 2));
            }
            if ((rc2 = pgd_spcg_direction(h, d_work, n_owned, block, sc, fl, stream))) return rc2;
            k_spcg_matvec_p2p<<<BK_CTAS_PER_SM * h->sm_count, BK_THREADS, BK_SMEM_BYTES, st>>>(d_rowptr, d_colidx, d_values, p, w_q, n_owned, sc,
                                                                                   fl, h->partials, h->counters, peers, hp, lay,
                                                                                   me, world, seq_halo, seq_ar);
            const int64_t n_nodes = n_owned / block;
            unsigned int vb = pgd_blocks(n_nodes, 256);
            unsigned int capv = (unsigned int)h->sm_count * 8;
            if (vb > capv) vb = capv;
            if (vb < 1) vb = 1;
            if (block == 1)
                k_spcg_update_p2p<1><<<vb, 256, 0, st>>>(d_x, w_r, w_z, p, w_q, w_minv, n_nodes, sc, fl, h->partials, h->counters,
                                                         peers, lay, me, world, seq_ar);
            else if (block == 2)
                k_spcg_update_p2p<2><<<vb, 256, 0, st>>>(d_x, w_r, w_z, p, w_q, w_minv, n_nodes, sc, fl, h->partials, h->counters,
                                                         peers, lay, me, world, seq_ar);
            else
                k_spcg_update_p2p<3><<<vb, 256, 0, st>>>(d_x, w_r, w_z, p, w_q, w_minv, n_nodes, sc, fl, h->partials, h->counters,
                                                         peers, lay, me, world, seq_ar);
            h->n_launches += 3;
            return (int32_t)cudaGetLastError();
        }
        if ((rc2 = pgd_spcg_direction(h, d_work, n_owned, block, sc, fl, stream))) return rc2;
        if (p2p) {
            if (n_send) {
                unsigned int pb = pgd_blocks(n_send, 256);
                if (pb > 64) pb = 64;
                k_halo_push<<<pb, 256, 0, st>>>(p, d_send_idx, n_send, peers, hp, lay, me, world, seq_halo,
                                                h->counters + (PGD_MAX_COUNTERS - 2), fl);
            }
            k_halo_wait<<<1, 32, 0, st>>>((unsigned char*)h->win_local, hp, lay, world, seq_halo, fl);
            h->n_launches += 2;
        } else if (world > 1) {
            if (n_send) {
                k_halo_pack<<<pgd_blocks(n_send, 256), 256, 0, st>>>(p, d_send_idx, n_send, sendbuf);
                h->n_launches += 1;
            }
            PGD_NCCL(h, g_nccl.GroupStart());
            int64_t so = 0, ro = 0;
            for (int r = 0; r < world; ++r) {
                if (h_send_counts[r]) PGD_NCCL(h, g_nccl.Send(sendbuf + so, (size_t)h_send_counts[r], ncclDouble, r, comm, st));
                if (h_recv_counts[r]) PGD_NCCL(h, g_nccl.Recv(p + n_owned + ro, (size_t)h_recv_counts[r], ncclDouble, r, comm, st));
                so += h_send_counts[r];
                ro += h_recv_counts[r];
            }
            PGD_NCCL(h, g_nccl.GroupEnd());
        }
        if ((rc2 = pgd_spcg_matvec(h, d_rowptr, d_colidx, d_values, d_work, n_owned, block, sc, stream))) return rc2;
        if ((rc2 = allreduce(sc + 2, 1, 1, st))) return rc2;
        if ((rc2 = pgd_spcg_update(h, d_x, d_work, n_owned, block, sc, fl, stream))) return rc2;
        if ((rc2 = allreduce(sc + 8, 2, 1, st))) return rc2;
        return pgd_spcg_rotate(h, sc, fl, stream);
    };
    // Peer-window path: the iteration is 8 small launches with device-resident sequence numbers and flags, so
    // it is captured ONCE per solve in a CUDA graph and replayed (the host was the bottleneck at ~7 us per
    // launch: 59 us per iteration on a 137 k-row slab whose kernels take ~15 us).
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    const bool use_graph = p2p && h->opt_graph;
    bool warmed = false;
    while (true) {
        PGD_CUDA(h, pgd_fetch(h, hf, fl, sizeof(hf), nullptr, nullptr, 0, st));
        if (hf[0] || launched >= maxit) break;
        int todo = maxit - launched;
        if (todo > check_every) todo = check_every;
        int i = 0;
        if (use_graph && !warmed) {  // first iteration eagerly (one-time function attributes), then capture
            if ((rc = enqueue_iteration(st))) return rc;
            warmed = true;
            ++i;
            if (i < todo) {
                // the caller's stream may be the legacy default stream, which cannot capture: record the
                // iteration on a private stream and launch the instantiated graph on the caller's stream
                if (!h->cap_stream) {
                    cudaStream_t cs;
                    PGD_CUDA(h, cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
                    h->cap_stream = cs;
                }
                cudaStream_t cs = (cudaStream_t)h->cap_stream;
                PGD_CUDA(h, cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
                int32_t rcc = enqueue_iteration(cs);
                cudaError_t ce = cudaStreamEndCapture(cs, &graph);
                if (rcc) return rcc;
                PGD_CUDA(h, ce);
                PGD_CUDA(h, cudaGraphInstantiate(&gexec, graph, 0));
            }
        }
        for (; i < todo; ++i) {
            if (gexec) {
                PGD_CUDA(h, cudaGraphLaunch(gexec, st));
            } else if ((rc = enqueue_iteration(st))) {
                return rc;
            }
        }
        launched += todo;
    }
    if (gexec) cudaGraphExecDestroy(gexec);
    if (graph) cudaGraphDestroy(graph);
    PGD_CUDA(h, cudaEventRecord(h->ev1, st));
    double hs[8];
    PGD_CUDA(h, pgd_fetch(h, hs, sc, sizeof(hs), nullptr, nullptr, 0, st));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->pcg_ms += ms;
    h->pcg_solves += 1;
    h->pcg_iters += hf[1];
    h->p_override = nullptr;
    if (h_iters) *h_iters = hf[1];
    if (h_relres) *h_relres = (hs[4] > 0.0) ? sqrt(hs[3] / hs[4]) : 0.0;
    if (hf[2] == 2) {
        // the ranks' sequence numbers may have diverged: the window is unusable from here on (fatal for the window, not
        // for the solver: later solves take the NCCL path until a new window is created and opened collectively)
        h->win_world = 0;
        snprintf(h->err, sizeof(h->err), "pgd_spcg_solve_sync: a peer rank did not arrive at a halo / all-reduce flag within "
                 "%d ms; the peer window has been disabled", h->opt_spin_ms);
        return -6;
    }
    if (hf[2]) {
        snprintf(h->err, sizeof(h->err), "pgd_spcg_solve_sync: NaN encountered (matrix not SPD?)");
        return -3;
    }
    return 0;
}
