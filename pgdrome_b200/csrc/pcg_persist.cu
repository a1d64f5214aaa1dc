// Persistent Jacobi / node-block-Jacobi PCG for the HBM-bound regime: ONE cooperative kernel per solve, on one
// GPU or on every GPU of a row-sharded system (the multi-GPU form of pcg_resident.cu).
//
// The matrix does not fit on chip here (configs[2]: 520 MB, configs[3]: 720 MB, 1/N of that per rank), so every
// iteration streams the local rows -- vector operators by the direct node-block walk of spmv_bsr.cuh (template value
// BSR = 2), scalar ones through the TMA ring of spmv_bulk.cuh (BSR = 0; BSR = 1 is the ring with node-block tiles);
// what the persistent form removes is everything BETWEEN the streams: kernel launches, host polling, and -- sharded --
// the 8 dependent launches and three host-visible flag round trips of pgd_spcg_solve_sync (55 us floor per
// iteration).  All CG scalars live in registers, identical in every CTA of every rank.
//
// One iteration (p = [owned | ghost] is purely local; the ghost values of z arrive through the peer window):
//   D  p = z + beta p on ALL local entries -- the ghost entries of p follow the same recurrence from the ghost
//      entries of z, which the neighbours sent during the previous reduction   -> grid barrier (arrival counter)
//   S  q = A_loc p, local p.q                                      -> reduce-broadcast #1 (alpha)
//   U  x += alpha p ; r -= alpha q ; z = M^-1 r ; local r.z, r.r   -> reduce-broadcast #2 (beta, stop test); inside it,
//      as soon as this rank's CTAs have all arrived, a few CTAs store the boundary entries of the new z straight
//      into the neighbours' windows over NVLink: the halo travels while the dot products cross the switch, and the
//      next D finds it in place.  (First version: halo of p pushed in D, every CTA fencing at system scope -- 9.6 us
//      for D and 6.7 us of barrier + halo wait per iteration at 2 GPUs.)
// reduce-broadcast: every CTA deposits its partials, arrives and adds the G partials itself in a fixed order; CTA 0
// sends the rank's sum to every other rank; every CTA of every rank collects the `world - 1` foreign sums from its
// LOCAL window and adds them in RANK ORDER => bitwise identical scalars everywhere, deterministic run to run, one
// uniform stop decision and no host round trip.  On a single GPU the scheme degenerates to a grid barrier.
// Cross-GPU transport ("ll" option, default): sums and halo entries travel as 8-byte words {32 data bits | 32-bit
// sequence number}; a value is complete when both of its words carry the expected sequence, so neither a system-scope
// fence nor a flag sits on the critical path (41 -> 29 us per iteration at 8 GPUs).  "ll" = 0: data, fence, release flag.
// Every spin carries a wall-clock budget (globaltimer) and watches a local abort word, so a missing peer ends
// the solve with -6 instead of hanging the GPU; the window is then marked unusable (sequence numbers may have
// diverged) and later solves use the NCCL path.
#include <cooperative_groups.h>

#include "common.cuh"
#include "pcg_blocks.cuh"
#include "peer_window.cuh"
#include "spmv_bulk.cuh"
#include "spmv_bsr.cuh"

#ifndef PS_THREADS
#define PS_THREADS 512
#endif

int32_t pgd_pcg_finish_impl(pgd_ctx* h, int32_t* h_iters, double* h_relres);

struct PersistArgs {
    const int32_t* rowptr;
    const int32_t* colidx;
    const double* vals;
    const double* b;
    double* x;                 // [n_local]; warm: initial guess with valid ghosts on entry
    int64_t n_owned, n_local;
    double rtol, atol;
    int maxit, warm;
    double *r, *q, *minv;      // [n_owned] (minv: n_owned * BS)
    double* z;                 // [n_local]; inside the peer window when world > 1 (neighbours write its ghost tail)
    double* s;                 // [n_owned], single-reduction variant only (A p by recurrence)
    double* p;                 // [n_local], local
    double* part;              // [2][4][G] partial sums, double-buffered by sync parity
    unsigned int* arrive;      // monotonic arrival counter, 0 at launch
    unsigned int* push_ctr;    // "last pushing CTA" ticket, 0 at launch and whenever idle
    int* abort_word;           // != 0: leave (set by a spin that ran out of budget)
    double* out_sc;            // [0] rr, [1] bb
    int* out_fl;               // [0] iterations, [1] status (0 ok, 2 NaN, 3 peer timeout)
    unsigned long long* seq_ar;    // device-resident sequence numbers (shared with pgd_spcg_solve_sync)
    unsigned long long* seq_halo;
    int world, me;
    PwPeers peers;             // window base of every rank (world == 1: base[0] = the handle's local mailbox block)
    PwLayout lay;
    PwHalo hp;
    const int64_t* send_idx;
    int64_t n_send;
    unsigned int n_push;       // CTAs that take part in a halo push (the LAST n_push of the grid)
    unsigned long long spin_ns;
    BsrPlan bsr;               // node-block walk of the CSR arrays (BSR template only)
    int ll;                    // 1: "LL" mailboxes and halo (data + sequence in one 8-byte word), 0: data, fence, flag
    unsigned long long* prof;  // optional [8]: ns spent (as seen by CTA 0) in D, barrier, S, reduce 1, U, reduce 2
};

#define PS_MARK(slot)                                              \
    do {                                                           \
        if (a.prof && blockIdx.x == 0 && tid == 0) {               \
            const unsigned long long _t = ps_now();                \
            a.prof[slot] += _t - t_mark;                           \
            t_mark = _t;                                           \
        }                                                          \
    } while (0)

__device__ __forceinline__ unsigned long long ps_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned int ps_ld_acquire_gpu(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

struct PsSync {
    unsigned int target;           // arrival target of this CTA (monotonic)
    unsigned long long seq_ar;     // last completed reduce-broadcast
    unsigned long long seq_halo;   // last completed halo exchange
    int parity;                    // partial-sum buffer of the next reduce
};

// Spin until pred() or the budget / abort word ends it; returns false when aborted (and raises the abort word).
template <class Pred>
__device__ __forceinline__ bool ps_spin(Pred&& pred, const PersistArgs& a) {
    if (pred()) return true;
    const unsigned long long t0 = ps_now();
    unsigned int k = 0;
    while (!pred()) {
        if ((++k & 63u) == 0u) {
            if (*((volatile int*)a.abort_word) != 0) return false;
            if (ps_now() - t0 > a.spin_ns) {
                atomicExch(a.abort_word, 1);
                return false;
            }
        }
    }
    return true;
}

// grid barrier: all threads call; false = aborted (uniform per CTA)
__device__ __forceinline__ bool ps_grid_barrier(const PersistArgs& a, PsSync& sy, unsigned int G) {
    __shared__ int s_ok;
    __syncthreads();
    sy.target += G;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(a.arrive, 1u);
        const unsigned int tgt = sy.target;
        s_ok = ps_spin([&] { return ps_ld_acquire_gpu(a.arrive) >= tgt; }, a) ? 1 : 0;
    }
    __syncthreads();
    return s_ok != 0;
}

// Deterministic all-CTA (and all-rank) sum of NV per-thread values; the result lands in v[0..NV) of every thread.
// Two L2 round trips on one GPU (arrival counter, then every CTA adds the G partials itself in the same fixed order:
// the first version, in which CTA 0 alone summed and everybody waited for its mailbox flag, cost 10.5 us per call on
// a B200 -- five dependent round trips and a system-scope fence -- against ~3 us for this one).  Sharded: CTA 0 also
// stores the rank's sum into the other ranks' mailboxes with a release flag, and every CTA acquires the `world - 1`
// foreign flags in its LOCAL window and adds the rank sums in RANK ORDER (its own from the partials).
// false = aborted.
__device__ __forceinline__ void ps_halo_issue(const PersistArgs& a, const double* v, unsigned int G);
__device__ __forceinline__ void ps_halo_commit(const PersistArgs& a, unsigned long long seq, unsigned int G);
__device__ __forceinline__ void ps_halo_issue_ll(const PersistArgs& a, const double* v, unsigned long long seq, unsigned int G);

// halo_v != NULL: the boundary entries of halo_v leave for the neighbours inside this reduction -- the stores are issued
// right after all CTAs of this rank have arrived (their global writes are visible), the flag is published after the
// rank's sum has been sent and BEFORE the wait for the other ranks, so that the acknowledgement round trip of the
// stores and the flight of the flag overlap with the reduction's own cross-rank latency.
template <int NV>
__device__ __forceinline__ bool ps_reduce_bcast(double (&v)[NV], const PersistArgs& a, PsSync& sy, unsigned int G,
                                                const double* halo_v = nullptr, unsigned long long halo_seq = 0) {
    __shared__ double s_res[4];
    __shared__ int s_ok;
    const int tid = threadIdx.x;
    if (tid == 0) s_ok = 1;  // (ordered before every later write by the barriers inside block_sum)
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = block_sum(v[k]);
    double* part = a.part + (size_t)sy.parity * 4 * G;
    sy.parity ^= 1;
    sy.target += G;
    const unsigned long long seq = ++sy.seq_ar;
    const int par = (int)(seq & 1ULL);
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) part[(size_t)k * G + blockIdx.x] = v[k];
        __threadfence();
        atomicAdd(a.arrive, 1u);
        const unsigned int tgt = sy.target;
        if (!ps_spin([&] { return ps_ld_acquire_gpu(a.arrive) >= tgt; }, a)) s_ok = 0;
    }
    __syncthreads();
    if (halo_v) {
        if (a.ll) ps_halo_issue_ll(a, halo_v, halo_seq, G);
        else ps_halo_issue(a, halo_v, G);
    }
    if (tid < 32) {  // every CTA: the rank's sum, same order everywhere
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double s = 0.0;
            for (unsigned int i = tid; i < G; i += 32) s += __ldcg(&part[(size_t)k * G + i]);
            s = warp_sum(s);
            if (tid == 0) s_res[k] = s;
        }
    }
    __syncthreads();
    if (a.world > 1 && a.ll) {
        __shared__ double s_peer[PW_MAXR][PW_AR_VALS];
        const unsigned int flag = (unsigned int)seq;  // differs from the slot's previous content (seq - 2, or 0 at start)
        if (blockIdx.x == 0 && tid < a.world && tid != a.me) {
            unsigned long long* ll = reinterpret_cast<unsigned long long*>(a.peers.base[tid] + a.lay.ll_off()) +
                                     ((size_t)par * PW_MAXR + a.me) * PW_AR_VALS * 2;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const unsigned long long bits = (unsigned long long)__double_as_longlong(s_res[k]);
                st_relaxed_sys(ll + 2 * k, (bits & 0xffffffffULL) | ((unsigned long long)flag << 32));
                st_relaxed_sys(ll + 2 * k + 1, (bits >> 32) | ((unsigned long long)flag << 32));
            }
        }
        if (tid < a.world && tid != a.me) {
            const unsigned long long* ll = reinterpret_cast<const unsigned long long*>(a.peers.base[a.me] + a.lay.ll_off()) +
                                           ((size_t)par * PW_MAXR + tid) * PW_AR_VALS * 2;
            unsigned long long w[2 * NV];
            const bool got = ps_spin(
                [&] {
                    bool all = true;
#pragma unroll
                    for (int j = 0; j < 2 * NV; ++j) {
                        w[j] = ld_relaxed_sys(ll + j);
                        all = all && ((unsigned int)(w[j] >> 32) == flag);
                    }
                    return all;
                },
                a);
            if (!got) s_ok = 0;
#pragma unroll
            for (int k = 0; k < NV; ++k)
                s_peer[tid][k] = __longlong_as_double((long long)((w[2 * k] & 0xffffffffULL) | (w[2 * k + 1] << 32)));
        }
        __syncthreads();
        if (tid == 0 && s_ok) {
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                double s = 0.0;
                for (int q = 0; q < a.world; ++q) s += (q == a.me) ? s_res[k] : s_peer[q][k];
                s_res[k] = s;
            }
        }
        __syncthreads();
    } else if (a.world > 1) {
        if (blockIdx.x == 0 && tid < a.world && tid != a.me) {
            double* slot = reinterpret_cast<double*>(a.peers.base[tid] + a.lay.slot_off()) +
                           ((size_t)par * PW_MAXR + a.me) * PW_AR_VALS;
#pragma unroll
            for (int k = 0; k < NV; ++k) slot[k] = s_res[k];
            __threadfence_system();
            st_release_sys(reinterpret_cast<unsigned long long*>(a.peers.base[tid] + a.lay.arflag_off()) + par * PW_MAXR + a.me,
                           seq);
        }
        if (halo_v) ps_halo_commit(a, halo_seq, G);
        if (tid < a.world && tid != a.me) {
            const unsigned long long* f =
                reinterpret_cast<const unsigned long long*>(a.peers.base[a.me] + a.lay.arflag_off()) + par * PW_MAXR + tid;
            if (!ps_spin([&] { return ld_acquire_sys(f) >= seq; }, a)) s_ok = 0;
        }
        __syncthreads();
        if (tid == 0 && s_ok) {
            const double* mine =
                reinterpret_cast<const double*>(a.peers.base[a.me] + a.lay.slot_off()) + (size_t)par * PW_MAXR * PW_AR_VALS;
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                double s = 0.0;
                for (int q = 0; q < a.world; ++q)
                    s += (q == a.me) ? s_res[k] : *((volatile const double*)&mine[q * PW_AR_VALS + k]);
                s_res[k] = s;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = s_res[k];
    const bool ok = s_ok != 0;
    __syncthreads();  // s_res / s_ok are reused by the next call
    return ok;
}

// Halo push in two halves, executed by the LAST n_push CTAs of the grid only (the others return at once):
//   issue   boundary entries of v -> the neighbours' ghost slots of their z buffer (window offset 0), plain remote stores;
//   commit  one system-scope fence per CTA (by then the stores have long been acknowledged: the cross-rank part of the
//           reduction sits between the two halves), ticket, and the last CTA publishes the halo flag on every destination.
// (Fencing right behind the stores put the NVLink round trip on the critical path of the pushing CTAs: 5.4 us of
// grid-barrier wait per iteration at 2 GPUs.)
__device__ __forceinline__ void ps_halo_issue(const PersistArgs& a, const double* v, unsigned int G) {
    if (a.n_push == 0 || blockIdx.x < G - a.n_push) return;
    const int64_t first = (int64_t)(blockIdx.x - (G - a.n_push)) * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)a.n_push * blockDim.x;
    for (int64_t s = first; s < a.n_send; s += stride) {
        int r = 0;
        while (s >= a.hp.seg_start[r + 1]) ++r;
        double* dst = reinterpret_cast<double*>(a.peers.base[r]) + a.hp.dst_off[r] + (s - a.hp.seg_start[r]);
        *dst = v[a.send_idx[s]];
    }
}
__device__ __forceinline__ void ps_halo_commit(const PersistArgs& a, unsigned long long seq, unsigned int G) {
    if (a.n_push == 0 || blockIdx.x < G - a.n_push) return;
    __shared__ bool s_last;
    __syncthreads();  // every store of this CTA is ordered before thread 0's fence (CTA barrier + cumulativity)
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int t = atomicAdd(a.push_ctr, 1u);
        s_last = (t == a.n_push - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        const int r = threadIdx.x;
        if (r < a.world && r != a.me && a.hp.seg_start[r + 1] > a.hp.seg_start[r])
            st_release_sys(reinterpret_cast<unsigned long long*>(a.peers.base[r] + a.lay.haloflag_off()) + a.me, seq);
        if (threadIdx.x == 0) *a.push_ctr = 0u;
    }
}
__device__ __forceinline__ void ps_halo_push(const PersistArgs& a, const double* v, unsigned long long seq, unsigned int G) {
    ps_halo_issue(a, v, G);
    ps_halo_commit(a, seq, G);
}

// "LL" halo (option "ll", default): every boundary double travels as two 8-byte words {32 data bits | 32-bit sequence
// number} into the destination's LL region (window offset llh_base, two parities), indexed by the destination's local
// ghost index.  8-byte stores are single-copy atomic, so an entry is valid as soon as both words carry the expected
// sequence: no system-scope fence, no ticket and no flag -- one one-way NVLink flight between the sender's store and the
// receiver's use.  The slot's previous content carries seq - 2 (or 0 after a window (re-)open), never seq.
__device__ __forceinline__ void ps_halo_issue_ll(const PersistArgs& a, const double* v, unsigned long long seq, unsigned int G) {
    if (a.n_push == 0 || blockIdx.x < G - a.n_push) return;
    const unsigned long long flag = (unsigned long long)(unsigned int)seq << 32;
    const size_t pbase = (size_t)(seq & 1ULL) * a.lay.llh_cap();
    const int64_t first = (int64_t)(blockIdx.x - (G - a.n_push)) * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)a.n_push * blockDim.x;
    for (int64_t s = first; s < a.n_send; s += stride) {
        int r = 0;
        while (s >= a.hp.seg_start[r + 1]) ++r;
        const size_t j = (size_t)(a.hp.dst_off[r] + (s - a.hp.seg_start[r]));
        unsigned long long* w = reinterpret_cast<unsigned long long*>(a.peers.base[r]) + a.lay.llh_base() + (pbase + j) * 2;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(v[a.send_idx[s]]);
        st_relaxed_sys(w, (bits & 0xffffffffULL) | flag);
        st_relaxed_sys(w + 1, (bits >> 32) | flag);
    }
}
// ghost entry j (local index) of halo `seq`; false = aborted
__device__ __forceinline__ bool ps_ll_read(const PersistArgs& a, unsigned long long seq, int64_t j, double& out) {
    const unsigned long long* w = reinterpret_cast<const unsigned long long*>(a.peers.base[a.me]) + a.lay.llh_base() +
                                  ((size_t)(seq & 1ULL) * a.lay.llh_cap() + (size_t)j) * 2;
    const unsigned int flag = (unsigned int)seq;
    unsigned long long w0 = 0, w1 = 0;
    const bool ok = ps_spin(
        [&] {
            w0 = ld_relaxed_sys(w);
            w1 = ld_relaxed_sys(w + 1);
            return (unsigned int)(w0 >> 32) == flag && (unsigned int)(w1 >> 32) == flag;
        },
        a);
    out = __longlong_as_double((long long)((w0 & 0xffffffffULL) | (w1 << 32)));
    return ok;
}

// all threads call; the threads r < world acquire the halo flag of rank r; false = aborted
__device__ __forceinline__ bool ps_halo_wait(const PersistArgs& a, unsigned long long seq) {
    __shared__ int s_ok2;
    if (a.world <= 1) return true;
    if (threadIdx.x == 0) s_ok2 = 1;
    __syncthreads();
    const int r = threadIdx.x;
    if (r < a.world && a.hp.recv_from[r]) {
        const unsigned long long* f = reinterpret_cast<const unsigned long long*>(a.peers.base[a.me] + a.lay.haloflag_off()) + r;
        if (!ps_spin([&] { return ld_acquire_sys(f) >= seq; }, a)) s_ok2 = 0;
    }
    __syncthreads();
    return s_ok2 != 0;
}

// Gather of the direction vector.  Plain global loads (L1-cached, NOT the read-only .nc path): p is rewritten every
// iteration by other CTAs and -- its ghost tail -- by peer GPUs; every phase that reads it starts behind an acquire
// (grid barrier / halo flag) followed by a CTA barrier, which orders these weak loads after the writers' stores.
struct PsGather {
    const double* x;
    __device__ __forceinline__ double operator()(int c) const { return x[c]; }
};

// q-type pass over the local rows: epi(row, (A v)[row]) through the CSR ring or the node-block walk
template <int BS, int BSR, class Epi>
__device__ __forceinline__ void ps_spmv(const PersistArgs& a, const double* v, Epi&& epi, unsigned char* smem, uint32_t* tile) {
    if constexpr (BSR == 2)
        bb_spmv_direct<BS, PsGather, Epi, PS_THREADS>(a.rowptr, a.vals, a.bsr, a.n_owned, PsGather{v}, epi);
    else if constexpr (BSR == 1)
        bb_spmv_rows<BS, PsGather, Epi, PS_THREADS>(a.rowptr, a.vals, a.bsr, a.n_owned, PsGather{v}, epi, smem, tile);
    else
        bk_spmv_rows<PsGather, Epi, PS_THREADS>(a.rowptr, a.colidx, a.vals, a.n_owned, PsGather{v}, epi, smem, tile);
}

template <int BS, int BSR>
__global__ void __launch_bounds__(PS_THREADS, BK_CTAS_PER_SM) k_pcg_persist(PersistArgs a) {
    extern __shared__ __align__(128) unsigned char bk_smem[];
    const unsigned int G = gridDim.x;
    const int tid = threadIdx.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + tid;
    const int64_t gstride = (int64_t)G * blockDim.x;
    const int64_t no = a.n_owned, n_nodes = a.n_owned / BS;
    PsSync sy;
    sy.target = 0;
    sy.seq_ar = *a.seq_ar;      // nobody writes these before the very end of the kernel
    sy.seq_halo = *a.seq_halo;
    sy.parity = 0;
    uint32_t tile = 0;
    if (tid == 0) {
        if constexpr (BSR == 1) bb_init_barriers(bk_smem);
        else if constexpr (BSR == 0) bk_init_barriers(bk_smem);
    }
    __syncthreads();
    int status = 0;
    int it = 0;
    double rz = 0.0, rz_old = 1.0, rr = 0.0, bb = 0.0, tol2 = 0.0;

    // ---- r = b - A x0 (warm) | b ; M^-1 ; z = M^-1 r ; p = 0
    if (a.warm) {
        auto epi = [&](int64_t row, double s) { a.q[row] = s; };
        ps_spmv<BS, BSR>(a, a.x, epi, bk_smem, &tile);
        __syncthreads();
        if (!ps_grid_barrier(a, sy, G)) status = 3;
    }
    unsigned long long hseq = sy.seq_halo;  // sequence number of the halo of the CURRENT z
    {
        double v[3] = {0.0, 0.0, 0.0};  // r.z, b.b, r.r
        if (status == 0) {
            for (int64_t nd = gtid; nd < n_nodes; nd += gstride) {
                double B[BS][BS], I[BS][BS], rb[BS];
#pragma unroll
                for (int i = 0; i < BS; ++i)
#pragma unroll
                    for (int k = 0; k < BS; ++k)
                        B[i][k] = csr_entry(a.rowptr, a.colidx, a.vals, (int)(nd * BS + i), (int)(nd * BS + k));
                invert_block<BS>(B, I);
#pragma unroll
                for (int i = 0; i < BS; ++i) {
                    rb[i] = a.b[nd * BS + i];
                    v[1] += rb[i] * rb[i];
                    if (a.warm) rb[i] -= a.q[nd * BS + i];
#pragma unroll
                    for (int k = 0; k < BS; ++k) a.minv[(nd * BS + i) * BS + k] = I[i][k];
                }
#pragma unroll
                for (int i = 0; i < BS; ++i) {
                    double zi = 0.0;
#pragma unroll
                    for (int k = 0; k < BS; ++k) zi += I[i][k] * rb[k];
                    const int64_t d = nd * BS + i;
                    a.r[d] = rb[i];
                    a.z[d] = zi;
                    if (!a.warm) a.x[d] = 0.0;
                    v[0] += rb[i] * zi;
                    v[2] += rb[i] * rb[i];
                }
            }
            for (int64_t i = gtid; i < a.n_local; i += gstride) a.p[i] = 0.0;
            hseq = ++sy.seq_halo;
            if (!ps_reduce_bcast<3>(v, a, sy, G, a.z, hseq)) status = 3;
        }
        rz = v[0];
        bb = v[1];
        rr = v[2];
        tol2 = a.rtol * a.rtol * bb;
        if (a.atol * a.atol > tol2) tol2 = a.atol * a.atol;
    }

    unsigned long long t_mark = ps_now();
    if (status == 0 && rr > tol2 && bb > 0.0) {
        while (it < a.maxit) {
            const double beta = (it == 0) ? 0.0 : rz / rz_old;
            // ---- D: p = z + beta p, owned entries first; the ghost entries of z (stored by the neighbours during the
            // reduction above) are acquired only then, and the ghost entries of p follow the same recurrence
            {
                const int64_t noe = no & ~(int64_t)1;  // vector part of the owned range
                if ((((uintptr_t)a.z | (uintptr_t)a.p) & 15) == 0) {
                    const double2* z2 = reinterpret_cast<const double2*>(a.z);
                    double2* p2 = reinterpret_cast<double2*>(a.p);
                    for (int64_t i = gtid; i < (noe >> 1); i += gstride) {
                        const double2 zv = z2[i], ov = p2[i];
                        p2[i] = make_double2(fma(beta, ov.x, zv.x), fma(beta, ov.y, zv.y));
                    }
                    if (noe < no && gtid == 0) a.p[noe] = fma(beta, a.p[noe], a.z[noe]);
                } else {
                    for (int64_t i = gtid; i < no; i += gstride) a.p[i] = fma(beta, a.p[i], a.z[i]);
                }
                if (a.ll) {
                    bool okl = true;
                    for (int64_t i = no + gtid; i < a.n_local; i += gstride) {
                        double zv;
                        okl = ps_ll_read(a, hseq, i, zv) && okl;
                        a.p[i] = fma(beta, a.p[i], zv);
                    }
                    if (__syncthreads_or(!okl)) {
                        status = 3;
                        break;
                    }
                } else {
                    if (!ps_halo_wait(a, hseq)) {
                        status = 3;
                        break;
                    }
                    for (int64_t i = no + gtid; i < a.n_local; i += gstride) a.p[i] = fma(beta, a.p[i], a.z[i]);
                }
            }
            PS_MARK(0);
            if (!ps_grid_barrier(a, sy, G)) {
                status = 3;
                break;
            }
            PS_MARK(1);
            // ---- S: q = A p, p.q
            double pq = 0.0;
            {
                auto epi = [&](int64_t row, double s) {
                    a.q[row] = s;
                    pq = fma(a.p[row], s, pq);
                };
                ps_spmv<BS, BSR>(a, a.p, epi, bk_smem, &tile);
            }
            double v1[1] = {pq};
            PS_MARK(2);
            if (!ps_reduce_bcast<1>(v1, a, sy, G)) {
                status = 3;
                break;
            }
            PS_MARK(3);
            const double alpha = rz / v1[0];
            // ---- U: x, r, z, r.z, r.r; the halo of the new z leaves inside the reduction
            double v2[2] = {0.0, 0.0};
            pcg_update_rows<BS>(a.x, a.r, a.z, a.p, a.q, a.minv, n_nodes, alpha, v2[0], v2[1]);
            PS_MARK(4);
            hseq = ++sy.seq_halo;
            if (!ps_reduce_bcast<2>(v2, a, sy, G, a.z, hseq)) {
                status = 3;
                break;
            }
            PS_MARK(5);
            rz_old = rz;
            rz = v2[0];
            rr = v2[1];
            ++it;
            if (!(rr == rr) || !(rz == rz)) {
                status = 2;
                break;
            }
            if (!(rr > tol2)) break;
        }
    }
    // ---- ghosts of the solution: the boundary entries of x travel through the ghost tail of z in the window
    if (a.world > 1 && status != 3) {
        // every rank has consumed (or will never read) the last z halo: drain it first so that the pushes of x below
        // cannot be overtaken by it, then push x behind one more cross-rank reduction
        if (!a.ll && !ps_halo_wait(a, hseq)) status = 3;  // (LL: the x halo below uses the other parity / another sequence)
        double v0[1] = {0.0};
        if (status != 3 && !ps_reduce_bcast<1>(v0, a, sy, G)) status = 3;
        if (status != 3) {
            if (!(bb > 0.0))
                for (int64_t i = gtid; i < no; i += gstride) a.x[i] = 0.0;  // b = 0: the solution is 0 whatever x0 was
            __syncthreads();
            if (!ps_grid_barrier(a, sy, G)) status = 3;
        }
        if (status != 3) {
            const unsigned long long xseq = ++sy.seq_halo;
            if (a.ll) {
                ps_halo_issue_ll(a, a.x, xseq, G);
                bool okl = true;
                for (int64_t i = no + gtid; i < a.n_local; i += gstride) {
                    double xv;
                    okl = ps_ll_read(a, xseq, i, xv) && okl;
                    a.x[i] = xv;
                }
                if (__syncthreads_or(!okl)) status = 3;
            } else {
                ps_halo_push(a, a.x, xseq, G);
                if (!ps_halo_wait(a, xseq)) status = 3;
                else
                    for (int64_t i = no + gtid; i < a.n_local; i += gstride) a.x[i] = a.z[i];
            }
            // nobody may start the next solve's pushes into the z tail before every rank has copied its ghosts
            if (status != 3 && !ps_reduce_bcast<1>(v0, a, sy, G)) status = 3;
        }
    } else if (!(bb > 0.0) && status == 0) {
        for (int64_t i = gtid; i < no; i += gstride) a.x[i] = 0.0;
    }
    if (blockIdx.x == 0 && tid == 0) {
        a.out_sc[0] = rr;
        a.out_sc[1] = bb;
        a.out_fl[0] = it;
        a.out_fl[1] = status;
        *a.seq_ar = sy.seq_ar;
        *a.seq_halo = sy.seq_halo;
    }
}

// ------------------------------------------------------------------------------------------------
// Single-reduction variant (Chronopoulos & Gear): the same Krylov iteration with ONE grid-wide reduction per step.
//     u = M^-1 r ; w = A u ; gamma = r.u ; delta = w.u ; beta = gamma / gamma_old ; alpha = gamma / (delta - beta gamma / alpha_old)
//     p = u + beta p ; s = w + beta s (= A p, by recurrence) ; x += alpha p ; r -= alpha s
// Per iteration:  V  all vector updates in one pass (p, s, x, r, u = M^-1 r, local r.u and r.r), boundary entries of u
//                    to the neighbours -> grid barrier (+ halo flags)
//                 S  w = A u through the TMA ring, local w.u -> reduce-broadcast of (r.u, w.u, r.r)
// i.e. one barrier and one reduction instead of one barrier and two reductions: on 8 GPUs the two reductions are 23 of
// the 40 us of an iteration (profiles/r02_pcg_bench_sharded8_*).  The price: s is carried by a recurrence instead of being
// recomputed as A p, so the iterates differ from the two-reduction form in the last bits (iteration counts within
// +-1 %, same stopping rule on the recursive residual); selected with pgd_set_option("single_reduction").
template <int BS, int BSR>
__global__ void __launch_bounds__(PS_THREADS, BK_CTAS_PER_SM) k_pcg_persist_sr(PersistArgs a) {
    extern __shared__ __align__(128) unsigned char bk_smem[];
    const unsigned int G = gridDim.x;
    const int tid = threadIdx.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + tid;
    const int64_t gstride = (int64_t)G * blockDim.x;
    const int64_t no = a.n_owned, n_nodes = a.n_owned / BS;
    PsSync sy;
    sy.target = 0;
    sy.seq_ar = *a.seq_ar;
    sy.seq_halo = *a.seq_halo;
    sy.parity = 0;
    uint32_t tile = 0;
    if (tid == 0) {
        if constexpr (BSR == 1) bb_init_barriers(bk_smem);
        else if constexpr (BSR == 0) bk_init_barriers(bk_smem);
    }
    __syncthreads();
    int status = 0, it = 0;
    double* const u = a.z;   // preconditioned residual [owned | ghost] (window when sharded)
    double* const w = a.q;   // A u
    double* const sv = a.s;  // A p by recurrence
    // ---- r = b - A x0 (warm) | b ; M^-1 ; u = M^-1 r ; p = s = 0
    if (a.warm) {
        auto epi = [&](int64_t row, double v) { w[row] = v; };
        ps_spmv<BS, BSR>(a, a.x, epi, bk_smem, &tile);
        __syncthreads();
        if (!ps_grid_barrier(a, sy, G)) status = 3;
    }
    double part[3] = {0.0, 0.0, 0.0};  // r.u, w.u, r.r of the CURRENT u (accumulated over V and S)
    double bb = 0.0;
    if (status == 0) {
        for (int64_t nd = gtid; nd < n_nodes; nd += gstride) {
            double B[BS][BS], I[BS][BS], rb[BS];
#pragma unroll
            for (int i = 0; i < BS; ++i)
#pragma unroll
                for (int k = 0; k < BS; ++k)
                    B[i][k] = csr_entry(a.rowptr, a.colidx, a.vals, (int)(nd * BS + i), (int)(nd * BS + k));
            invert_block<BS>(B, I);
#pragma unroll
            for (int i = 0; i < BS; ++i) {
                rb[i] = a.b[nd * BS + i];
                bb += rb[i] * rb[i];
                if (a.warm) rb[i] -= w[nd * BS + i];
#pragma unroll
                for (int k = 0; k < BS; ++k) a.minv[(nd * BS + i) * BS + k] = I[i][k];
            }
#pragma unroll
            for (int i = 0; i < BS; ++i) {
                double zi = 0.0;
#pragma unroll
                for (int k = 0; k < BS; ++k) zi += I[i][k] * rb[k];
                const int64_t d = nd * BS + i;
                a.r[d] = rb[i];
                u[d] = zi;
                a.p[d] = 0.0;
                sv[d] = 0.0;
                if (!a.warm) a.x[d] = 0.0;
                part[0] += rb[i] * zi;
                part[2] += rb[i] * rb[i];
            }
        }
    }
    double gamma = 0.0, gamma_old = 1.0, alpha = 1.0, beta = 0.0, rr = 0.0, tol2 = 0.0;
    unsigned long long t_mark = ps_now();
    bool first = true;
    while (status == 0) {
        // ---- halo of the current u, grid barrier (every u entry written), halo flags
        const unsigned long long hseq = ++sy.seq_halo;
        PS_MARK(4);
        if (!ps_grid_barrier(a, sy, G)) {
            status = 3;
            break;
        }
        ps_halo_push(a, u, hseq, G);
        if (!ps_halo_wait(a, hseq)) {
            status = 3;
            break;
        }
        PS_MARK(1);
        // ---- S: w = A u, w.u
        {
            double wu = 0.0;
            auto epi = [&](int64_t row, double v) {
                w[row] = v;
                wu = fma(u[row], v, wu);
            };
            ps_spmv<BS, BSR>(a, u, epi, bk_smem, &tile);
            part[1] = wu;
        }
        PS_MARK(2);
        double v4[4] = {part[0], part[1], part[2], first ? bb : 0.0};
        if (!ps_reduce_bcast<4>(v4, a, sy, G)) {
            status = 3;
            break;
        }
        PS_MARK(3);
        if (first) {
            bb = v4[3];
            tol2 = a.rtol * a.rtol * bb;
            if (a.atol * a.atol > tol2) tol2 = a.atol * a.atol;
            first = false;
        }
        const double gamma_new = v4[0], delta = v4[1];
        rr = v4[2];
        if (!(rr == rr) || !(gamma_new == gamma_new)) {
            status = 2;
            break;
        }
        if (!(rr > tol2) || !(bb > 0.0) || it >= a.maxit) break;
        beta = (it == 0) ? 0.0 : gamma_new / gamma_old;
        alpha = (it == 0) ? gamma_new / delta : gamma_new / (delta - beta * gamma_new / alpha);
        gamma_old = gamma_new;
        gamma = gamma_new;
        // ---- V: p = u + beta p ; s = w + beta s ; x += alpha p ; r -= alpha s ; u = M^-1 r ; local r.u, r.r
        part[0] = part[1] = part[2] = 0.0;
        {
            constexpr int CH = (32 / BS) * BS;
            const int lane = tid & 31;
            const int l0 = (lane / BS) * BS;
            const int64_t nwarps = gstride >> 5;
            for (int64_t base = (gtid >> 5) * CH; base < no; base += nwarps * CH) {
                const int64_t d = base + lane;
                const bool act = lane < CH && d < no;
                double rn = 0.0;
                if (act) {
                    const double pn = fma(beta, a.p[d], u[d]);
                    const double sn = fma(beta, sv[d], w[d]);
                    a.p[d] = pn;
                    sv[d] = sn;
                    a.x[d] = fma(alpha, pn, a.x[d]);
                    rn = fma(-alpha, sn, a.r[d]);
                    a.r[d] = rn;
                }
                double rk[BS];
#pragma unroll
                for (int k = 0; k < BS; ++k) rk[k] = __shfl_sync(0xffffffffu, rn, (l0 + k) & 31);
                if (act) {
                    double zi = 0.0;
#pragma unroll
                    for (int k = 0; k < BS; ++k) zi = fma(__ldg(&a.minv[d * BS + k]), rk[k], zi);
                    u[d] = zi;
                    part[0] = fma(rn, zi, part[0]);
                    part[2] = fma(rn, rn, part[2]);
                }
            }
        }
        ++it;
    }
    (void)gamma;
    // ---- ghosts of the solution (as in k_pcg_persist)
    if (a.world > 1 && status != 3) {
        double v0[1] = {0.0};
        if (!ps_reduce_bcast<1>(v0, a, sy, G)) status = 3;
        if (status != 3) {
            if (!(bb > 0.0))
                for (int64_t i = gtid; i < no; i += gstride) a.x[i] = 0.0;
            __syncthreads();
            if (!ps_grid_barrier(a, sy, G)) status = 3;
        }
        if (status != 3) {
            const unsigned long long xseq = ++sy.seq_halo;
            ps_halo_push(a, a.x, xseq, G);
            if (!ps_halo_wait(a, xseq)) status = 3;
            else
                for (int64_t i = no + gtid; i < a.n_local; i += gstride) a.x[i] = u[i];
            if (status != 3 && !ps_reduce_bcast<1>(v0, a, sy, G)) status = 3;
        }
    } else if (!(bb > 0.0) && status == 0) {
        for (int64_t i = gtid; i < no; i += gstride) a.x[i] = 0.0;
    }
    if (blockIdx.x == 0 && tid == 0) {
        a.out_sc[0] = rr;
        a.out_sc[1] = bb;
        a.out_fl[0] = it;
        a.out_fl[1] = status;
        *a.seq_ar = sy.seq_ar;
        *a.seq_halo = sy.seq_halo;
    }
}

/* d_work: r q (stride ns = n_owned rounded up to even) | M^-1 (even(n_owned*block)) | p, z (even(n_local) each; with a
 * peer window z lives there instead) => 2*ns + even(n_owned*block) + 2*even(n_local) + 8 doubles. */
static int32_t persist_run(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                           const double* d_b, double* d_x, int64_t n_owned, int64_t n_local, int32_t block, double rtol,
                           double atol, int32_t maxit, int32_t warm, double* d_work, const int64_t* d_send_idx,
                           const int64_t* h_send_counts, const int64_t* h_recv_counts, const int64_t* h_peer_ghost_base,
                           const int32_t* d_bcol, int32_t max_blocks_per_row, int32_t* h_iters, double* h_relres,
                           void* stream, int defer) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_rowptr && d_colidx && d_values && d_b && d_x && d_work, "null pointer");
    PGD_ARG(h, n_owned > 0 && n_local >= n_owned && block >= 1 && block <= 3 && n_owned % block == 0, "bad sizes");
    PGD_ARG(h, (((uintptr_t)d_colidx | (uintptr_t)d_values) & 15) == 0, "colidx / values must be 16-byte aligned (bulk copies)");
    cudaStream_t st = (cudaStream_t)stream;
    {  // a solve started earlier with pgd_pcg_start and never collected: finish it first
        int32_t rc0 = pgd_pcg_finish_impl(h, nullptr, nullptr);
        if (rc0 < 0) return rc0;
    }
    const bool multi = h_send_counts && h_recv_counts && h_peer_ghost_base && h->win_local && h->win_world > 1;
    PGD_ARG(h, multi || n_local == n_owned, "ghost columns need an opened peer window (pgd_peer_window_open)");
    const int world = multi ? h->win_world : 1;
    PersistArgs a;
    memset(&a, 0, sizeof(a));
    a.rowptr = d_rowptr;
    a.colidx = d_colidx;
    a.vals = d_values;
    a.b = d_b;
    a.x = d_x;
    a.n_owned = n_owned;
    a.n_local = n_local;
    a.rtol = rtol;
    a.atol = atol;
    a.maxit = maxit;
    a.warm = warm ? 1 : 0;
    const int64_t ns = (n_owned + 1) & ~(int64_t)1;
    const int64_t nl2 = (n_local + 1) & ~(int64_t)1;
    a.r = d_work;
    a.q = a.r + ns;
    a.minv = a.q + ns;
    a.p = a.minv + ((n_owned * block + 1) & ~(int64_t)1);
    a.z = a.p + nl2;
    a.s = a.z + nl2;  // single GPU: behind z; sharded (z lives in the window): the local z slot
    a.world = world;
    a.me = multi ? h->win_rank : 0;
    a.lay = PwLayout{multi ? h->win_pcap : 0};
    a.ll = 0;
    const bool sr_requested = h->opt_single_reduction >= 2 || (h->opt_single_reduction == 1 && multi);
    if (multi) {
        PGD_ARG(h, n_local <= h->win_pcap, "peer window too small");
        if (h->opt_ll && !sr_requested) {
            // the protocol must be the same on every rank: a window that is too small for THIS rank is an error, not a
            // silent switch to the flag protocol (the capacity is uniform, n_local is not)
            PGD_ARG(h, (size_t)n_local <= a.lay.llh_base() && (size_t)n_local <= a.lay.llh_cap(),
                    "peer window too small for the LL halo: capacity >= 6 n_local + 16 doubles of the largest rank");
            a.ll = 1;
        }
        a.s = a.z;
        a.z = reinterpret_cast<double*>(h->win_local);
        int64_t n_send = 0;
        a.hp.seg_start[0] = 0;
        for (int r = 0; r < PW_MAXR; ++r) {
            a.peers.base[r] = (unsigned char*)(r < world ? h->win_peer[r] : nullptr);
            a.hp.seg_start[r + 1] = a.hp.seg_start[r] + (r < world ? h_send_counts[r] : 0);
            a.hp.dst_off[r] = r < world ? h_peer_ghost_base[r] : 0;
            a.hp.recv_from[r] = (r < world && h_recv_counts[r] > 0) ? 1 : 0;
            if (r < world) n_send += h_send_counts[r];
        }
        PGD_ARG(h, n_send == 0 || d_send_idx, "send index list required");
        a.send_idx = d_send_idx;
        a.n_send = n_send;
    } else {
        a.peers.base[0] = reinterpret_cast<unsigned char*>(h->mailbox);
    }
    a.part = h->partials;
    a.arrive = h->counters + (PGD_MAX_COUNTERS - 3);
    a.push_ctr = h->counters + (PGD_MAX_COUNTERS - 4);
    a.abort_word = h->flags + 12;
    a.out_sc = h->scalars + 32;
    a.out_fl = h->flags + 8;
    a.seq_halo = reinterpret_cast<unsigned long long*>(h->scalars + 40);
    a.seq_ar = reinterpret_cast<unsigned long long*>(h->scalars + (multi ? 41 : 42));  // the local mailbox has its own sequence
    a.spin_ns = (unsigned long long)h->opt_spin_ms * 1000000ULL;
    a.prof = h->opt_prof ? reinterpret_cast<unsigned long long*>(h->scalars + 48) : nullptr;  // read back with pgd_get_phase_ns
    PGD_CUDA(h, cudaMemsetAsync(a.arrive - 1, 0, 2 * sizeof(unsigned int), st));  // push_ctr, arrive
    PGD_CUDA(h, cudaMemsetAsync(a.out_fl, 0, 2 * sizeof(int), st));
    PGD_CUDA(h, cudaMemsetAsync(a.abort_word, 0, sizeof(int), st));
    // node-block walk ("bsr" option): 1 = tiles through the TMA ring (lanes per block row = the power of two covering
    // the longest block row, block rows per tile = CTA size / lanes, the tile must fit one stage); 2 = direct walk
    // (no shared memory; lanes per block row = the power of two covering a third of the longest row segment)
    int bsr = 0;
    if (d_bcol && block > 1 && max_blocks_per_row > 0 && h->opt_bsr == 1) {
        int lpr = 1;
        while (lpr < 32 && lpr < max_blocks_per_row) lpr *= 2;
        const int nbr = PS_THREADS / lpr;
        if (nbr <= BB_MAXNBR && (int64_t)nbr * block * block * max_blocks_per_row <= BK_CAP &&
            (((uintptr_t)d_bcol) & 15) == 0) {
            bsr = 1;
            a.bsr.bcol = d_bcol;
            a.bsr.nbr = nbr;
            a.bsr.lpr = lpr;
        }
    } else if (d_bcol && max_blocks_per_row > 0 && h->opt_bsr >= 2) {
        // (block == 1 with bcol = colidx is the plain CSR row walk; measured SLOWER than the ring on 15-entry rows --
        // 0.361 vs 0.216 ms at configs[3] -- so the host side only passes a block-column list for vector operators)
        int lpr = 4;
        while (lpr < 32 && 3 * lpr < block * max_blocks_per_row) lpr *= 2;
        bsr = 2;
        a.bsr.bcol = d_bcol;
        a.bsr.nbr = PS_THREADS / lpr;
        a.bsr.lpr = lpr;
    }
    const bool sr = sr_requested;  // "single_reduction": 0 = never (default), 1 = sharded solves, 2 = always
    const void* fn = nullptr;
#define PS_PICK(K, B) (bsr == 2 ? (const void*)K<B, 2> : bsr == 1 ? (const void*)K<B, 1> : (const void*)K<B, 0>)
    if (sr) {
        if (block == 1) fn = bsr == 2 ? (const void*)k_pcg_persist_sr<1, 2> : (const void*)k_pcg_persist_sr<1, 0>;
        else if (block == 2) fn = PS_PICK(k_pcg_persist_sr, 2);
        else fn = PS_PICK(k_pcg_persist_sr, 3);
    } else if (block == 1) fn = bsr == 2 ? (const void*)k_pcg_persist<1, 2> : (const void*)k_pcg_persist<1, 0>;
    else if (block == 2) fn = PS_PICK(k_pcg_persist, 2);
    else fn = PS_PICK(k_pcg_persist, 3);
#undef PS_PICK
    const size_t smem = bsr == 2 ? 0 : bsr == 1 ? BB_SMEM_BYTES : BK_SMEM_BYTES;
    PGD_CUDA(h, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PGD_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, PS_THREADS, smem));
    if (per_sm < 1) {
        snprintf(h->err, sizeof(h->err), "pgd_pcg_persist_sync: kernel does not fit an SM");
        return -4;
    }
    if (per_sm > BK_CTAS_PER_SM) per_sm = BK_CTAS_PER_SM;
    const int G = per_sm * h->sm_count;
    if ((size_t)G * 8 > PGD_MAX_PARTIALS) return -4;
    a.n_push = multi ? (unsigned int)((a.n_send + PS_THREADS - 1) / PS_THREADS) : 0u;
    if (a.n_push > (unsigned int)G) a.n_push = (unsigned int)G;
    void* kargs[] = {&a};
    PGD_CUDA(h, cudaEventRecord(h->ev0, st));
    PGD_CUDA(h, cudaLaunchCooperativeKernel(fn, dim3(G), dim3(PS_THREADS), kargs, smem, st));
    h->n_launches += 1;
    PGD_CUDA(h, cudaEventRecord(h->ev1, st));
    if (defer) {
        // single GPU only (checked by the caller): results -> the handle's page-locked block, collected by pgd_pcg_finish
        // (exactly the hand-over of the SM-resident solver, pcg_resident.cu); the host goes on recording the next
        // sub-problem while the kernel iterates
        char* pin = static_cast<char*>(h->pinned) + 256;
        PGD_CUDA(h, cudaMemcpyAsync(pin, a.out_fl, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
        PGD_CUDA(h, cudaMemcpyAsync(pin + 64, a.out_sc, 2 * sizeof(double), cudaMemcpyDeviceToHost, st));
        PGD_CUDA(h, cudaEventRecord(h->ev_done, st));
        h->res_pending = 1;
        h->res_kind = 2;
        h->res_stream = (void*)st;
        if (h_iters) *h_iters = -1;
        return 0;
    }
    int hf[2] = {0, 0};
    double hs[2] = {0.0, 0.0};
    PGD_CUDA(h, pgd_fetch(h, hf, a.out_fl, sizeof(hf), hs, a.out_sc, sizeof(hs), st));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->pcg_ms += ms;
    h->pcg_solves += 1;
    h->pcg_iters += hf[0];
    if (h_iters) *h_iters = hf[0];
    if (h_relres) *h_relres = (hs[1] > 0.0) ? sqrt(hs[0] / hs[1]) : 0.0;
    if (hf[1] == 3) {
        // sequence numbers of the ranks may have diverged: the window is not usable any more (NCCL path from now on)
        if (multi) h->win_world = 0;
        snprintf(h->err, sizeof(h->err),
                 "pgd_pcg_persist_sync: a CTA or peer rank did not arrive within %d ms; the peer window has been disabled",
                 h->opt_spin_ms);
        return -6;
    }
    if (hf[1] == 2) {
        snprintf(h->err, sizeof(h->err), "pgd_pcg_persist_sync: NaN encountered (matrix not SPD?)");
        return -3;
    }
    return 0;
}

extern "C" int32_t pgd_pcg_persist_sync(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                                        const double* d_b, double* d_x, int64_t n_owned, int64_t n_local, int32_t block,
                                        double rtol, double atol, int32_t maxit, int32_t warm, double* d_work,
                                        const int64_t* d_send_idx, const int64_t* h_send_counts, const int64_t* h_recv_counts,
                                        const int64_t* h_peer_ghost_base, const int32_t* d_bcol, int32_t max_blocks_per_row,
                                        int32_t* h_iters, double* h_relres, void* stream) {
    return persist_run(h, d_rowptr, d_colidx, d_values, d_b, d_x, n_owned, n_local, block, rtol, atol, maxit, warm, d_work,
                       d_send_idx, h_send_counts, h_recv_counts, h_peer_ghost_base, d_bcol, max_blocks_per_row, h_iters,
                       h_relres, stream, 0);
}

/* Single-GPU form that does not wait: enqueue the solve and return; iterations / residual / errors are collected by
 * pgd_pcg_finish (one solve in flight per handle; every other solver entry point finishes a pending one first). */
extern "C" int32_t pgd_pcg_persist_start(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                                         const double* d_b, double* d_x, int64_t n, int32_t block, double rtol, double atol,
                                         int32_t maxit, int32_t warm, double* d_work, const int32_t* d_bcol,
                                         int32_t max_blocks_per_row, void* stream) {
    return persist_run(h, d_rowptr, d_colidx, d_values, d_b, d_x, n, n, block, rtol, atol, maxit, warm, d_work, nullptr, nullptr,
                       nullptr, nullptr, d_bcol, max_blocks_per_row, nullptr, nullptr, stream, 1);
}

/* Phase profile of the persistent kernel (pgd_set_option "prof" = 1): accumulated nanoseconds, as seen by CTA 0, in
 * [0] direction + halo push, [1] grid barrier + halo wait, [2] SpMV, [3] reduce-broadcast of p.q, [4] vector update,
 * [5] reduce-broadcast of r.z / r.r since the last reset.  Diagnostic only (tools/pcg_bench.py). */
extern "C" int32_t pgd_get_phase_ns(pgd_handle_t h, int64_t* h_ns6, int32_t reset) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, h_ns6 != nullptr, "null pointer");
    PGD_CUDA(h, cudaMemcpy(h_ns6, h->scalars + 48, 6 * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (reset) PGD_CUDA(h, cudaMemset(h->scalars + 48, 0, 8 * sizeof(double)));
    return 0;
}
