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

// building blocks defined in pcg.cu
extern "C" int32_t pgd_spcg_init(pgd_handle_t, const int32_t*, const int32_t*, const double*, const double*, double*, int64_t,
                                 int64_t, int32_t, double*, double*, int32_t*, void*);

extern "C" int32_t pgd_spcg_solve_sync(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                                       const double* d_b, double* d_x, int64_t n_owned, int64_t n_local, int32_t block,
                                       const int64_t* d_send_idx, const int64_t* h_send_counts, const int64_t* h_recv_counts,
                                       double rtol, double atol, int32_t maxit, int32_t check_every, double* d_work,
                                       int32_t* h_iters, double* h_relres, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_rowptr && d_colidx && d_values && d_b && d_x && d_work, "null pointer");
    PGD_ARG(h, n_owned >= 0 && n_local >= n_owned && block >= 1 && block <= 3, "bad sizes");
    const int world = h->comm ? h->comm_world : 1;
    PGD_ARG(h, world == 1 || (h_send_counts && h_recv_counts), "split sizes required with more than one rank");
    cudaStream_t st = (cudaStream_t)stream;
    ncclComm_t comm = (ncclComm_t)h->comm;
    int64_t n_send = 0;
    if (world > 1)
        for (int r = 0; r < world; ++r) n_send += h_send_counts[r];
    PGD_ARG(h, n_send == 0 || d_send_idx, "send index list required");
    // work: r z q minv p[n_local] sendbuf[n_send]
    double* p = d_work + n_owned * (3 + block);
    double* sendbuf = p + n_local;
    double* sc = h->scalars;
    int* fl = h->flags;
    int32_t rc = pgd_spcg_init(h, d_rowptr, d_colidx, d_values, d_b, d_x, n_owned, n_local, block, d_work, sc, fl, stream);
    if (rc) return rc;
    if (world > 1) PGD_NCCL(h, g_nccl.AllReduce(sc + 8, sc + 8, 3, ncclDouble, ncclSum, comm, st));
    rc = pgd_spcg_init_fin(h, sc, fl, rtol, atol, stream);
    if (rc) return rc;
    if (check_every < 1) check_every = 1;
    int hf[4] = {0, 0, 0, 0};
    int launched = 0;
    PGD_CUDA(h, cudaEventRecord(h->ev0, st));
    while (true) {
        PGD_CUDA(h, cudaMemcpyAsync(hf, fl, sizeof(hf), cudaMemcpyDeviceToHost, st));
        PGD_CUDA(h, cudaStreamSynchronize(st));
        if (hf[0] || launched >= maxit) break;
        int todo = maxit - launched;
        if (todo > check_every) todo = check_every;
        for (int i = 0; i < todo; ++i) {
            if ((rc = pgd_spcg_direction(h, d_work, n_owned, block, sc, fl, stream))) return rc;
            if (world > 1) {
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
            if ((rc = pgd_spcg_matvec(h, d_rowptr, d_colidx, d_values, d_work, n_owned, block, sc, stream))) return rc;
            if (world > 1) PGD_NCCL(h, g_nccl.AllReduce(sc + 2, sc + 2, 1, ncclDouble, ncclSum, comm, st));
            if ((rc = pgd_spcg_update(h, d_x, d_work, n_owned, block, sc, fl, stream))) return rc;
            if (world > 1) PGD_NCCL(h, g_nccl.AllReduce(sc + 8, sc + 8, 2, ncclDouble, ncclSum, comm, st));
            if ((rc = pgd_spcg_rotate(h, sc, fl, stream))) return rc;
        }
        launched += todo;
    }
    PGD_CUDA(h, cudaEventRecord(h->ev1, st));
    double hs[8];
    PGD_CUDA(h, cudaMemcpyAsync(hs, sc, sizeof(hs), cudaMemcpyDeviceToHost, st));
    PGD_CUDA(h, cudaStreamSynchronize(st));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->pcg_ms += ms;
    h->pcg_solves += 1;
    h->pcg_iters += hf[1];
    if (h_iters) *h_iters = hf[1];
    if (h_relres) *h_relres = (hs[4] > 0.0) ? sqrt(hs[3] / hs[4]) : 0.0;
    if (hf[2]) {
        snprintf(h->err, sizeof(h->err), "pgd_spcg_solve_sync: NaN encountered (matrix not SPD?)");
        return -3;
    }
    return 0;
}
