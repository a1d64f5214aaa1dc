// CSR sparsity = union of per-cell dof cliques (DOLFIN SparsityPatternBuilder semantics), built by
// sorting the n_cells*ndl^2 (row,col) contribution keys.  The stable sorted order doubles as the
// deterministic gather list (gptr/gidx) that assembly reduces over -- no floating-point atomics.
#include <cub/cub.cuh>

#include "common.cuh"

void pgd_free_pattern(pgd_ctx* h) {
    if (!h->pat_in_arena) {
        cudaFree(h->pat_rowptr);
        cudaFree(h->pat_colidx);
        cudaFree(h->pat_gptr);
        cudaFree(h->pat_gidx);
    }
    h->pat_in_arena = 0;
    h->pat_rowptr = h->pat_colidx = h->pat_gidx = nullptr;
    h->pat_gptr = nullptr;
    h->pat_nnz = 0;
}

__global__ void k_make_pair_keys(const int32_t* __restrict__ cell_dofs, int64_t n_contrib, int ndl, int64_t n_dofs,
                                 int64_t* keys, int32_t* vals) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_contrib) return;
    int64_t cell = t / (ndl * ndl);
    int loc = (int)(t - cell * ndl * ndl);
    int a = loc / ndl, b = loc - a * ndl;
    int64_t r = cell_dofs[cell * ndl + a], c = cell_dofs[cell * ndl + b];
    keys[t] = r * n_dofs + c;
    vals[t] = (int32_t)t;
}

__global__ void k_make_dof_keys(const int32_t* __restrict__ cell_dofs, int64_t n, int64_t* keys, int32_t* vals) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    keys[t] = cell_dofs[t];
    vals[t] = (int32_t)t;
}

__global__ void k_head_flags(const int64_t* __restrict__ keys, int64_t n, int32_t* flags) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    flags[t] = (t == 0 || keys[t] != keys[t - 1]) ? 1 : 0;
}

// gid[t] = inclusive scan of head flags (1-based group id). Thread of each head writes group data.
__global__ void k_groups(const int64_t* __restrict__ keys, const int32_t* __restrict__ gid, int64_t n, int64_t n_dofs,
                         int64_t n_groups, int64_t* gptr, int32_t* colidx, int32_t* rowptr) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    bool head = (t == 0) || (keys[t] != keys[t - 1]);
    if (!head) return;
    int64_t g = gid[t] - 1;
    gptr[g] = t;
    int64_t row = keys[t] / n_dofs;
    if (colidx) colidx[g] = (int32_t)(keys[t] - row * n_dofs);
    if (rowptr) {
        int64_t prev_row = (t == 0) ? -1 : keys[t - 1] / n_dofs;
        for (int64_t r = prev_row + 1; r <= row; ++r) rowptr[r] = (int32_t)g;
    }
    if (g == n_groups - 1) {
        gptr[n_groups] = n;
        if (rowptr)
            for (int64_t r = row + 1; r <= n_dofs; ++r) rowptr[r] = (int32_t)n_groups;
    }
}

// keyed by plain dof: group g == dof value (dense 0..n_dofs-1, empty groups allowed)
__global__ void k_vecmap_ptr(const int64_t* __restrict__ keys, int64_t n, int64_t n_dofs, int64_t* vptr) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    int64_t k = keys[t];
    int64_t prev = (t == 0) ? -1 : keys[t - 1];
    for (int64_t r = prev + 1; r <= k; ++r) vptr[r] = t;
    if (t == n - 1)
        for (int64_t r = k + 1; r <= n_dofs; ++r) vptr[r] = n;
}

static int bits_for(int64_t v) {
    int b = 1;
    while (b < 63 && ((int64_t)1 << b) <= v) ++b;
    return b;
}

struct SortBufs {
    int64_t *k0 = nullptr, *k1 = nullptr;
    int32_t *v0 = nullptr, *v1 = nullptr;
    void* temp = nullptr;
    bool owned = true;  // false: carved out of the handle's arena
    ~SortBufs() {
        if (!owned) return;
        cudaFree(k0);
        cudaFree(k1);
        cudaFree(v0);
        cudaFree(v1);
        cudaFree(temp);
    }
};

static inline size_t up256(size_t b) { return (b + 255) & ~(size_t)255; }

// bump allocator over one block
struct Carver {
    char* p;
    template <typename T>
    T* take(size_t count) {
        T* r = reinterpret_cast<T*>(p);
        p += up256(sizeof(T) * count);
        return r;
    }
};

extern "C" int32_t pgd_pattern_build_sync(pgd_handle_t h, const int32_t* d_cell_dofs, int64_t n_cells, int32_t ndl,
                                          int64_t n_dofs, int64_t* h_nnz, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_cell_dofs && n_cells > 0 && ndl > 0 && n_dofs > 0 && h_nnz, "bad arguments");
    int64_t n = n_cells * ndl * ndl;
    PGD_ARG(h, n < ((int64_t)1 << 31), "n_cells*ndl^2 must be < 2^31");
    PGD_ARG(h, n_dofs < ((int64_t)1 << 31), "n_dofs must be < 2^31");
    cudaStream_t st = (cudaStream_t)stream;
    pgd_set_device(h);
    pgd_free_pattern(h);
    SortBufs B;
    size_t tb = 0, tb2 = 0;
    const int end_bit = bits_for(n_dofs * n_dofs);
    PGD_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, tb, B.k0, B.k1, B.v0, B.v1, (int)n, 0, end_bit, st));
    PGD_CUDA(h, cub::DeviceScan::InclusiveSum(nullptr, tb2, B.v0, B.v0, (int)n, st));
    if (tb2 > tb) tb = tb2;
    // everything (sort buffers, CUB scratch, the pattern with nnz bounded by the contribution count) in the handle's
    // arena when that is small enough to keep; otherwise per-call allocations sized by the real nnz
    const size_t need = 2 * up256(sizeof(int64_t) * n) + 2 * up256(sizeof(int32_t) * n) + up256(tb) +
                        up256(sizeof(int32_t) * (n_dofs + 1)) + up256(sizeof(int32_t) * n) + up256(sizeof(int64_t) * (n + 1));
    const bool in_arena = need <= PGD_ARENA_KEEP;
    Carver cv{nullptr};
    if (in_arena) {
        void* base = nullptr;
        bool temporary = false;
        PGD_CUDA(h, pgd_arena(h, need, &base, &temporary));
        cv.p = static_cast<char*>(base);
        B.owned = false;
        B.k0 = cv.take<int64_t>(n);
        B.k1 = cv.take<int64_t>(n);
        B.v0 = cv.take<int32_t>(n);
        B.v1 = cv.take<int32_t>(n);
        B.temp = cv.take<char>(tb);
    } else {
        PGD_CUDA(h, cudaMalloc(&B.k0, sizeof(int64_t) * n));
        PGD_CUDA(h, cudaMalloc(&B.k1, sizeof(int64_t) * n));
        PGD_CUDA(h, cudaMalloc(&B.v0, sizeof(int32_t) * n));
        PGD_CUDA(h, cudaMalloc(&B.v1, sizeof(int32_t) * n));
        PGD_CUDA(h, cudaMalloc(&B.temp, tb));
    }
    k_make_pair_keys<<<pgd_blocks(n, 256), 256, 0, st>>>(d_cell_dofs, n, ndl, n_dofs, B.k0, B.v0);
    PGD_LAUNCH_OK(h);
    PGD_CUDA(h, cub::DeviceRadixSort::SortPairs(B.temp, tb, B.k0, B.k1, B.v0, B.v1, (int)n, 0, end_bit, st));
    // sorted keys in k1, contribution ids in v1.  reuse v0 as flags/gid.
    k_head_flags<<<pgd_blocks(n, 256), 256, 0, st>>>(B.k1, n, B.v0);
    PGD_LAUNCH_OK(h);
    PGD_CUDA(h, cub::DeviceScan::InclusiveSum(B.temp, tb, B.v0, B.v0, (int)n, st));
    int32_t nnz32 = 0;
    PGD_CUDA(h, pgd_fetch(h, &nnz32, B.v0 + (n - 1), sizeof(int32_t), nullptr, nullptr, 0, st));
    int64_t nnz = nnz32;
    if (in_arena) {
        h->pat_rowptr = cv.take<int32_t>(n_dofs + 1);
        h->pat_colidx = cv.take<int32_t>(n);
        h->pat_gptr = cv.take<int64_t>(n + 1);
        h->pat_in_arena = 1;
    } else {
        PGD_CUDA(h, cudaMalloc(&h->pat_rowptr, sizeof(int32_t) * (n_dofs + 1)));
        PGD_CUDA(h, cudaMalloc(&h->pat_colidx, sizeof(int32_t) * nnz));
        PGD_CUDA(h, cudaMalloc(&h->pat_gptr, sizeof(int64_t) * (nnz + 1)));
    }
    k_groups<<<pgd_blocks(n, 256), 256, 0, st>>>(B.k1, B.v0, n, n_dofs, nnz, h->pat_gptr, h->pat_colidx, h->pat_rowptr);
    PGD_LAUNCH_OK(h);
    h->pat_gidx = B.v1;  // keep the sorted contribution ids
    if (B.owned) B.v1 = nullptr;
    h->pat_nnz = nnz;
    h->pat_ndofs = n_dofs;
    h->pat_ncontrib = n;
    PGD_CUDA(h, cudaStreamSynchronize(st));
    *h_nnz = nnz;
    return 0;
}

extern "C" int32_t pgd_pattern_export(pgd_handle_t h, int32_t* d_rowptr, int32_t* d_colidx, int64_t* d_gptr,
                                      int32_t* d_gidx, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, h->pat_rowptr != nullptr, "no pending pattern (call pgd_pattern_build_sync first)");
    cudaStream_t st = (cudaStream_t)stream;
    const cudaMemcpyKind k = cudaMemcpyDeviceToDevice;
    if (d_rowptr) PGD_CUDA(h, cudaMemcpyAsync(d_rowptr, h->pat_rowptr, sizeof(int32_t) * (h->pat_ndofs + 1), k, st));
    if (d_colidx) PGD_CUDA(h, cudaMemcpyAsync(d_colidx, h->pat_colidx, sizeof(int32_t) * h->pat_nnz, k, st));
    if (d_gptr) PGD_CUDA(h, cudaMemcpyAsync(d_gptr, h->pat_gptr, sizeof(int64_t) * (h->pat_nnz + 1), k, st));
    if (d_gidx) PGD_CUDA(h, cudaMemcpyAsync(d_gidx, h->pat_gidx, sizeof(int32_t) * h->pat_ncontrib, k, st));
    PGD_CUDA(h, cudaStreamSynchronize(st));
    pgd_free_pattern(h);
    // a new sparsity structure exists: whatever the solver remembered about an older one (nnz, "fits the SM-resident
    // PCG") may belong to freed memory whose address the caller's allocator hands out again
    h->nnz_key = nullptr;
    h->fit_key = nullptr;
    return 0;
}

extern "C" int32_t pgd_vecmap_build_sync(pgd_handle_t h, const int32_t* d_cell_dofs, int64_t n_cells, int32_t ndl,
                                         int64_t n_dofs, int64_t* d_vptr, int32_t* d_vidx, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_cell_dofs && n_cells > 0 && ndl > 0 && n_dofs > 0 && d_vptr && d_vidx, "bad arguments");
    int64_t n = n_cells * ndl;
    PGD_ARG(h, n < ((int64_t)1 << 31), "n_cells*ndl must be < 2^31");
    cudaStream_t st = (cudaStream_t)stream;
    pgd_set_device(h);
    PGD_ARG(h, !h->pat_in_arena, "a pattern is pending in the handle's scratch (call pgd_pattern_export first)");
    SortBufs B;
    size_t tb = 0;
    const int end_bit = bits_for(n_dofs);
    PGD_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, tb, B.k0, B.k1, B.v0, d_vidx, (int)n, 0, end_bit, st));
    const size_t need = 2 * up256(sizeof(int64_t) * n) + up256(sizeof(int32_t) * n) + up256(tb);
    if (need <= PGD_ARENA_KEEP) {
        void* base = nullptr;
        bool temporary = false;
        PGD_CUDA(h, pgd_arena(h, need, &base, &temporary));
        Carver cv{static_cast<char*>(base)};
        B.owned = false;
        B.k0 = cv.take<int64_t>(n);
        B.k1 = cv.take<int64_t>(n);
        B.v0 = cv.take<int32_t>(n);
        B.temp = cv.take<char>(tb);
    } else {
        PGD_CUDA(h, cudaMalloc(&B.k0, sizeof(int64_t) * n));
        PGD_CUDA(h, cudaMalloc(&B.k1, sizeof(int64_t) * n));
        PGD_CUDA(h, cudaMalloc(&B.v0, sizeof(int32_t) * n));
        PGD_CUDA(h, cudaMalloc(&B.temp, tb));
    }
    k_make_dof_keys<<<pgd_blocks(n, 256), 256, 0, st>>>(d_cell_dofs, n, B.k0, B.v0);
    PGD_LAUNCH_OK(h);
    PGD_CUDA(h, cub::DeviceRadixSort::SortPairs(B.temp, tb, B.k0, B.k1, B.v0, d_vidx, (int)n, 0, end_bit, st));
    k_vecmap_ptr<<<pgd_blocks(n, 256), 256, 0, st>>>(B.k1, n, n_dofs, d_vptr);
    PGD_LAUNCH_OK(h);
    PGD_CUDA(h, cudaStreamSynchronize(st));
    return 0;
}
