// Handle lifecycle + small elementwise kernels (lincomb, set_entries, dot, panel dots).
#include "common.cuh"

extern "C" int32_t pgd_abi_version(void) { return 1; }

extern "C" int32_t pgd_create(int32_t device, pgd_handle_t* out) {
    if (!out) return -1;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return (int32_t)e;
    if (device < 0 || device >= count) return -2;
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return (int32_t)e;
    pgd_ctx* h = new pgd_ctx();
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->opt_resident = 1;
    h->opt_stream = 2;
    h->opt_p2p = 1;
    h->opt_graph = 1;
    h->opt_pcg3 = 1;
    h->opt_fused = 0;
    h->opt_persist = 1;
    h->opt_bsr = 2;
    h->opt_ll = 1;
    h->opt_spin_ms = 20000;
    h->opt_single_reduction = 0;
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    if ((e = cudaMalloc(&h->partials, sizeof(double) * PGD_MAX_PARTIALS)) != cudaSuccess ||
        (e = cudaMalloc(&h->counters, sizeof(unsigned int) * PGD_MAX_COUNTERS)) != cudaSuccess ||
        (e = cudaMalloc(&h->scalars, sizeof(double) * 64)) != cudaSuccess ||
        (e = cudaMalloc(&h->flags, sizeof(int) * 16)) != cudaSuccess ||
        (e = cudaMalloc(&h->mailbox, 2048)) != cudaSuccess ||
        (e = cudaMallocHost(&h->pinned, 512)) != cudaSuccess) {
        delete h;
        return (int32_t)e;
    }
    cudaEventCreate(&h->ev0);
    cudaEventCreate(&h->ev1);
    cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming);
    cudaMemset(h->counters, 0, sizeof(unsigned int) * PGD_MAX_COUNTERS);
    cudaMemset(h->scalars, 0, sizeof(double) * 64);
    cudaMemset(h->flags, 0, sizeof(int) * 16);
    cudaMemset(h->mailbox, 0, 2048);
    cudaDeviceSynchronize();
    *out = h;
    return 0;
}

extern "C" int32_t pgd_destroy(pgd_handle_t h) {
    PGD_CHECK_HANDLE(h);
    cudaSetDevice(h->device);
    pgd_free_pattern(h);
    cudaFree(h->partials);
    if (h->arena) cudaFree(h->arena);
    if (h->pinned) cudaFreeHost(h->pinned);
    cudaFree(h->counters);
    cudaFree(h->scalars);
    cudaFree(h->flags);
    cudaFree(h->mailbox);
    cudaEventDestroy(h->ev0);
    cudaEventDestroy(h->ev1);
    cudaEventDestroy(h->ev_done);
    delete h;
    return 0;
}

extern "C" int32_t pgd_get_stats(pgd_handle_t h, int64_t* h_counts, double* h_pcg_ms, int32_t reset) {
    PGD_CHECK_HANDLE(h);
    if (h_counts) {
        h_counts[0] = h->n_launches;
        h_counts[1] = h->pcg_solves;
        h_counts[2] = h->pcg_iters;
        h_counts[3] = h->pcg_resident_solves;
    }
    if (h_pcg_ms) *h_pcg_ms = h->pcg_ms;
    if (reset) {
        h->n_launches = h->pcg_solves = h->pcg_iters = h->pcg_resident_solves = 0;
        h->pcg_ms = 0.0;
    }
    return 0;
}

extern "C" const char* pgd_last_error(pgd_handle_t h) { return h ? h->err : "null handle"; }

extern "C" int32_t pgd_set_option(pgd_handle_t h, const char* name, int64_t value) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, name != nullptr, "null option name");
    if (strcmp(name, "pcg_resident") == 0) {
        h->opt_resident = value ? 1 : 0;
        return 0;
    }
    if (strcmp(name, "fused") == 0) {
        h->opt_fused = value ? 1 : 0;
        return 0;
    }
    if (strcmp(name, "pcg3") == 0) {
        h->opt_pcg3 = value ? 1 : 0;
        return 0;
    }
    if (strcmp(name, "graph") == 0) {
        h->opt_graph = value ? 1 : 0;
        return 0;
    }
    if (strcmp(name, "p2p") == 0) {
        h->opt_p2p = value ? 1 : 0;
        return 0;
    }
    if (strcmp(name, "persist") == 0) {
        h->opt_persist = value < 0 ? 0 : (value > 2 ? 2 : (int)value);
        return 0;
    }
    if (strcmp(name, "ll") == 0) {
        h->opt_ll = value ? 1 : 0;
        return 0;
    }
    if (strcmp(name, "bsr") == 0) {
        h->opt_bsr = value < 0 ? 0 : value > 2 ? 2 : (int)value;
        return 0;
    }
    if (strcmp(name, "single_reduction") == 0) {
        h->opt_single_reduction = value < 0 ? 0 : (value > 2 ? 2 : (int)value);
        return 0;
    }
    if (strcmp(name, "prof") == 0) {
        h->opt_prof = value ? 1 : 0;
        return 0;
    }
    if (strcmp(name, "spin_ms") == 0) {
        h->opt_spin_ms = value < 1 ? 1 : (value > 600000 ? 600000 : (int)value);
        return 0;
    }
    if (strcmp(name, "spmv_stream") == 0) {
        h->opt_stream = (value < 0 || value > 2) ? 2 : (int)value;
        return 0;
    }
    snprintf(h->err, sizeof(h->err), "pgd_set_option: unknown option '%s'", name);
    return -2;
}

// ----------------------------------------------------------------------------- lincomb
#define LC_MAX 24
struct LcArgs {
    const double* x[LC_MAX];
    double c[LC_MAX];
    const double* dc;  // when set: the coefficients of this chunk live on the device (pgd_lincomb_dev)
    int n_terms;
};

// VEC: every pointer is 16-byte aligned -> 128-bit loads/stores, two element pairs in flight per thread
// (torch.dot reads at 7.1 TB/s on this box, the scalar version of this kernel reached 5.2: tools/micro/bw_probe.py).
// Every array is touched exactly once: streaming (evict-first) loads and stores.  ncu on the first version
// (profiles/r02_kernels_128cube_ncu_summary.json): 179 us for 2 atoms of 31.8 M entries, 511 MB read but 469 MB WRITTEN
// for a 254 MB result -- 5.5 TB/s of actual DRAM traffic (85 %), i.e. the kernel was at the memory limit with ~215 MB
// of write-back traffic it did not need.
template <bool VEC>
__global__ void __launch_bounds__(256) k_lincomb(LcArgs a, int64_t n, double* __restrict__ out, int accumulate) {
    if (a.dc) {
#pragma unroll 4
        for (int t = 0; t < a.n_terms; ++t) a.c[t] = __ldg(&a.dc[t]);
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (VEC) {
        const int64_t n2 = n >> 1;
        double2* __restrict__ o2 = reinterpret_cast<double2*>(out);
#pragma unroll 2
        for (int64_t i = gtid; i < n2; i += stride) {
            double2 s = accumulate ? o2[i] : make_double2(0.0, 0.0);
#pragma unroll 4
            for (int t = 0; t < a.n_terms; ++t) {
                const double2 v = __ldcs(reinterpret_cast<const double2*>(a.x[t]) + i);
                s.x += a.c[t] * v.x;
                s.y += a.c[t] * v.y;
            }
            __stcs(o2 + i, s);
        }
        if ((n & 1) && gtid == 0) {
            double s = accumulate ? out[n - 1] : 0.0;
            for (int t = 0; t < a.n_terms; ++t) s += a.c[t] * a.x[t][n - 1];
            out[n - 1] = s;
        }
        return;
    }
    for (int64_t i = gtid; i < n; i += stride) {
        double s = accumulate ? out[i] : 0.0;
#pragma unroll 4
        for (int t = 0; t < a.n_terms; ++t) s += a.c[t] * __ldg(&a.x[t][i]);
        out[i] = s;
    }
}

static int32_t lincomb_impl(pgd_ctx* h, int32_t n_terms, const double* const* h_xs, const double* h_coefs,
                            const double* d_coefs, int64_t n, double* d_out, int32_t accumulate, void* stream) {
    PGD_ARG(h, n_terms >= 0 && n >= 0, "bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) return 0;  // empty vectors are a no-op (their device pointers may be NULL)
    PGD_ARG(h, d_out, "null output");
    if (n_terms == 0 && !accumulate) {
        PGD_CUDA(h, cudaMemsetAsync(d_out, 0, sizeof(double) * n, st));
        return 0;
    }
    unsigned int blocks = pgd_blocks(n, 512);
    unsigned int cap = (unsigned int)h->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    for (int t0 = 0; t0 < n_terms; t0 += LC_MAX) {
        LcArgs a;
        a.n_terms = (n_terms - t0 < LC_MAX) ? (n_terms - t0) : LC_MAX;
        a.dc = d_coefs ? d_coefs + t0 : nullptr;
        uintptr_t al = reinterpret_cast<uintptr_t>(d_out);
        for (int t = 0; t < a.n_terms; ++t) {
            a.x[t] = h_xs[t0 + t];
            a.c[t] = h_coefs ? h_coefs[t0 + t] : 0.0;
            al |= reinterpret_cast<uintptr_t>(a.x[t]);
        }
        if ((al & 15) == 0) k_lincomb<true><<<blocks, 256, 0, st>>>(a, n, d_out, (accumulate || t0 > 0) ? 1 : 0);
        else k_lincomb<false><<<blocks, 256, 0, st>>>(a, n, d_out, (accumulate || t0 > 0) ? 1 : 0);
        PGD_LAUNCH_OK(h);
    }
    return 0;
}

extern "C" int32_t pgd_lincomb(pgd_handle_t h, int32_t n_terms, const double* const* h_xs, const double* h_coefs,
                               int64_t n, double* d_out, int32_t accumulate, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, n_terms == 0 || h_coefs, "null coefficients");
    return lincomb_impl(h, n_terms, h_xs, h_coefs, nullptr, n, d_out, accumulate, stream);
}

extern "C" int32_t pgd_lincomb_dev(pgd_handle_t h, int32_t n_terms, const double* const* h_xs, const double* d_coefs,
                                   int64_t n, double* d_out, int32_t accumulate, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, n_terms == 0 || d_coefs, "null coefficients");
    return lincomb_impl(h, n_terms, h_xs, nullptr, d_coefs, n, d_out, accumulate, stream);
}

// ----------------------------------------------------------------------------- scalar programs
// The separated form's coefficients  c_g = (+-) c0 * prod_j <mode integral j>  are tiny arithmetic expressions over
// mode integrals that already sit in device memory.  Evaluating them here (one thread per coefficient, postfix code
// passed as kernel parameters) instead of on the host removes the device->host round trip between the functionals of
// a sub-problem and its operator / right-hand-side assembly.  Every operation is a separately rounded IEEE double
// operation in the order the host would apply it, so the result is bitwise the host's.
#define SP_MAX_CODE 640
#define SP_MAX_CONST 128
#define SP_MAX_PROG 32
struct SpArgs {
    int n_prog;
    int off[SP_MAX_PROG + 1];
    int code[SP_MAX_CODE];  // (opcode << 24) | operand
    double consts[SP_MAX_CONST];
};
enum { SP_CONST = 0, SP_LOAD = 1, SP_MUL = 2, SP_ADD = 3, SP_SUB = 4, SP_DIV = 5, SP_NEG = 6 };

__global__ void k_scalar_programs(SpArgs a, const double* __restrict__ pool, double* __restrict__ out) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= a.n_prog) return;
    double st[16];
    int sp = 0;
    for (int i = a.off[g]; i < a.off[g + 1]; ++i) {
        const int op = a.code[i] >> 24, arg = a.code[i] & 0xffffff;
        if (op == SP_CONST) st[sp++ & 15] = a.consts[arg];
        else if (op == SP_LOAD) st[sp++ & 15] = pool[arg];
        else if (op == SP_NEG) st[(sp - 1) & 15] = -st[(sp - 1) & 15];
        else {
            const double y = st[(sp - 1) & 15], x = st[(sp - 2) & 15];
            double r;
            if (op == SP_MUL) r = __dmul_rn(x, y);
            else if (op == SP_ADD) r = __dadd_rn(x, y);
            else if (op == SP_SUB) r = __dsub_rn(x, y);
            else r = __ddiv_rn(x, y);
            sp -= 1;
            st[(sp - 1) & 15] = r;
        }
    }
    out[g] = st[(sp - 1) & 15];
}

extern "C" int32_t pgd_scalar_programs(pgd_handle_t h, int32_t n_prog, const int32_t* h_off, const int32_t* h_code,
                                       const double* h_consts, int32_t n_consts, const double* d_pool, double* d_out,
                                       void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, n_prog >= 0 && n_prog <= SP_MAX_PROG && h_off && d_out, "bad arguments (at most 32 programs per call)");
    if (n_prog == 0) return 0;
    PGD_ARG(h, h_off[n_prog] <= SP_MAX_CODE && n_consts <= SP_MAX_CONST && h_off[0] == 0, "program too long");
    SpArgs a;
    a.n_prog = n_prog;
    for (int i = 0; i <= n_prog; ++i) a.off[i] = h_off[i];
    // validate: operands in range, stack depth within 1..16, exactly one value left
    for (int g = 0; g < n_prog; ++g) {
        int depth = 0;
        for (int i = h_off[g]; i < h_off[g + 1]; ++i) {
            const int op = h_code[i] >> 24, arg = h_code[i] & 0xffffff;
            if (op == SP_CONST) {
                PGD_ARG(h, arg < n_consts, "constant index out of range");
                ++depth;
            } else if (op == SP_LOAD) {
                PGD_ARG(h, d_pool != nullptr, "program loads from a null pool");
                ++depth;
            } else if (op == SP_NEG) {
                PGD_ARG(h, depth >= 1, "stack underflow");
            } else {
                PGD_ARG(h, op >= SP_MUL && op <= SP_DIV && depth >= 2, "bad opcode / stack underflow");
                --depth;
            }
            PGD_ARG(h, depth <= 16, "expression too deep");
            a.code[i] = h_code[i];
        }
        PGD_ARG(h, depth == 1, "program must leave exactly one value");
    }
    for (int i = 0; i < n_consts; ++i) a.consts[i] = h_consts[i];
    k_scalar_programs<<<1, SP_MAX_PROG, 0, (cudaStream_t)stream>>>(a, d_pool, d_out);
    PGD_LAUNCH_OK(h);
    return 0;
}

// ----------------------------------------------------------------------------- set_entries
__global__ void k_set_entries(double* x, const int32_t* idx, const double* vals, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[idx[i]] = vals ? vals[i] : 0.0;
}

extern "C" int32_t pgd_set_entries(pgd_handle_t h, double* d_x, const int32_t* d_idx, const double* d_vals,
                                   int64_t n_idx, void* stream) {
    PGD_CHECK_HANDLE(h);
    if (n_idx <= 0) return 0;
    k_set_entries<<<pgd_blocks(n_idx, 256), 256, 0, (cudaStream_t)stream>>>(d_x, d_idx, d_vals, n_idx);
    PGD_LAUNCH_OK(h);
    return 0;
}

// ----------------------------------------------------------------------------- dot / panel dots
// grid (gx, n_vecs): block (bx, m) reduces a slice of P[m,:].x ; last block of row m finishes.
__global__ void __launch_bounds__(256) k_panel_dots(const double* __restrict__ P, int64_t ld, const double* __restrict__ x,
                                                    int64_t n, double* out, double* part, unsigned int* counters) {
    const unsigned int m = blockIdx.y;
    const double* row = P + (size_t)m * ld;
    double s = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (((reinterpret_cast<uintptr_t>(row) | reinterpret_cast<uintptr_t>(x)) & 15) == 0) {  // uniform per row
        const int64_t n2 = n >> 1;
        const double2* __restrict__ r2 = reinterpret_cast<const double2*>(row);
        const double2* __restrict__ x2 = reinterpret_cast<const double2*>(x);
        double s1 = 0.0;
#pragma unroll 4
        for (int64_t i = gtid; i < n2; i += stride) {
            const double2 a = __ldcs(r2 + i), b = __ldg(x2 + i);
            s = fma(a.x, b.x, s);
            s1 = fma(a.y, b.y, s1);
        }
        s += s1;
        if ((n & 1) && gtid == 0) s += row[n - 1] * x[n - 1];
    } else {
#pragma unroll 4
        for (int64_t i = gtid; i < n; i += stride) s += __ldcs(&row[i]) * __ldg(&x[i]);
    }
    s = block_sum(s);
    double v[1] = {s};
    grid_sum_finish<1>(v, part + (size_t)m * gridDim.x, counters + m, out + m, blockIdx.x, gridDim.x);
}

extern "C" int32_t pgd_panel_dots(pgd_handle_t h, const double* d_P, int64_t ld, int32_t n_vecs, const double* d_x,
                                  int64_t n, double* d_out, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, n_vecs >= 0 && n >= 0 && ld >= n, "bad arguments");
    if (n_vecs == 0) return 0;
    PGD_ARG(h, n_vecs <= PGD_MAX_COUNTERS, "too many vectors");
    unsigned int gx = pgd_blocks(n, 256 * 8);
    unsigned int cap = (unsigned int)(PGD_MAX_PARTIALS / n_vecs);
    unsigned int want = (unsigned int)(h->sm_count * 8 / n_vecs) + 1;
    if (gx > want) gx = want;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    dim3 grid(gx, (unsigned int)n_vecs);
    k_panel_dots<<<grid, 256, 0, (cudaStream_t)stream>>>(d_P, ld, d_x, n, d_out, h->partials, h->counters);
    PGD_LAUNCH_OK(h);
    return 0;
}

extern "C" int32_t pgd_dot(pgd_handle_t h, const double* d_x, const double* d_y, int64_t n, double* d_out, void* stream) {
    return pgd_panel_dots(h, d_x, n, 1, d_y, n, d_out, stream);
}
