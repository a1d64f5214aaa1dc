// Separated-form assembly kernels (fp64).
//   k_elem_bilinear / k_elem_linear : generic Lagrange element kernels, tables staged in shared
//       memory, one thread per (cell, node pair) / (cell, node); output reduced by k_gather.
//   k_gather        : deterministic fixed-order reduction of element contributions into CSR values.
//   k_assemble_p1   : fused P1 simplex mass+stiffness+advection with folded scalar coefficients,
//       written straight into the CSR pattern through the gather list (4 lanes per nonzero).
#include "common.cuh"

template <int TDIM, int GDIM>
struct Geo {
    double Jinv[TDIM][GDIM];  // grad_g phi = sum_t dphi_t * Jinv[t][g]   (only when TDIM == GDIM)
    double detJ;              // |det J| (or sqrt(det J^T J) for embedded facets)
};

template <int TDIM, int GDIM>
__device__ __forceinline__ void load_geo(const double* __restrict__ coords, const int32_t* __restrict__ cv, int64_t cell,
                                         Geo<TDIM, GDIM>& G) {
    double X[TDIM + 1][GDIM];
#pragma unroll
    for (int v = 0; v <= TDIM; ++v) {
        int64_t vi = cv[cell * (TDIM + 1) + v];
#pragma unroll
        for (int g = 0; g < GDIM; ++g) X[v][g] = __ldg(&coords[vi * GDIM + g]);
    }
    double J[GDIM][TDIM];
#pragma unroll
    for (int g = 0; g < GDIM; ++g)
#pragma unroll
        for (int t = 0; t < TDIM; ++t) J[g][t] = X[t + 1][g] - X[0][g];
    if constexpr (TDIM == 1 && GDIM == 1) {
        G.detJ = fabs(J[0][0]);
        G.Jinv[0][0] = 1.0 / J[0][0];
    } else if constexpr (TDIM == 2 && GDIM == 2) {
        double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        double id = 1.0 / det;
        G.detJ = fabs(det);
        G.Jinv[0][0] = J[1][1] * id;
        G.Jinv[0][1] = -J[0][1] * id;
        G.Jinv[1][0] = -J[1][0] * id;
        G.Jinv[1][1] = J[0][0] * id;
    } else if constexpr (TDIM == 3 && GDIM == 3) {
        double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
        double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
        double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
        double id = 1.0 / det;
        G.detJ = fabs(det);
        // inverse = adj / det ; Jinv[t][g] = (J^-1)[t][g]
        G.Jinv[0][0] = c00 * id;
        G.Jinv[1][0] = c01 * id;
        G.Jinv[2][0] = c02 * id;
        G.Jinv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
        G.Jinv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
        G.Jinv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
        G.Jinv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
        G.Jinv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
        G.Jinv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    } else if constexpr (TDIM == 1) {  // segment embedded in 2-D / 3-D
        double s = 0.0;
#pragma unroll
        for (int g = 0; g < GDIM; ++g) s += J[g][0] * J[g][0];
        G.detJ = sqrt(s);
#pragma unroll
        for (int g = 0; g < GDIM; ++g) G.Jinv[0][g] = 0.0;
    } else {  // triangle embedded in 3-D
        double cx = J[1][0] * J[2][1] - J[2][0] * J[1][1];
        double cy = J[2][0] * J[0][1] - J[0][0] * J[2][1];
        double cz = J[0][0] * J[1][1] - J[1][0] * J[0][1];
        G.detJ = sqrt(cx * cx + cy * cy + cz * cz);
#pragma unroll
        for (int t = 0; t < TDIM; ++t)
#pragma unroll
            for (int g = 0; g < GDIM; ++g) G.Jinv[t][g] = 0.0;
    }
}

// D[0] = phi_a(q), D[1+g] = d phi_a / d x_g
template <int TDIM, int GDIM>
__device__ __forceinline__ void slots(const double* s_phi, const double* s_dphi, int q, int a, int nd,
                                      const Geo<TDIM, GDIM>& G, double (&D)[GDIM + 1]) {
    D[0] = s_phi[q * nd + a];
#pragma unroll
    for (int g = 0; g < GDIM; ++g) {
        double s = 0.0;
#pragma unroll
        for (int t = 0; t < TDIM; ++t) s += s_dphi[(q * nd + a) * TDIM + t] * G.Jinv[t][g];
        D[1 + g] = s;
    }
}

template <int TDIM, int GDIM, int BS>
__global__ void __launch_bounds__(128) k_elem_bilinear(const double* __restrict__ coords, const int32_t* __restrict__ cv,
                                                       int64_t n_cells, int nd, int nq, const double* __restrict__ phi,
                                                       const double* __restrict__ dphi, const double* __restrict__ qw,
                                                       const double* __restrict__ wq, const double* __restrict__ T,
                                                       double* __restrict__ Ae) {
    extern __shared__ double sm[];
    constexpr int S = GDIM + 1;
    double* s_phi = sm;
    double* s_dphi = s_phi + nq * nd;
    double* s_qw = s_dphi + nq * nd * TDIM;
    double* s_T = s_qw + nq;
    for (int i = threadIdx.x; i < nq * nd; i += blockDim.x) s_phi[i] = phi[i];
    for (int i = threadIdx.x; i < nq * nd * TDIM; i += blockDim.x) s_dphi[i] = dphi[i];
    for (int i = threadIdx.x; i < nq; i += blockDim.x) s_qw[i] = qw[i];
    for (int i = threadIdx.x; i < BS * S * BS * S; i += blockDim.x) s_T[i] = T[i];
    __syncthreads();
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = n_cells * nd * nd;
    if (t >= total) return;
    int64_t cell = t / (nd * nd);
    int ab = (int)(t - cell * nd * nd);
    int a = ab / nd, b = ab - a * nd;
    Geo<TDIM, GDIM> G;
    load_geo<TDIM, GDIM>(coords, cv, cell, G);
    double acc[BS][BS];
#pragma unroll
    for (int i = 0; i < BS; ++i)
#pragma unroll
        for (int k = 0; k < BS; ++k) acc[i][k] = 0.0;
    for (int q = 0; q < nq; ++q) {
        double w = s_qw[q] * G.detJ;
        if (wq) w *= wq[cell * nq + q];
        double Da[S], Db[S];
        slots<TDIM, GDIM>(s_phi, s_dphi, q, a, nd, G, Da);
        slots<TDIM, GDIM>(s_phi, s_dphi, q, b, nd, G, Db);
#pragma unroll
        for (int i = 0; i < BS; ++i)
#pragma unroll
            for (int k = 0; k < BS; ++k) {
                double s = 0.0;
#pragma unroll
                for (int j = 0; j < S; ++j)
#pragma unroll
                    for (int l = 0; l < S; ++l) s += s_T[((i * S + j) * BS + k) * S + l] * Da[j] * Db[l];
                acc[i][k] += w * s;
            }
    }
    const int ndl = nd * BS;
    double* out = Ae + cell * ndl * ndl;
#pragma unroll
    for (int i = 0; i < BS; ++i)
#pragma unroll
        for (int k = 0; k < BS; ++k) out[(a * BS + i) * ndl + (b * BS + k)] = acc[i][k];
}

template <int TDIM, int GDIM, int BS>
__global__ void __launch_bounds__(128) k_elem_linear(const double* __restrict__ coords, const int32_t* __restrict__ cv,
                                                     int64_t n_cells, int nd, int nq, const double* __restrict__ phi,
                                                     const double* __restrict__ dphi, const double* __restrict__ qw,
                                                     const double* __restrict__ wq, const double* __restrict__ L,
                                                     double* __restrict__ be) {
    extern __shared__ double sm[];
    constexpr int S = GDIM + 1;
    double* s_phi = sm;
    double* s_dphi = s_phi + nq * nd;
    double* s_qw = s_dphi + nq * nd * TDIM;
    double* s_L = s_qw + nq;
    for (int i = threadIdx.x; i < nq * nd; i += blockDim.x) s_phi[i] = phi[i];
    for (int i = threadIdx.x; i < nq * nd * TDIM; i += blockDim.x) s_dphi[i] = dphi[i];
    for (int i = threadIdx.x; i < nq; i += blockDim.x) s_qw[i] = qw[i];
    for (int i = threadIdx.x; i < BS * S; i += blockDim.x) s_L[i] = L[i];
    __syncthreads();
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_cells * nd) return;
    int64_t cell = t / nd;
    int a = (int)(t - cell * nd);
    Geo<TDIM, GDIM> G;
    load_geo<TDIM, GDIM>(coords, cv, cell, G);
    double acc[BS];
#pragma unroll
    for (int i = 0; i < BS; ++i) acc[i] = 0.0;
    for (int q = 0; q < nq; ++q) {
        double w = s_qw[q] * G.detJ;
        if (wq) w *= wq[cell * nq + q];
        double Da[S];
        slots<TDIM, GDIM>(s_phi, s_dphi, q, a, nd, G, Da);
#pragma unroll
        for (int i = 0; i < BS; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < S; ++j) s += s_L[i * S + j] * Da[j];
            acc[i] += w * s;
        }
    }
#pragma unroll
    for (int i = 0; i < BS; ++i) be[cell * nd * BS + a * BS + i] = acc[i];
}

#define ELEM_DISPATCH(KERNEL, ...)                                                    \
    do {                                                                              \
        bool ok = true;                                                               \
        if (tdim == 1 && gdim == 1 && bs == 1) KERNEL<1, 1, 1> __VA_ARGS__;           \
        else if (tdim == 2 && gdim == 2 && bs == 1) KERNEL<2, 2, 1> __VA_ARGS__;      \
        else if (tdim == 2 && gdim == 2 && bs == 2) KERNEL<2, 2, 2> __VA_ARGS__;      \
        else if (tdim == 3 && gdim == 3 && bs == 1) KERNEL<3, 3, 1> __VA_ARGS__;      \
        else if (tdim == 3 && gdim == 3 && bs == 3) KERNEL<3, 3, 3> __VA_ARGS__;      \
        else if (tdim == 1 && gdim == 2 && bs == 1) KERNEL<1, 2, 1> __VA_ARGS__;      \
        else if (tdim == 1 && gdim == 2 && bs == 2) KERNEL<1, 2, 2> __VA_ARGS__;      \
        else if (tdim == 2 && gdim == 3 && bs == 1) KERNEL<2, 3, 1> __VA_ARGS__;      \
        else if (tdim == 2 && gdim == 3 && bs == 3) KERNEL<2, 3, 3> __VA_ARGS__;      \
        else ok = false;                                                              \
        PGD_ARG(h, ok, "unsupported (tdim, gdim, bs) combination");                   \
    } while (0)

extern "C" int32_t pgd_elem_bilinear(pgd_handle_t h, const double* d_coords, const int32_t* d_cell_verts, int64_t n_cells,
                                     int32_t tdim, int32_t gdim, int32_t bs, int32_t nd, int32_t nq, const double* d_phi,
                                     const double* d_dphi, const double* d_qw, const double* d_wq, const double* d_T,
                                     double* d_Ae, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_coords && d_cell_verts && d_phi && d_dphi && d_qw && d_T && d_Ae, "null pointer");
    PGD_ARG(h, n_cells >= 0 && nd > 0 && nq > 0, "bad sizes");
    if (n_cells == 0) return 0;
    int S = gdim + 1;
    size_t smem = sizeof(double) * ((size_t)nq * nd * (1 + tdim) + nq + (size_t)bs * S * bs * S);
    PGD_ARG(h, smem <= 48 * 1024, "element tables exceed 48 KB of shared memory");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int blocks = pgd_blocks(n_cells * nd * nd, 128);
    ELEM_DISPATCH(k_elem_bilinear, <<<blocks, 128, smem, st>>>(d_coords, d_cell_verts, n_cells, nd, nq, d_phi, d_dphi, d_qw,
                                                                d_wq, d_T, d_Ae));
    PGD_LAUNCH_OK(h);
    return 0;
}

extern "C" int32_t pgd_elem_linear(pgd_handle_t h, const double* d_coords, const int32_t* d_cell_verts, int64_t n_cells,
                                   int32_t tdim, int32_t gdim, int32_t bs, int32_t nd, int32_t nq, const double* d_phi,
                                   const double* d_dphi, const double* d_qw, const double* d_wq, const double* d_L,
                                   double* d_be, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_coords && d_cell_verts && d_phi && d_dphi && d_qw && d_L && d_be, "null pointer");
    PGD_ARG(h, n_cells >= 0 && nd > 0 && nq > 0, "bad sizes");
    if (n_cells == 0) return 0;
    int S = gdim + 1;
    size_t smem = sizeof(double) * ((size_t)nq * nd * (1 + tdim) + nq + (size_t)bs * S);
    PGD_ARG(h, smem <= 48 * 1024, "element tables exceed 48 KB of shared memory");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int blocks = pgd_blocks(n_cells * nd, 128);
    ELEM_DISPATCH(k_elem_linear, <<<blocks, 128, smem, st>>>(d_coords, d_cell_verts, n_cells, nd, nq, d_phi, d_dphi, d_qw, d_wq,
                                                              d_L, d_be));
    PGD_LAUNCH_OK(h);
    return 0;
}

// ----------------------------------------------------------------------------- gather
__global__ void __launch_bounds__(256) k_gather(const double* __restrict__ src, const int64_t* __restrict__ gptr,
                                                const int32_t* __restrict__ gidx, int64_t n_out, double* __restrict__ out) {
    int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_out) return;
    int64_t k0 = gptr[g], k1 = gptr[g + 1];
    double s = 0.0;
    for (int64_t k = k0; k < k1; ++k) s += __ldg(&src[gidx[k]]);
    out[g] = s;
}

extern "C" int32_t pgd_gather_values(pgd_handle_t h, const double* d_src, const int64_t* d_gptr, const int32_t* d_gidx,
                                     int64_t n_out, double* d_out, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_src && d_gptr && d_gidx && d_out && n_out >= 0, "bad arguments");
    if (n_out == 0) return 0;
    k_gather<<<pgd_blocks(n_out, 256), 256, 0, (cudaStream_t)stream>>>(d_src, d_gptr, d_gidx, n_out, d_out);
    PGD_LAUNCH_OK(h);
    return 0;
}

// ----------------------------------------------------------------------------- fused P1 operator
struct P1Coef {
    double cm, ck, cadv[3];
};

template <int G>
__global__ void __launch_bounds__(256) k_assemble_p1(const double* __restrict__ coords, const int32_t* __restrict__ cv,
                                                     const int64_t* __restrict__ gptr, const int32_t* __restrict__ gidx,
                                                     int64_t nnz, P1Coef cf, double* __restrict__ values) {
    constexpr int NV = G + 1;
    constexpr int LPN = 4;  // lanes per nonzero
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t g = t / LPN;
    int l = (int)(t % LPN);
    double acc = 0.0;
    if (g < nnz) {
        int64_t k1 = gptr[g + 1];
        for (int64_t k = gptr[g] + l; k < k1; k += LPN) {
            int32_t c = gidx[k];
            int64_t cell = c / (NV * NV);
            int ab = c - (int)cell * (NV * NV);
            int a = ab / NV, b = ab - a * NV;  // a: test (row), b: trial (col)
            Geo<G, G> geo;
            load_geo<G, G>(coords, cv, cell, geo);
            double ga[G], gb[G];
#pragma unroll
            for (int m = 0; m < G; ++m) {
                double s0 = 0.0;
#pragma unroll
                for (int tt = 0; tt < G; ++tt) s0 -= geo.Jinv[tt][m];
                ga[m] = (a == 0) ? s0 : geo.Jinv[a - 1][m];
                gb[m] = (b == 0) ? s0 : geo.Jinv[b - 1][m];
            }
            double fact = (G == 1) ? 1.0 : (G == 2 ? 2.0 : 6.0);
            double vol = geo.detJ / fact;
            double dotg = 0.0, adv = 0.0;
#pragma unroll
            for (int m = 0; m < G; ++m) {
                dotg += ga[m] * gb[m];
                adv += cf.cadv[m] * gb[m];
            }
            double mass = vol * ((a == b) ? 2.0 : 1.0) / (double)((G + 1) * (G + 2));
            acc += cf.cm * mass + cf.ck * vol * dotg + adv * vol / (double)(G + 1);
        }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (g < nnz && l == 0) values[g] = acc;
}

extern "C" int32_t pgd_assemble_p1(pgd_handle_t h, const double* d_coords, const int32_t* d_cell_verts, int64_t n_cells,
                                   int32_t gdim, double c_mass, double c_stiff, const double* h_c_adv,
                                   const int64_t* d_gptr, const int32_t* d_gidx, int64_t nnz, double* d_values,
                                   void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_coords && d_cell_verts && d_gptr && d_gidx && d_values, "null pointer");
    PGD_ARG(h, gdim >= 1 && gdim <= 3, "gdim must be 1, 2 or 3");
    (void)n_cells;
    if (nnz <= 0) return 0;
    P1Coef cf;
    cf.cm = c_mass;
    cf.ck = c_stiff;
    for (int m = 0; m < 3; ++m) cf.cadv[m] = (h_c_adv && m < gdim) ? h_c_adv[m] : 0.0;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned int blocks = pgd_blocks(nnz * 4, 256);
    if (gdim == 1) k_assemble_p1<1><<<blocks, 256, 0, st>>>(d_coords, d_cell_verts, d_gptr, d_gidx, nnz, cf, d_values);
    else if (gdim == 2) k_assemble_p1<2><<<blocks, 256, 0, st>>>(d_coords, d_cell_verts, d_gptr, d_gidx, nnz, cf, d_values);
    else k_assemble_p1<3><<<blocks, 256, 0, st>>>(d_coords, d_cell_verts, d_gptr, d_gidx, nnz, cf, d_values);
    PGD_LAUNCH_OK(h);
    return 0;
}
