// Row-owner fused P1 assembly: values = c_mass*M + c_stiff*K + sum_m c_adv[m] * int (d_m u) v written
// straight into the CSR pattern, one kernel, no element buffer, no atomics.
//
// The first fused kernel (k_assemble_p1, assemble.cu) owns one NONZERO per 4 lanes and re-derives
// the element geometry for every one of the (nv*nv) local entries of every cell: ncu showed it
// FP64/latency-bound at 1.3 % of the HBM roofline (profiles/README.md).  Here a thread owns one ROW
// (= mesh node): it walks the cells around its node in the fixed order of the vecmap, computes the
// geometry of each cell once, forms the nv entries of the local row and adds them at precomputed
// positions (one byte per entry, packed next to the cell id in the plan) into the CTA's slice of the
// value array held in shared memory; the slice is then written out fully coalesced.  The geometry is
// evaluated nv times per cell instead of nv*nv times, each thread has 3*nv independent coordinate
// loads in flight per cell, and the summation order is fixed => bitwise reproducible.
//
// Measured and rejected (round 2, profiles/README.md): a variant that first copies the coordinates of the row's ~15
// column nodes into shared memory so that a cell visit needs only its plan entry (45 instead of ~290 global loads per
// row).  360 B of shared memory per thread cap the occupancy at 18.5 % warps active and the kernel got SLOWER
// (0.869 ms against 0.746 ms on the 128^3 mesh): this kernel is issue-bound (72 % issue slots, FP64 pipe 47 %) on its
// 24 geometry evaluations per node, not on the gathers alone.
#include "common.cuh"

#define AR_ROWS 128       // rows (threads) per CTA
#define AR_CAP 3072       // doubles of shared memory for the CTA's value slice (24 KB => 9 CTAs per SM)

// ------------------------------------------------------------------------------------ plan
template <int NV>
__global__ void __launch_bounds__(256) k_p1_rowplan(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                    const int32_t* __restrict__ cd, const int64_t* __restrict__ vptr,
                                                    const int32_t* __restrict__ vidx, int64_t n_nodes, int2* __restrict__ vent,
                                                    int* flag) {
    int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_nodes) return;
    const int k0 = rowptr[row], k1 = rowptr[row + 1];
    if (k1 - k0 > 255) {
        *flag = 1;
        return;
    }
    for (int64_t e = vptr[row]; e < vptr[row + 1]; ++e) {
        const int ent = vidx[e];
        const int cell = ent / NV;
        unsigned int packed = 0;
#pragma unroll
        for (int b = 0; b < NV; ++b) {
            const int c = cd[(int64_t)cell * NV + b];
            int lo = k0, hi = k1 - 1, pos = 0;
            while (lo <= hi) {
                const int mid = (lo + hi) >> 1;
                const int cc = colidx[mid];
                if (cc == c) {
                    pos = mid - k0;
                    break;
                }
                if (cc < c) lo = mid + 1;
                else hi = mid - 1;
            }
            packed |= (unsigned int)pos << (8 * b);
        }
        vent[e] = make_int2(ent, (int)packed);
    }
}

extern "C" int32_t pgd_p1_rowplan_build_sync(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx,
                                             const int32_t* d_cell_dofs, int64_t n_cells, int32_t nv, const int64_t* d_vptr,
                                             const int32_t* d_vidx, int64_t n_nodes, int32_t* d_vent, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_rowptr && d_colidx && d_cell_dofs && d_vptr && d_vidx && d_vent && n_nodes > 0 && n_cells > 0, "bad arguments");
    PGD_ARG(h, nv >= 2 && nv <= 4, "nv must be 2, 3 or 4 (P1 interval / triangle / tetrahedron)");
    cudaStream_t st = (cudaStream_t)stream;
    int* flag = h->flags + 12;
    PGD_CUDA(h, cudaMemsetAsync(flag, 0, sizeof(int), st));
    unsigned int blocks = pgd_blocks(n_nodes, 256);
    int2* vent = reinterpret_cast<int2*>(d_vent);
    if (nv == 2) k_p1_rowplan<2><<<blocks, 256, 0, st>>>(d_rowptr, d_colidx, d_cell_dofs, d_vptr, d_vidx, n_nodes, vent, flag);
    else if (nv == 3) k_p1_rowplan<3><<<blocks, 256, 0, st>>>(d_rowptr, d_colidx, d_cell_dofs, d_vptr, d_vidx, n_nodes, vent, flag);
    else k_p1_rowplan<4><<<blocks, 256, 0, st>>>(d_rowptr, d_colidx, d_cell_dofs, d_vptr, d_vidx, n_nodes, vent, flag);
    PGD_LAUNCH_OK(h);
    int hf = 0;
    PGD_CUDA(h, pgd_fetch(h, &hf, flag, sizeof(int), nullptr, nullptr, 0, st));
    if (hf) {
        snprintf(h->err, sizeof(h->err), "pgd_p1_rowplan_build_sync: a row has more than 255 entries");
        return -4;
    }
    return 0;
}

// ------------------------------------------------------------------------------------ kernel
struct ArCoef {
    double cm, ck, cadv[3];
};

// SOA: coords is [G][n_verts] (component-major).  A warp = 32 consecutive rows whose k-th cells are
// neighbouring cells on a mesh-ordered numbering, so a warp-wide load of one coordinate component touches
// 2 cache lines instead of the 6-8 of the interleaved [n_verts][G] layout (ncu: L1TEX was the top unit).
template <int G, bool SOA, bool ADV>
__global__ void __launch_bounds__(AR_ROWS) k_assemble_p1_rows(const double* __restrict__ coords, int64_t n_verts,
                                                              const int32_t* __restrict__ cv,
                                                              const int32_t* __restrict__ rowptr,
                                                              const int64_t* __restrict__ vptr, const int2* __restrict__ vent,
                                                              int64_t n_nodes, ArCoef cf, double* __restrict__ values) {
    constexpr int NV = G + 1;
    __shared__ double s_val[AR_CAP];
    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * AR_ROWS;
    const int nr = (int)min((int64_t)AR_ROWS, n_nodes - r0);
    const int kbase = __ldg(&rowptr[r0]);
    const int kcnt = __ldg(&rowptr[r0 + nr]) - kbase;
    const bool in_smem = kcnt <= AR_CAP;  // uniform; otherwise accumulate in global memory (thread-private rows)
    double* acc = in_smem ? s_val : (values + kbase);
    for (int j = tid; j < kcnt; j += AR_ROWS) acc[j] = 0.0;
    __syncthreads();
    if (tid < nr) {
        const int64_t row = r0 + tid;
        double* arow = acc + (__ldg(&rowptr[row]) - kbase);
        const int64_t e0 = __ldg(&vptr[row]), e1 = __ldg(&vptr[row + 1]);
        constexpr double fact = (G == 1) ? 1.0 : (G == 2 ? 2.0 : 6.0);
        auto coord = [&](int vid, int g) -> double {
            return SOA ? __ldg(&coords[(int64_t)g * n_verts + vid]) : __ldg(&coords[(int64_t)vid * G + g]);
        };
        // the row's own vertex is a vertex of every cell in its list: fetch its coordinates once
        double own[G];
#pragma unroll
        for (int g = 0; g < G; ++g) own[g] = 0.0;
        if (e0 < e1) {
            const int ent0 = __ldg(&vent[e0]).x;
            const int vid = __ldg(&cv[ent0]);  // cv[cell * NV + a]
#pragma unroll
            for (int g = 0; g < G; ++g) own[g] = coord(vid, g);
        }
        // software pipeline over the cells of the row: the plan entry is fetched two cells ahead and the
        // cell's vertex ids one cell ahead, so that only the coordinate loads of the current cell are
        // on the critical path (ncu: the vent -> cell_verts -> coords chain was the top stall).
        auto load_verts = [&](int cell, int (&vv)[NV]) {
            if constexpr (NV == 4) {
                const int4 q = __ldg(reinterpret_cast<const int4*>(cv) + cell);
                vv[0] = q.x, vv[1] = q.y, vv[2] = q.z, vv[3] = q.w;
            } else if constexpr (NV == 2) {
                const int2 q = __ldg(reinterpret_cast<const int2*>(cv) + cell);
                vv[0] = q.x, vv[1] = q.y;
            } else {
#pragma unroll
                for (int v = 0; v < NV; ++v) vv[v] = __ldg(&cv[(int64_t)cell * NV + v]);
            }
        };
        int2 en1 = (e0 < e1) ? __ldg(&vent[e0]) : make_int2(0, 0);
        int2 en2 = (e0 + 1 < e1) ? __ldg(&vent[e0 + 1]) : make_int2(0, 0);
        int vi1[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) vi1[v] = 0;
        if (e0 < e1) load_verts(en1.x / NV, vi1);
        for (int64_t e = e0; e < e1; ++e) {
            const int2 en = en1;
            int vi[NV];
#pragma unroll
            for (int v = 0; v < NV; ++v) vi[v] = vi1[v];
            en1 = en2;
            if (e + 2 < e1) en2 = __ldg(&vent[e + 2]);
            if (e + 1 < e1) load_verts(en1.x / NV, vi1);
            const int cell = en.x / NV;
            const int a = en.x - cell * NV;
            const unsigned int packed = (unsigned int)en.y;
            double X[NV][G];
#pragma unroll
            for (int v = 0; v < NV; ++v)
#pragma unroll
                for (int g = 0; g < G; ++g) X[v][g] = (v == a) ? own[g] : coord(vi[v], g);
            // J[g][t] = X[t+1][g] - X[0][g];  grad phi_{t+1} = row t of J^-1, grad phi_0 = -sum_t
            double Jinv[G][G], det;
            if constexpr (G == 1) {
                const double j00 = X[1][0] - X[0][0];
                det = j00;
                Jinv[0][0] = 1.0 / j00;
            } else if constexpr (G == 2) {
                const double j00 = X[1][0] - X[0][0], j01 = X[2][0] - X[0][0];
                const double j10 = X[1][1] - X[0][1], j11 = X[2][1] - X[0][1];
                det = j00 * j11 - j01 * j10;
                const double id = 1.0 / det;
                Jinv[0][0] = j11 * id;
                Jinv[0][1] = -j01 * id;
                Jinv[1][0] = -j10 * id;
                Jinv[1][1] = j00 * id;
            } else {
                double J[3][3];
#pragma unroll
                for (int g = 0; g < 3; ++g)
#pragma unroll
                    for (int t = 0; t < 3; ++t) J[g][t] = X[t + 1][g] - X[0][g];
                const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
                const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
                const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
                det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
                const double id = 1.0 / det;
                Jinv[0][0] = c00 * id;
                Jinv[1][0] = c01 * id;
                Jinv[2][0] = c02 * id;
                Jinv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
                Jinv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
                Jinv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
                Jinv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
                Jinv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
                Jinv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
            }
            const double vol = fabs(det) / fact;
            double grad[NV][G];
#pragma unroll
            for (int m = 0; m < G; ++m) {
                double s0 = 0.0;
#pragma unroll
                for (int t = 0; t < G; ++t) {
                    s0 -= Jinv[t][m];
                    grad[t + 1][m] = Jinv[t][m];
                }
                grad[0][m] = s0;
            }
            double ga[G];
#pragma unroll
            for (int m = 0; m < G; ++m) {
                double v = grad[0][m];
#pragma unroll
                for (int t = 1; t < NV; ++t) v = (a == t) ? grad[t][m] : v;
                ga[m] = v;
            }
#pragma unroll
            for (int b = 0; b < NV; ++b) {
                double dotg = 0.0, adv = 0.0;
#pragma unroll
                for (int m = 0; m < G; ++m) {
                    dotg += ga[m] * grad[b][m];
                    if (ADV) adv += cf.cadv[m] * grad[b][m];
                }
                const double mass = vol * ((a == b) ? 2.0 : 1.0) / (double)((G + 1) * (G + 2));
                double val = cf.cm * mass + cf.ck * vol * dotg;
                if (ADV) val += adv * vol / (double)(G + 1);
                arow[(packed >> (8 * b)) & 255u] += val;
            }
        }
    }
    if (in_smem) {
        __syncthreads();
        for (int j = tid; j < kcnt; j += AR_ROWS) values[kbase + j] = s_val[j];
    }
}

extern "C" int32_t pgd_assemble_p1_rows(pgd_handle_t h, const double* d_coords, const int32_t* d_cell_verts, int64_t n_cells,
                                        int32_t gdim, double c_mass, double c_stiff, const double* h_c_adv,
                                        const int32_t* d_rowptr, const int64_t* d_vptr, const int32_t* d_vent, int64_t n_nodes,
                                        double* d_values, const double* d_coords_soa, int64_t n_verts, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, (d_coords || d_coords_soa) && d_cell_verts && d_rowptr && d_vptr && d_vent && d_values, "null pointer");
    PGD_ARG(h, gdim >= 1 && gdim <= 3, "gdim must be 1, 2 or 3");
    PGD_ARG(h, !d_coords_soa || n_verts > 0, "n_verts required with component-major coordinates");
    (void)n_cells;
    if (n_nodes <= 0) return 0;
    ArCoef cf;
    cf.cm = c_mass;
    cf.ck = c_stiff;
    bool adv = false;
    for (int m = 0; m < 3; ++m) {
        cf.cadv[m] = (h_c_adv && m < gdim) ? h_c_adv[m] : 0.0;
        adv = adv || cf.cadv[m] != 0.0;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned int blocks = pgd_blocks(n_nodes, AR_ROWS);
    const int2* vent = reinterpret_cast<const int2*>(d_vent);
    const bool soa = d_coords_soa != nullptr;
    const double* xyz = soa ? d_coords_soa : d_coords;
#define AR_LAUNCH(G, S, A) \
    k_assemble_p1_rows<G, S, A><<<blocks, AR_ROWS, 0, st>>>(xyz, n_verts, d_cell_verts, d_rowptr, d_vptr, vent, n_nodes, cf, d_values)
#define AR_DISPATCH(G)                           \
    do {                                         \
        if (soa && adv) AR_LAUNCH(G, true, true);        \
        else if (soa) AR_LAUNCH(G, true, false);         \
        else if (adv) AR_LAUNCH(G, false, true);         \
        else AR_LAUNCH(G, false, false);                 \
    } while (0)
    if (gdim == 1) AR_DISPATCH(1);
    else if (gdim == 2) AR_DISPATCH(2);
    else AR_DISPATCH(3);
#undef AR_DISPATCH
#undef AR_LAUNCH
    PGD_LAUNCH_OK(h);
    return 0;
}

// ------------------------------------------------------------------------------------ general form tensor, vector spaces
// Row-owner assembly of ANY constant-coefficient P1 atom
//     A[(a,iv),(b,iu)] = sum_cells w_cell int sum_{jv,ju} T[iv][jv][iu][ju] D_jv phi_a D_ju phi_b
// (slot 0 = value, 1+m = d/dx_m: mass, stiffness, first derivatives, Voigt elasticity, any mix) on scalar or
// node-blocked vector spaces, with an optional per-cell coefficient (degree-0 Expression: material zones), straight
// into the CSR pattern -- no element-matrix buffer, no gather pass (the generic path writes n_cells * (nv*bs)^2
// doubles and reads them back through a 10x larger gather: 1.55 GB + 2.67 GB for one scalar atom on 12.6 M tets).
// One thread per DOF ROW (node nd, component iv): it walks the cells around its node in the fixed order of the NODE
// plan (pgd_p1_rowplan_build_sync on the node-level pattern), evaluates the P1 geometry of the cell and adds the
// BS * NV entries of its row at  rowstart + BS * pos_b + iu  (pos_b = position of vertex b's node in the node's
// block-column list, one byte each in the plan) into the CTA's value slice in shared memory.  Exact affine-P1
// integrals: int phi_a phi_b = |K| (1 + d_ab) / ((g+1)(g+2)), int phi_a d_m phi_b = |K| G_bm / (g+1),
// int d_j phi_a d_m phi_b = |K| G_aj G_bm.  Fixed summation order => bitwise reproducible.
#define ART_ROWS 128
#define ART_CAP 6016  // doubles of shared memory for the CTA's value slice (47 KB)

template <int G, int BS>
struct ArTensor {
    double t[BS][G + 1][BS][G + 1];
};

template <int G, int BS>
__global__ void __launch_bounds__(ART_ROWS) k_assemble_p1_tensor(const double* __restrict__ xyz, int64_t n_verts,
                                                                 const int32_t* __restrict__ cv,
                                                                 const int32_t* __restrict__ rowptr,
                                                                 const int64_t* __restrict__ vptr, const int2* __restrict__ vent,
                                                                 const double* __restrict__ w_cell, int64_t n_rows,
                                                                 ArTensor<G, BS> T, double* __restrict__ values) {
    constexpr int NV = G + 1;
    __shared__ double s_val[ART_CAP];
    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * ART_ROWS;
    const int nr = (int)min((int64_t)ART_ROWS, n_rows - r0);
    const int kbase = __ldg(&rowptr[r0]);
    const int kcnt = __ldg(&rowptr[r0 + nr]) - kbase;
    const bool in_smem = kcnt <= ART_CAP;
    double* acc = in_smem ? s_val : (values + kbase);
    for (int j = tid; j < kcnt; j += ART_ROWS) acc[j] = 0.0;
    __syncthreads();
    if (tid < nr) {
        const int64_t row = r0 + tid;
        const int64_t nd = row / BS;
        const int iv = (int)(row - nd * BS);
        double* arow = acc + (__ldg(&rowptr[row]) - kbase);
        const int64_t e0 = __ldg(&vptr[nd]), e1 = __ldg(&vptr[nd + 1]);
        constexpr double fact = (G == 1) ? 1.0 : (G == 2 ? 2.0 : 6.0);
        for (int64_t e = e0; e < e1; ++e) {
            const int2 en = __ldg(&vent[e]);
            const int cell = en.x / NV;
            const int a = en.x - cell * NV;
            const unsigned int packed = (unsigned int)en.y;
            double X[NV][G];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int vid = __ldg(&cv[(int64_t)cell * NV + v]);
#pragma unroll
                for (int g = 0; g < G; ++g) X[v][g] = __ldg(&xyz[(int64_t)g * n_verts + vid]);
            }
            double Jinv[G][G], det;
            if constexpr (G == 1) {
                const double j00 = X[1][0] - X[0][0];
                det = j00;
                Jinv[0][0] = 1.0 / j00;
            } else if constexpr (G == 2) {
                const double j00 = X[1][0] - X[0][0], j01 = X[2][0] - X[0][0];
                const double j10 = X[1][1] - X[0][1], j11 = X[2][1] - X[0][1];
                det = j00 * j11 - j01 * j10;
                const double id = 1.0 / det;
                Jinv[0][0] = j11 * id;
                Jinv[0][1] = -j01 * id;
                Jinv[1][0] = -j10 * id;
                Jinv[1][1] = j00 * id;
            } else {
                double J[3][3];
#pragma unroll
                for (int g = 0; g < 3; ++g)
#pragma unroll
                    for (int t = 0; t < 3; ++t) J[g][t] = X[t + 1][g] - X[0][g];
                const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
                const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
                const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
                det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
                const double id = 1.0 / det;
                Jinv[0][0] = c00 * id;
                Jinv[1][0] = c01 * id;
                Jinv[2][0] = c02 * id;
                Jinv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
                Jinv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
                Jinv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
                Jinv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
                Jinv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
                Jinv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
            }
            double vol = fabs(det) / fact;
            if (w_cell) vol *= __ldg(&w_cell[cell]);
            double grad[NV][G];
#pragma unroll
            for (int m = 0; m < G; ++m) {
                double s0 = 0.0;
#pragma unroll
                for (int t = 0; t < G; ++t) {
                    s0 -= Jinv[t][m];
                    grad[t + 1][m] = Jinv[t][m];
                }
                grad[0][m] = s0;
            }
            double ga[G];
#pragma unroll
            for (int m = 0; m < G; ++m) {
                double v = grad[0][m];
#pragma unroll
                for (int t = 1; t < NV; ++t) v = (a == t) ? grad[t][m] : v;
                ga[m] = v;
            }
            // row (a, iv): t0[iu][ju] = T[iv][0][iu][ju], t1[iu][ju] = sum_{jv>=1} T[iv][jv][iu][ju] G_a[jv-1]
            double t0[BS][G + 1], t1[BS][G + 1];
#pragma unroll
            for (int iu = 0; iu < BS; ++iu)
#pragma unroll
                for (int ju = 0; ju <= G; ++ju) {
                    double s = 0.0;
#pragma unroll
                    for (int jv = 1; jv <= G; ++jv) {
                        double tv = 0.0;
#pragma unroll
                        for (int q = 0; q < BS; ++q) tv = (iv == q) ? T.t[q][jv][iu][ju] : tv;
                        s = fma(tv, ga[jv - 1], s);
                    }
                    t1[iu][ju] = s;
                    double tz = 0.0;
#pragma unroll
                    for (int q = 0; q < BS; ++q) tz = (iv == q) ? T.t[q][0][iu][ju] : tz;
                    t0[iu][ju] = tz;
                }
#pragma unroll
            for (int b = 0; b < NV; ++b) {
                const int pos = (packed >> (8 * b)) & 255u;
                const double mab = ((a == b) ? 2.0 : 1.0) / (double)((G + 1) * (G + 2));
#pragma unroll
                for (int iu = 0; iu < BS; ++iu) {
                    double s = t0[iu][0] * mab + t1[iu][0] / (double)(G + 1);
#pragma unroll
                    for (int ju = 1; ju <= G; ++ju) s += (t0[iu][ju] / (double)(G + 1) + t1[iu][ju]) * grad[b][ju - 1];
                    arow[BS * pos + iu] += vol * s;
                }
            }
        }
    }
    if (in_smem) {
        __syncthreads();
        for (int j = tid; j < kcnt; j += ART_ROWS) values[kbase + j] = s_val[j];
    }
}

template <int G, int BS>
static int32_t launch_p1_tensor(pgd_ctx* h, const double* xyz, int64_t n_verts, const int32_t* cv, const int32_t* rowptr,
                                const int64_t* vptr, const int32_t* vent, const double* w_cell, int64_t n_rows, const double* h_T,
                                double* values, cudaStream_t st) {
    ArTensor<G, BS> T;
    memcpy(&T, h_T, sizeof(T));
    k_assemble_p1_tensor<G, BS><<<pgd_blocks(n_rows, ART_ROWS), ART_ROWS, 0, st>>>(xyz, n_verts, cv, rowptr, vptr,
                                                                                  reinterpret_cast<const int2*>(vent), w_cell,
                                                                                  n_rows, T, values);
    PGD_LAUNCH_OK(h);
    return 0;
}

extern "C" int32_t pgd_assemble_p1_tensor(pgd_handle_t h, int32_t gdim, int32_t bs, const double* h_T, const double* d_coords_soa,
                                          int64_t n_verts, const int32_t* d_cell_verts, const int32_t* d_rowptr,
                                          const int64_t* d_node_vptr, const int32_t* d_node_vent, const double* d_w_cell,
                                          int64_t n_rows, double* d_values, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, h_T && d_coords_soa && d_cell_verts && d_rowptr && d_node_vptr && d_node_vent && d_values, "null pointer");
    PGD_ARG(h, gdim >= 1 && gdim <= 3 && bs >= 1 && bs <= 3 && n_verts > 0 && n_rows % bs == 0, "gdim, bs must be 1..3");
    if (n_rows <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
#define ART_CASE(G_, B_)  \
    if (gdim == G_ && bs == B_) \
        return launch_p1_tensor<G_, B_>(h, d_coords_soa, n_verts, d_cell_verts, d_rowptr, d_node_vptr, d_node_vent, d_w_cell, n_rows, \
                                        h_T, d_values, st);
    ART_CASE(1, 1) ART_CASE(1, 2) ART_CASE(1, 3) ART_CASE(2, 1) ART_CASE(2, 2) ART_CASE(2, 3) ART_CASE(3, 1) ART_CASE(3, 2)
    ART_CASE(3, 3)
#undef ART_CASE
    return -2;
}
