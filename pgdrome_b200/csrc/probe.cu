// Sensor evaluation (PGD.evaluate_sensor_response, model.py:862-953; eval_fixed_modes, model.py:107-130).
//
// The reference hands the sensor coordinates to fenicstools.Probes, which locates each point in the
// spatial mesh (DOLFIN bounding-box tree -> first colliding cell) and evaluates every mode's finite
// element basis there.  On the device this is two kernels:
//
//   k_locate<G>     one thread per simplex: vertex gather (12..16 B index row + G+1 coordinate rows),
//                   then the sensor points (staged in shared memory PRB_CHUNK per pass, rank-sorted by x) whose x
//                   falls into the cell's padded bounding box are found by binary search, checked against the other
//                   axes and, only if inside, tested with their barycentric coordinates; a hit does
//                   atomicMin(cell id) for that point => the LOWEST containing cell wins, independent
//                   of scheduling (points on a facet / vertex belong to several cells).
//                   HBM-bound: 4 (G+1) n_cells + 8 G n_verts bytes per pass over the mesh.
//   k_locate_fin<G> one thread per point: barycentric coordinates in the winning cell (-1 = outside).
//
//   k_probe_modes   E[k, r] = sum_j w[r, j] X[k, dof[r, j]]: the basis weights w (Lagrange P1/P2 at the
//                   located reference point, tabulated on the host for the handful of sensors) applied to
//                   all R modes; the [R, n_rows] result is exactly the X operand pgd_eval_gemv /
//                   pgd_eval_gemm_f64 take, so the sensor response is one more GEMV.
#include <limits.h>

#include "common.cuh"

#define PRB_CHUNK 512

template <int G>
__device__ __forceinline__ bool prb_inverse(const double (&X)[G + 1][G], double (&inv)[G][G]) {
    // rows of inv map (x - X0) to the reference coordinates xi_1..xi_G
    if constexpr (G == 1) {
        const double d = X[1][0] - X[0][0];
        if (d == 0.0) return false;
        inv[0][0] = 1.0 / d;
    } else if constexpr (G == 2) {
        const double a = X[1][0] - X[0][0], b = X[2][0] - X[0][0];
        const double c = X[1][1] - X[0][1], d = X[2][1] - X[0][1];
        const double det = a * d - b * c;
        if (det == 0.0) return false;
        const double id = 1.0 / det;
        inv[0][0] = d * id;
        inv[0][1] = -b * id;
        inv[1][0] = -c * id;
        inv[1][1] = a * id;
    } else {
        double J[3][3];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) J[r][c] = X[c + 1][r] - X[0][r];
        const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
        const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
        const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
        if (det == 0.0) return false;
        const double id = 1.0 / det;
        inv[0][0] = c00 * id;
        inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id;
        inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id;
        inv[1][0] = c01 * id;
        inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id;
        inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id;
        inv[2][0] = c02 * id;
        inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id;
        inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id;
    }
    return true;
}

template <int G>
__device__ __forceinline__ double prb_bary(const double (&X0)[G], const double (&inv)[G][G], const double* p,
                                           double (&lam)[G + 1]) {
    double s = 0.0, mn;
#pragma unroll
    for (int r = 0; r < G; ++r) {
        double xi = 0.0;
#pragma unroll
        for (int c = 0; c < G; ++c) xi = fma(inv[r][c], p[c] - X0[c], xi);
        lam[r + 1] = xi;
        s += xi;
    }
    lam[0] = 1.0 - s;
    mn = lam[0];
#pragma unroll
    for (int r = 1; r <= G; ++r) mn = fmin(mn, lam[r]);
    return mn;
}

template <int G>
__global__ void __launch_bounds__(256) k_locate(const double* __restrict__ coords, const int32_t* __restrict__ cells,
                                                int64_t n_cells, const double* __restrict__ pts, int n_pts, double tol,
                                                int* __restrict__ cell_out) {
    __shared__ double s_raw[PRB_CHUNK * G];  // the chunk as given
    __shared__ double s_p[PRB_CHUNK * G];    // the chunk sorted by its first coordinate
    __shared__ double s_x[PRB_CHUNK];        // sorted first coordinates (binary-searched per cell)
    __shared__ int s_id[PRB_CHUNK];          // sorted position -> point index within the chunk
    for (int p0 = 0; p0 < n_pts; p0 += PRB_CHUNK) {
        const int np = min(PRB_CHUNK, n_pts - p0);
        __syncthreads();
        for (int i = threadIdx.x; i < np * G; i += blockDim.x) s_raw[i] = pts[(size_t)p0 * G + i];
        __syncthreads();
        // rank sort by x (np <= 512: a counting pass per point is cheaper than anything smarter); ties by index
        for (int i = threadIdx.x; i < np; i += blockDim.x) {
            const double xi = s_raw[i * G];
            int rank = 0;
            for (int j = 0; j < np; ++j) {
                const double xj = s_raw[j * G];
                rank += (xj < xi) || (xj == xi && j < i) || (xj != xj && xi == xi);  // NaNs first: never matched below
            }
            if (xi != xi) {  // rank among the NaNs themselves
                rank = 0;
                for (int j = 0; j < i; ++j) rank += (s_raw[j * G] != s_raw[j * G]);
            }
            s_x[rank] = xi;
            s_id[rank] = i;
#pragma unroll
            for (int d = 0; d < G; ++d) s_p[rank * G + d] = s_raw[i * G + d];
        }
        __syncthreads();
        // the index row of the NEXT cell of this thread is requested one iteration ahead: the vertex gathers of a
        // cell then only wait for one memory latency instead of two dependent ones (index row -> coordinates)
        auto load_row = [&](int64_t c, int32_t (&v)[G + 1]) {
            if constexpr (G == 3) {  // one 16-byte row
                const int4 q = __ldcs(reinterpret_cast<const int4*>(cells) + c);
                v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
            } else if constexpr (G == 1) {
                const int2 q = __ldcs(reinterpret_cast<const int2*>(cells) + c);
                v[0] = q.x, v[1] = q.y;
            } else {
#pragma unroll
                for (int k = 0; k <= G; ++k) v[k] = __ldcs(&cells[c * (G + 1) + k]);
            }
        };
        const int64_t cstride = (int64_t)gridDim.x * blockDim.x;
        const int64_t c_first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        int32_t vnext[G + 1];
        if (c_first < n_cells) load_row(c_first, vnext);
        for (int64_t c = c_first; c < n_cells; c += cstride) {
            double X[G + 1][G];
            double lo[G], hi[G];
            int32_t vids[G + 1];
#pragma unroll
            for (int k = 0; k <= G; ++k) vids[k] = vnext[k];
            if (c + cstride < n_cells) load_row(c + cstride, vnext);
#pragma unroll
            for (int v = 0; v <= G; ++v) {
                const int32_t vid = vids[v];
#pragma unroll
                for (int d = 0; d < G; ++d) {
                    X[v][d] = __ldg(&coords[(size_t)vid * G + d]);
                    lo[d] = v ? fmin(lo[d], X[v][d]) : X[v][d];
                    hi[d] = v ? fmax(hi[d], X[v][d]) : X[v][d];
                }
            }
            // lambda_i >= -tol admits points up to ~2 (G+1) tol * (cell extent) outside the hull: pad the box test generously
            double ext = 0.0;
#pragma unroll
            for (int d = 0; d < G; ++d) ext += hi[d] - lo[d];
            const double pad = 4.0 * (G + 1) * tol * ext + 1e-300;
#pragma unroll
            for (int d = 0; d < G; ++d) {
                lo[d] -= pad;
                hi[d] += pad;
            }
            // first sorted point with x >= lo[0]
            int a = 0, b = np;
            while (a < b) {
                const int mid = (a + b) >> 1;
                if (s_x[mid] >= lo[0]) b = mid;
                else a = mid + 1;
            }
            // candidates: x inside the padded box; then the other axes; the inverse Jacobian is only formed when needed
            double inv[G][G];
            int have_inv = 0;
            for (int i = a; i < np && s_x[i] <= hi[0]; ++i) {
                const double* p = &s_p[i * G];
                bool in = true;
#pragma unroll
                for (int d = 1; d < G; ++d) in = in && (p[d] >= lo[d]) && (p[d] <= hi[d]);
                if (!in) continue;
                if (!have_inv) have_inv = prb_inverse<G>(X, inv) ? 1 : -1;
                if (have_inv < 0) break;  // degenerate cell
                double lam[G + 1];
                if (prb_bary<G>(X[0], inv, p, lam) >= -tol) atomicMin(&cell_out[p0 + s_id[i]], (int)c);
            }
        }
    }
}

template <int G>
__global__ void k_locate_fin(const double* __restrict__ coords, const int32_t* __restrict__ cells,
                             const double* __restrict__ pts, int n_pts, int* __restrict__ cell_out,
                             double* __restrict__ bary) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pts) return;
    int c = cell_out[i];
    double lam[G + 1];
#pragma unroll
    for (int r = 0; r <= G; ++r) lam[r] = 0.0;
    if (c == INT_MAX) {
        c = -1;
    } else {
        double X[G + 1][G], inv[G][G], p[G];
#pragma unroll
        for (int v = 0; v <= G; ++v)
#pragma unroll
            for (int d = 0; d < G; ++d) X[v][d] = coords[(size_t)cells[(size_t)c * (G + 1) + v] * G + d];
#pragma unroll
        for (int d = 0; d < G; ++d) p[d] = pts[(size_t)i * G + d];
        prb_inverse<G>(X, inv);
        prb_bary<G>(X[0], inv, p, lam);
    }
    cell_out[i] = c;
#pragma unroll
    for (int r = 0; r <= G; ++r) bary[(size_t)i * (G + 1) + r] = lam[r];
}

__global__ void k_fill_int(int* p, int n, int v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

extern "C" int32_t pgd_locate_points(pgd_handle_t h, const double* d_coords, const int32_t* d_cells, int64_t n_cells,
                                     int32_t gdim, const double* d_points, int32_t n_points, double tol, int32_t* d_cell,
                                     double* d_bary, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, gdim >= 1 && gdim <= 3, "gdim must be 1, 2 or 3");
    PGD_ARG(h, n_cells >= 0 && n_points >= 0 && tol >= 0.0, "bad arguments");
    if (n_points == 0) return 0;
    PGD_ARG(h, d_points && d_cell && d_bary, "null pointer");
    PGD_ARG(h, n_cells == 0 || (d_coords && d_cells), "null pointer");
    PGD_ARG(h, gdim == 2 || (reinterpret_cast<uintptr_t>(d_cells) % (4 * (gdim + 1))) == 0,
            "d_cells must be aligned to one row (8 B for intervals, 16 B for tetrahedra)");
    cudaStream_t st = (cudaStream_t)stream;
    k_fill_int<<<pgd_blocks(n_points, 256), 256, 0, st>>>(d_cell, n_points, INT_MAX);
    PGD_LAUNCH_OK(h);
    if (n_cells > 0) {
        // persistent-style grid: a multiple of the SM count, capped by the work
        unsigned int grid = pgd_blocks(n_cells, 256);
        const unsigned int cap = (unsigned int)h->sm_count * 4u;  // one resident wave (59 registers x 256 threads)
        if (grid > cap) grid = cap;
        if (gdim == 1) k_locate<1><<<grid, 256, 0, st>>>(d_coords, d_cells, n_cells, d_points, n_points, tol, d_cell);
        else if (gdim == 2) k_locate<2><<<grid, 256, 0, st>>>(d_coords, d_cells, n_cells, d_points, n_points, tol, d_cell);
        else k_locate<3><<<grid, 256, 0, st>>>(d_coords, d_cells, n_cells, d_points, n_points, tol, d_cell);
        PGD_LAUNCH_OK(h);
    }
    const unsigned int gp = pgd_blocks(n_points, 128);
    if (gdim == 1) k_locate_fin<1><<<gp, 128, 0, st>>>(d_coords, d_cells, d_points, n_points, d_cell, d_bary);
    else if (gdim == 2) k_locate_fin<2><<<gp, 128, 0, st>>>(d_coords, d_cells, d_points, n_points, d_cell, d_bary);
    else k_locate_fin<3><<<gp, 128, 0, st>>>(d_coords, d_cells, d_points, n_points, d_cell, d_bary);
    PGD_LAUNCH_OK(h);
    return 0;
}

// E[k, r] = sum_j w[r, j] X[k, dof[r, j]]
__global__ void k_probe_modes(const double* __restrict__ X, int64_t ldx, int R, const int32_t* __restrict__ dofs,
                              const double* __restrict__ w, int64_t n_rows, int nd, double* __restrict__ E, int64_t lde) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (r >= n_rows || k >= R) return;
    double s = 0.0;
    for (int j = 0; j < nd; ++j) s = fma(w[r * nd + j], X[(size_t)k * ldx + dofs[r * nd + j]], s);
    E[(size_t)k * lde + r] = s;
}

extern "C" int32_t pgd_probe_modes(pgd_handle_t h, const double* d_X, int64_t ldx, int32_t R, const int32_t* d_dofs,
                                   const double* d_w, int64_t n_rows, int32_t nd, double* d_E, int64_t lde, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, R > 0 && R <= 65535 && nd > 0 && n_rows >= 0 && lde >= n_rows && ldx > 0, "bad arguments");
    if (n_rows == 0) return 0;
    PGD_ARG(h, d_X && d_dofs && d_w && d_E, "null pointer");
    dim3 grid(pgd_blocks(n_rows, 128), (unsigned int)R);
    k_probe_modes<<<grid, 128, 0, (cudaStream_t)stream>>>(d_X, ldx, R, d_dofs, d_w, n_rows, nd, d_E, lde);
    PGD_LAUNCH_OK(h);
    return 0;
}
