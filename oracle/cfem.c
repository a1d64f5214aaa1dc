/* TEST INFRASTRUCTURE (oracle): plain-C restatement of the three CPU pieces of the PGD hot path whose NumPy form
 * (oracle/fem.py, scipy.sparse.linalg.cg) is too slow at the sizes bench.py and the large parity tests use:
 *
 *   cfem_p1_bilinear   element assembly of one separated-form atom on P1 simplices,
 *                        A[row = test dof, col = trial dof] = int w sum T[iv,jv,iu,ju] D_jv v_iv D_ju u_iu dx
 *                      (what dolfin.assemble + the FFC tabulate_tensor behind every lhs_fct / rhs_fct callback computes,
 *                      e.g. tests/integration/test_elastic.py:71-219; same form tensor convention as oracle/fem.py) with
 *                      the exact affine-P1 integrals  int phi_a phi_b = |K| (1 + d_ab) / ((g+1)(g+2)),
 *                      int phi_a d_m phi_b = |K| G_bm / (g+1),  int d_j phi_a d_m phi_b = |K| G_aj G_bm,
 *                      summed into a given CSR pattern in cell order (deterministic);
 *   cfem_spmv          y = A x (OpenMP over rows);
 *   cfem_pcg           Jacobi / node-block-Jacobi preconditioned CG, the algorithm PETSc runs when the reference forwards
 *                      settings={"linear_solver": "cg", "preconditioner": "jacobi"} (pgdrome/solver.py:634-635), stop
 *                      ||r|| <= rtol ||b||; `max_iters` caps the work for bench.py's bounded CPU sample.
 *
 * Checked against oracle/fem.py (quadrature-based NumPy assembly) and scipy in tests/test_oracle_c.py.  Only tests/,
 * __graft_entry__ and bench.py's CPU legs load this; nothing under pgdrome_b200/ does.
 * Build: gcc -O3 -fopenmp -shared -fPIC oracle/cfem.c -o oracle/_build/libcfem.so   (oracle/cfem.py) */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int find_col(const int32_t* colidx, int lo, int hi, int c) {
    hi -= 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        int cc = colidx[mid];
        if (cc == c) return mid;
        if (cc < c) lo = mid + 1;
        else hi = mid - 1;
    }
    return -1;
}

/* inverse-transpose Jacobian gradients of the P1 basis on one simplex: G[a][m], |det|/g! in *vol; returns 0 if degenerate */
static int p1_geometry(int g, const double* X /* (g+1) x g */, double G[4][3], double* vol) {
    double J[3][3], inv[3][3], det;
    for (int i = 0; i < g; ++i)
        for (int j = 0; j < g; ++j) J[i][j] = X[(j + 1) * g + i] - X[i];  /* J[i][j] = d x_i / d xi_j */
    if (g == 1) {
        det = J[0][0];
        inv[0][0] = 1.0 / det;
    } else if (g == 2) {
        det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        inv[0][0] = J[1][1] / det;
        inv[0][1] = -J[0][1] / det;
        inv[1][0] = -J[1][0] / det;
        inv[1][1] = J[0][0] / det;
    } else {
        double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
        double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
        double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
        det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
        inv[0][0] = c00 / det;
        inv[1][0] = c01 / det;
        inv[2][0] = c02 / det;
        inv[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
        inv[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
        inv[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
        inv[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
        inv[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
        inv[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
    }
    if (det == 0.0) return 0;
    /* grad phi_a (a >= 1) = row a-1 of J^-1 (d xi_{a-1} / d x_m); grad phi_0 = - sum of the others */
    for (int m = 0; m < g; ++m) {
        double s = 0.0;
        for (int a = 1; a <= g; ++a) {
            G[a][m] = inv[a - 1][m];
            s += inv[a - 1][m];
        }
        G[0][m] = -s;
    }
    static const double fact[4] = {1.0, 1.0, 2.0, 6.0};
    *vol = fabs(det) / fact[g];
    return 1;
}

/* cells: vertex ids [n_cells, g+1] into coords [n_verts, g]; cell_nodes: node ids [n_cells, g+1] (dof = node*bs+comp);
 * w_cell: per-cell coefficient or NULL; values must be zero-initialised by the caller.  Returns the number of
 * contributions that fell outside the pattern (0 for a correct pattern). */
int64_t cfem_p1_bilinear(int32_t g, int32_t bs, int64_t n_cells, const double* coords, const int32_t* cells,
                         const int64_t* cell_nodes, const double* T, const double* w_cell, const int32_t* rowptr,
                         const int32_t* colidx, double* values) {
    const int nv = g + 1, s1 = g + 1;
    int64_t missed = 0;
    for (int64_t e = 0; e < n_cells; ++e) {
        double X[12], G[4][3], vol;
        for (int a = 0; a < nv; ++a)
            for (int m = 0; m < g; ++m) X[a * g + m] = coords[(int64_t)cells[e * nv + a] * g + m];
        if (!p1_geometry(g, X, G, &vol)) continue;
        const double w = (w_cell ? w_cell[e] : 1.0) * vol;
        for (int a = 0; a < nv; ++a)
            for (int iv = 0; iv < bs; ++iv) {
                const int64_t row = cell_nodes[e * nv + a] * bs + iv;
                for (int b = 0; b < nv; ++b)
                    for (int iu = 0; iu < bs; ++iu) {
                        /* sum_{jv,ju} T[iv,jv,iu,ju] int D_jv phi_a D_ju phi_b */
                        const double* t = T + ((int64_t)iv * s1 * bs + iu) * s1;  /* T[iv][jv][iu][ju] = t[jv*bs*s1 + ju] */
                        double acc = t[0] * ((a == b ? 2.0 : 1.0) / ((g + 1.0) * (g + 2.0)));
                        for (int m = 0; m < g; ++m) {
                            acc += t[1 + m] * (G[b][m] / (g + 1.0));                 /* v value, u derivative */
                            acc += t[(int64_t)(1 + m) * bs * s1] * (G[a][m] / (g + 1.0)); /* v derivative, u value */
                            for (int q = 0; q < g; ++q) acc += t[(int64_t)(1 + m) * bs * s1 + 1 + q] * G[a][m] * G[b][q];
                        }
                        if (acc == 0.0) continue;
                        const int64_t col = cell_nodes[e * nv + b] * bs + iu;
                        const int k = find_col(colidx, rowptr[row], rowptr[row + 1], (int)col);
                        if (k < 0) ++missed;
                        else values[k] += w * acc;
                    }
            }
    }
    return missed;
}

void cfem_spmv(int64_t n, const int32_t* rowptr, const int32_t* colidx, const double* vals, const double* x, double* y) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double s = 0.0;
        for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) s += vals[k] * x[colidx[k]];
        y[i] = s;
    }
}

static double dotp(int64_t n, const double* a, const double* b) {
    double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
    for (int64_t i = 0; i < n; ++i) s += a[i] * b[i];
    return s;
}

static int invert_small(int bs, const double* B, double* I) {
    if (bs == 1) {
        I[0] = 1.0 / B[0];
        return 1;
    }
    if (bs == 2) {
        double d = B[0] * B[3] - B[1] * B[2];
        I[0] = B[3] / d;
        I[1] = -B[1] / d;
        I[2] = -B[2] / d;
        I[3] = B[0] / d;
        return 1;
    }
    double c00 = B[4] * B[8] - B[5] * B[7], c01 = B[5] * B[6] - B[3] * B[8], c02 = B[3] * B[7] - B[4] * B[6];
    double d = B[0] * c00 + B[1] * c01 + B[2] * c02;
    I[0] = c00 / d;
    I[3] = c01 / d;
    I[6] = c02 / d;
    I[1] = (B[2] * B[7] - B[1] * B[8]) / d;
    I[4] = (B[0] * B[8] - B[2] * B[6]) / d;
    I[7] = (B[1] * B[6] - B[0] * B[7]) / d;
    I[2] = (B[1] * B[5] - B[2] * B[4]) / d;
    I[5] = (B[2] * B[3] - B[0] * B[5]) / d;
    I[8] = (B[0] * B[4] - B[1] * B[3]) / d;
    return 1;
}

/* x: initial guess on entry (zero it for a cold start), solution on return.  block = 1 (point Jacobi) or bs <= 3
 * (node-block Jacobi).  Stops when ||r|| <= rtol ||b|| or after max_iters iterations.  Returns iterations; *relres. */
int32_t cfem_pcg(int64_t n, const int32_t* rowptr, const int32_t* colidx, const double* vals, const double* b, double* x,
                 int32_t block, double rtol, int32_t max_iters, double* relres) {
    double* r = (double*)malloc(sizeof(double) * n * 4);
    double* minv = (double*)malloc(sizeof(double) * n * block);
    if (!r || !minv) {
        free(r);
        free(minv);
        return -1;
    }
    double *z = r + n, *p = z + n, *q = p + n;
    const int64_t nn = n / block;
#pragma omp parallel for schedule(static)
    for (int64_t nd = 0; nd < nn; ++nd) {
        double B[9], I[9];
        for (int i = 0; i < block; ++i)
            for (int k = 0; k < block; ++k) {
                int64_t row = nd * block + i;
                int kk = find_col(colidx, rowptr[row], rowptr[row + 1], (int)(nd * block + k));
                B[i * block + k] = kk >= 0 ? vals[kk] : 0.0;
            }
        invert_small(block, B, I);
        for (int i = 0; i < block * block; ++i) minv[nd * block * block + i] = I[i];
    }
    cfem_spmv(n, rowptr, colidx, vals, x, q);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) r[i] = b[i] - q[i];
    const double bb = dotp(n, b, b);
    const double tol2 = rtol * rtol * bb;
    double rr = dotp(n, r, r), rz = 0.0, rz_old = 1.0;
    int32_t it = 0;
    memset(p, 0, sizeof(double) * n);
    while (it < max_iters && rr > tol2 && bb > 0.0) {
#pragma omp parallel for schedule(static)
        for (int64_t nd = 0; nd < nn; ++nd)
            for (int i = 0; i < block; ++i) {
                double s = 0.0;
                for (int k = 0; k < block; ++k) s += minv[(nd * block + i) * block + k] * r[nd * block + k];
                z[nd * block + i] = s;
            }
        rz = dotp(n, r, z);
        const double beta = it == 0 ? 0.0 : rz / rz_old;
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) p[i] = z[i] + beta * p[i];
        cfem_spmv(n, rowptr, colidx, vals, p, q);
        const double alpha = rz / dotp(n, p, q);
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            x[i] += alpha * p[i];
            r[i] -= alpha * q[i];
        }
        rr = dotp(n, r, r);
        rz_old = rz;
        ++it;
    }
    if (relres) *relres = bb > 0.0 ? sqrt(rr / bb) : 0.0;
    free(r);
    free(minv);
    return it;
}

int32_t cfem_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
