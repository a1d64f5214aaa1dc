// Device building blocks shared by the PCG translation units (pcg.cu, pcg_persist.cu): scalar / flag slots,
// node-block inverse, CSR entry lookup and the fused vector update  x += alpha p ; r -= alpha q ; z = M^-1 r.
#pragma once
#include "common.cuh"

enum { S_RZ_OLD = 0, S_RZ_NEW = 1, S_PQ = 2, S_RR = 3, S_BB = 4, S_TOL2 = 5, S_TMP = 8 };
enum { F_DONE = 0, F_ITER = 1, F_BAD = 2 };

template <int BS>
__device__ __forceinline__ void invert_block(const double (&B)[BS][BS], double (&I)[BS][BS]) {
    if constexpr (BS == 1) {
        I[0][0] = 1.0 / B[0][0];
    } else if constexpr (BS == 2) {
        double id = 1.0 / (B[0][0] * B[1][1] - B[0][1] * B[1][0]);
        I[0][0] = B[1][1] * id;
        I[0][1] = -B[0][1] * id;
        I[1][0] = -B[1][0] * id;
        I[1][1] = B[0][0] * id;
    } else {
        double c00 = B[1][1] * B[2][2] - B[1][2] * B[2][1];
        double c01 = B[1][2] * B[2][0] - B[1][0] * B[2][2];
        double c02 = B[1][0] * B[2][1] - B[1][1] * B[2][0];
        double id = 1.0 / (B[0][0] * c00 + B[0][1] * c01 + B[0][2] * c02);
        I[0][0] = c00 * id;
        I[1][0] = c01 * id;
        I[2][0] = c02 * id;
        I[0][1] = (B[0][2] * B[2][1] - B[0][1] * B[2][2]) * id;
        I[1][1] = (B[0][0] * B[2][2] - B[0][2] * B[2][0]) * id;
        I[2][1] = (B[0][1] * B[2][0] - B[0][0] * B[2][1]) * id;
        I[0][2] = (B[0][1] * B[1][2] - B[0][2] * B[1][1]) * id;
        I[1][2] = (B[0][2] * B[1][0] - B[0][0] * B[1][2]) * id;
        I[2][2] = (B[0][0] * B[1][1] - B[0][1] * B[1][0]) * id;
    }
}

__device__ __forceinline__ double csr_entry(const int32_t* rowptr, const int32_t* colidx, const double* vals, int r, int c) {
    int lo = rowptr[r], hi = rowptr[r + 1] - 1;
    while (lo <= hi) {
        int mid = (lo + hi) >> 1;
        int cc = colidx[mid];
        if (cc == c) return vals[mid];
        if (cc < c) lo = mid + 1;
        else hi = mid - 1;
    }
    return 0.0;
}

// x += alpha p ; r -= alpha q ; z = M^-1 r over the rows of this grid; accumulates the thread's r.z and r.r.
// Point Jacobi takes two rows per thread with 128-bit accesses when the six arrays are 16-byte aligned.
template <int BS>
__device__ __forceinline__ void pcg_update_rows(double* __restrict__ x, double* __restrict__ r, double* __restrict__ z,
                                                const double* __restrict__ p, const double* __restrict__ q,
                                                const double* __restrict__ minv, int64_t n_nodes, double alpha, double& rz,
                                                double& rr) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t gtid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool vec = false;
    if constexpr (BS == 1) {
        // point Jacobi: two rows per thread with 128-bit loads/stores when the six arrays are 16-byte aligned
        vec = (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(z) |
                 reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(minv)) & 15) == 0);
        if (vec) {
            const int64_t n2 = n_nodes >> 1;
            double2* x2 = reinterpret_cast<double2*>(x);
            double2* r2 = reinterpret_cast<double2*>(r);
            double2* z2 = reinterpret_cast<double2*>(z);
            const double2* p2 = reinterpret_cast<const double2*>(p);
            const double2* q2 = reinterpret_cast<const double2*>(q);
            const double2* m2 = reinterpret_cast<const double2*>(minv);
            double rz1 = 0.0, rr1 = 0.0;
#pragma unroll 2
            for (int64_t i = gtid; i < n2; i += stride) {
                const double2 pv = p2[i], qv = q2[i], mv = __ldg(m2 + i);
                double2 xv = x2[i], rv = r2[i];
                xv.x = fma(alpha, pv.x, xv.x);
                xv.y = fma(alpha, pv.y, xv.y);
                rv.x = fma(-alpha, qv.x, rv.x);
                rv.y = fma(-alpha, qv.y, rv.y);
                const double2 zv = make_double2(mv.x * rv.x, mv.y * rv.y);
                x2[i] = xv;
                r2[i] = rv;
                z2[i] = zv;
                rr = fma(rv.x, rv.x, rr);
                rr1 = fma(rv.y, rv.y, rr1);
                rz = fma(rv.x, zv.x, rz);
                rz1 = fma(rv.y, zv.y, rz1);
            }
            rr += rr1;
            rz += rz1;
            if ((n_nodes & 1) && gtid == 0) {
                const int64_t d = n_nodes - 1;
                x[d] = fma(alpha, p[d], x[d]);
                const double rn = fma(-alpha, q[d], r[d]);
                r[d] = rn;
                const double zi = minv[d] * rn;
                z[d] = zi;
                rr = fma(rn, rn, rr);
                rz = fma(rn, zi, rz);
            }
        }
    }
    if constexpr (BS == 1) {
        for (int64_t d = vec ? n_nodes : gtid; d < n_nodes; d += stride) {
            x[d] = fma(alpha, p[d], x[d]);
            const double rn = fma(-alpha, q[d], r[d]);
            r[d] = rn;
            const double zi = __ldg(&minv[d]) * rn;
            z[d] = zi;
            rr = fma(rn, rn, rr);
            rz = fma(rn, zi, rz);
        }
    } else {
        // node blocks: one lane per dof, a warp takes CH = (32 / BS) * BS consecutive dofs (whole nodes) so that every
        // array is read and written with unit stride; the BS residuals of a node are exchanged by shuffles.
        // (The earlier thread-per-node form read x, r, p, q with stride BS and M^-1 with stride BS*BS: 28.6 us for
        // 985 527 dofs against 8 us for the streaming direction update, ncu profiles/r02_pcg_bs3_v0.)
        constexpr int CH = (32 / BS) * BS;
        const int lane = threadIdx.x & 31;
        const int l0 = (lane / BS) * BS;
        const int64_t n = n_nodes * BS;
        const int64_t nwarps = stride >> 5;
        for (int64_t base = (gtid >> 5) * CH; base < n; base += nwarps * CH) {
            const int64_t d = base + lane;
            const bool act = lane < CH && d < n;
            double rn = 0.0;
            if (act) {
                x[d] = fma(alpha, p[d], x[d]);
                rn = fma(-alpha, q[d], r[d]);
                r[d] = rn;
            }
            double rk[BS];
#pragma unroll
            for (int k = 0; k < BS; ++k) rk[k] = __shfl_sync(0xffffffffu, rn, (l0 + k) & 31);
            if (act) {
                double zi = 0.0;
#pragma unroll
                for (int k = 0; k < BS; ++k) zi = fma(__ldg(&minv[d * BS + k]), rk[k], zi);
                z[d] = zi;
                rr = fma(rn, rn, rr);
                rz = fma(rn, zi, rz);
            }
        }
    }
}
