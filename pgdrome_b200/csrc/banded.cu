// General banded LU with partial pivoting for the 1-D parameter/time dimensions (N <= a few
// thousand, bandwidth 1-2, possibly non-symmetric: int u' v, upwind D1_up).  One CTA: all threads
// scatter CSR -> LAPACK band storage (in shared memory when it fits), warp 0 runs the dgbtf2/dgbtrs
// recurrences with lanes spread over the (kl x (kl+ku)) update block, all threads un-permute.
#include "common.cuh"

__global__ void __launch_bounds__(256) k_banded_solve(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                                      const double* __restrict__ vals, const double* __restrict__ b,
                                                      double* __restrict__ x, int n, const int32_t* __restrict__ perm, int kl,
                                                      int ku, double* work, int use_smem, int* info) {
    extern __shared__ double smem[];
    const int kv = kl + ku;
    const int ldab = 2 * kl + ku + 1;
    double* AB = use_smem ? smem : work;
    double* bw = AB + (size_t)ldab * n;
    int* ipiv = (int*)(bw + n);
    int* inv = ipiv + n;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < ldab * n; i += nt) AB[i] = 0.0;
    for (int i = tid; i < n; i += nt) inv[perm[i]] = i;
    __syncthreads();
    for (int r = tid; r < n; r += nt) {
        int i = inv[r];
        bw[i] = b[r];
        for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) {
            int j = inv[colidx[k]];
            int d = i - j;
            if (d <= kl && -d <= ku) AB[(size_t)j * ldab + kv + d] = vals[k];
            else if (vals[k] != 0.0) atomicExch(info, -1);  // entry outside the declared band
        }
    }
    __syncthreads();
    if (kl <= 3 && ku <= 3) {
        // narrow bands (every 1-D P1/P2 and finite-difference operator of the path): the recurrences touch <= 4 x 7
        // entries per column, so one thread running them back to back out of shared memory beats a warp that needs
        // ~10 shuffles and 5 warp syncs per column (143 us -> ~40 us for n = 200)
        if (tid == 0) {
            int ju = 0;
            for (int j = 0; j < n; ++j) {
                const int km = min(kl, n - 1 - j);
                double* colj = AB + (size_t)j * ldab + kv;
                int jp = 0;
                double best = fabs(colj[0]);
                for (int t = 1; t <= km; ++t) {
                    const double av = fabs(colj[t]);
                    if (av > best) {
                        best = av;
                        jp = t;
                    }
                }
                ipiv[j] = j + jp;
                if (best == 0.0) {
                    atomicCAS(info, 0, j + 1);
                    continue;
                }
                ju = max(ju, min(j + ku + jp, n - 1));
                if (jp != 0) {
                    for (int c = j; c <= ju; ++c) {
                        double* pc = AB + (size_t)c * ldab + kv - (c - j);
                        const double t0 = pc[0];
                        pc[0] = pc[jp];
                        pc[jp] = t0;
                    }
                }
                const double rp = 1.0 / colj[0];
                for (int t = 1; t <= km; ++t) colj[t] *= rp;
                for (int c = j + 1; c <= ju; ++c) {
                    double* pc = AB + (size_t)c * ldab + kv - (c - j);
                    const double p0 = pc[0];
                    for (int t = 1; t <= km; ++t) pc[t] -= colj[t] * p0;
                }
            }
            for (int j = 0; j < n - 1; ++j) {
                const int lm = min(kl, n - 1 - j);
                const int l = ipiv[j];
                if (l != j) {
                    const double t0 = bw[l];
                    bw[l] = bw[j];
                    bw[j] = t0;
                }
                const double bj = bw[j];
                const double* colj = AB + (size_t)j * ldab + kv;
                for (int t = 1; t <= lm; ++t) bw[j + t] -= bj * colj[t];
            }
            for (int j = n - 1; j >= 0; --j) {
                const double* colj = AB + (size_t)j * ldab + kv;
                const double bj = bw[j] / colj[0];
                bw[j] = bj;
                const int m = min(kv, j);
                for (int t = 1; t <= m; ++t) bw[j - t] -= bj * colj[-t];
            }
        }
    } else if (tid < 32) {
        const int lane = tid;
        int ju = 0;
        for (int j = 0; j < n; ++j) {
            const int km = min(kl, n - 1 - j);
            double* colj = AB + (size_t)j * ldab + kv;
            // pivot search (first maximum, like idamax)
            double best = -1.0;
            int bi = 0;
            for (int t = lane; t <= km; t += 32) {
                double a = fabs(colj[t]);
                if (a > best) {
                    best = a;
                    bi = t;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double ob = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) {
                    best = ob;
                    bi = oi;
                }
            }
            const int jp = bi;
            if (lane == 0) ipiv[j] = j + jp;
            if (best == 0.0) {
                if (lane == 0) atomicCAS(info, 0, j + 1);
                continue;
            }
            ju = max(ju, min(j + ku + jp, n - 1));
            if (jp != 0) {
                for (int c = j + lane; c <= ju; c += 32) {
                    double* pc = AB + (size_t)c * ldab + kv - (c - j);
                    double t0 = pc[0], t1 = pc[jp];
                    pc[0] = t1;
                    pc[jp] = t0;
                }
            }
            __syncwarp();
            const double rp = 1.0 / colj[0];
            __syncwarp();
            for (int t = 1 + lane; t <= km; t += 32) colj[t] *= rp;
            __syncwarp();
            const int ncol = ju - j;
            const int total = ncol * km;
            for (int idx = lane; idx < total; idx += 32) {
                int cp = idx / km + 1;
                int t = idx - (cp - 1) * km + 1;
                double* pc = AB + (size_t)(j + cp) * ldab + kv - cp;
                pc[t] -= colj[t] * pc[0];
            }
            __syncwarp();
        }
        // forward substitution with row interchanges
        for (int j = 0; j < n - 1; ++j) {
            const int lm = min(kl, n - 1 - j);
            const int l = ipiv[j];
            if (lane == 0 && l != j) {
                double t0 = bw[l];
                bw[l] = bw[j];
                bw[j] = t0;
            }
            __syncwarp();
            const double bj = bw[j];
            const double* colj = AB + (size_t)j * ldab + kv;
            for (int t = 1 + lane; t <= lm; t += 32) bw[j + t] -= bj * colj[t];
            __syncwarp();
        }
        // back substitution, kv super-diagonals
        for (int j = n - 1; j >= 0; --j) {
            const double* colj = AB + (size_t)j * ldab + kv;
            double bj = bw[j] / colj[0];
            __syncwarp();
            if (lane == 0) bw[j] = bj;
            const int m = min(kv, j);
            for (int t = 1 + lane; t <= m; t += 32) bw[j - t] -= bj * colj[-t];
            __syncwarp();
        }
    }
    __syncthreads();
    for (int i = tid; i < n; i += nt) x[perm[i]] = bw[i];
}

extern "C" int32_t pgd_banded_solve(pgd_handle_t h, const int32_t* d_rowptr, const int32_t* d_colidx, const double* d_values,
                                    const double* d_b, double* d_x, int32_t n, const int32_t* d_perm, int32_t kl, int32_t ku,
                                    double* d_work, int32_t* d_info, void* stream) {
    PGD_CHECK_HANDLE(h);
    PGD_ARG(h, d_rowptr && d_colidx && d_values && d_b && d_x && d_perm && d_work && d_info, "null pointer");
    PGD_ARG(h, n > 0 && kl >= 0 && ku >= 0 && kl <= 64 && ku <= 64, "bad sizes (bandwidths must be <= 64)");
    cudaStream_t st = (cudaStream_t)stream;
    PGD_CUDA(h, cudaMemsetAsync(d_info, 0, sizeof(int32_t), st));
    size_t ldab = 2 * (size_t)kl + ku + 1;
    size_t bytes = sizeof(double) * (ldab + 1) * n + sizeof(int) * 2 * (size_t)n;
    int use_smem = bytes <= 200 * 1024;
    if (use_smem && bytes > 48 * 1024)
        PGD_CUDA(h, cudaFuncSetAttribute(k_banded_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    k_banded_solve<<<1, 256, use_smem ? bytes : 0, st>>>(d_rowptr, d_colidx, d_values, d_b, d_x, n, d_perm, kl, ku, d_work,
                                                         use_smem, d_info);
    PGD_LAUNCH_OK(h);
    return 0;
}
