// Micro-benchmark: latency of grid-wide barrier(+reduction) variants on B200 (cooperative launch,
// one CTA per SM, 512 threads).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o barrier_bench barrier_bench.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_release(unsigned* p) { asm volatile("red.release.gpu.global.add.u32 [%0], 1;" :: "l"(p) : "memory"); }
__device__ __forceinline__ double warp_sum(double v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }

// variant 0: cg grid.sync ; 1: fence+atomicAdd+ld.acquire poll ; 2: red.release + ld.relaxed poll + fence
// variant 3: per-CTA flag slots (value + epoch), every CTA polls all slots = barrier and reduction in one
__global__ void __launch_bounds__(512, 1) k(int variant, int iters, unsigned* bar, double* vals, unsigned* flags, double* out, double* gdata) {
    cg::grid_group grid = cg::this_grid();
    const unsigned G = gridDim.x, b = blockIdx.x;
    __shared__ double s_w[32];
    __shared__ double s_tot;
    unsigned target = 0;
    double acc = 0.0;
    for (int it = 1; it <= iters; ++it) {
        // some global writes that the barrier has to publish
        gdata[(size_t)b * 512 + threadIdx.x] = it;
        double mine = 1.0 + b;
        if (variant == 0) {
            if (threadIdx.x == 0) vals[(it & 1) * G + b] = mine;
            grid.sync();
        } else if (variant == 1 || variant == 2) {
            if (threadIdx.x == 0) vals[(it & 1) * G + b] = mine;
            __syncthreads();
            target += G;
            if (threadIdx.x == 0) {
                if (variant == 1) { __threadfence(); atomicAdd(bar, 1u); while (ld_acquire(bar) < target) {} }
                else { red_release(bar); while (ld_relaxed(bar) < target) {} __threadfence(); }
            }
            __syncthreads();
        }
        if (variant <= 2) {
            if (threadIdx.x < 32) {
                double s = 0.0;
                for (unsigned i = threadIdx.x; i < G; i += 32) s += __ldcg(&vals[(it & 1) * G + i]);
                s = warp_sum(s);
                if (threadIdx.x == 0) s_tot = s;
            }
            __syncthreads();
        } else {
            __syncthreads();
            if (threadIdx.x == 0) { vals[(it & 1) * G + b] = mine; st_release(&flags[b], (unsigned)it); }
            double v = 0.0;
            if (threadIdx.x < G) {
                while (ld_relaxed(&flags[threadIdx.x]) < (unsigned)it) {}
            }
            __threadfence();
            if (threadIdx.x < G) v = __ldcg(&vals[(it & 1) * G + threadIdx.x]);
            v = warp_sum(v);
            if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
            __syncthreads();
            if (threadIdx.x == 0) { double s = 0.0; for (unsigned w = 0; w < (G + 31) / 32; ++w) s += s_w[w]; s_tot = s; }
            __syncthreads();
        }
        acc += s_tot;
    }
    if (threadIdx.x == 0 && b == 0) out[0] = acc;
}

int main() {
    int dev = 0; cudaSetDevice(dev);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned *bar, *flags; double *vals, *out, *gdata;
    cudaMalloc(&bar, 4); cudaMalloc(&flags, 4 * 1024); cudaMalloc(&vals, 8 * 2 * 1024); cudaMalloc(&out, 8); cudaMalloc(&gdata, 8 * 512 * 1024);
    const int iters = 20000;
    for (int variant = 0; variant < 4; ++variant) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(bar, 0, 4); cudaMemset(flags, 0, 4 * 1024);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            int it = iters; void* args[] = {&variant, &it, &bar, &vals, &flags, &out, &gdata};
            cudaEventRecord(e0);
            cudaError_t e = cudaLaunchCooperativeKernel((void*)k, dim3(sms), dim3(512), args, 0, 0);
            cudaEventRecord(e1); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
            if (rep == 1) printf("variant %d: %.3f us per barrier+reduce (%s) check=%.1f expected=%.1f\n", variant, 1e3 * ms / iters, cudaGetErrorString(e), h, (double)iters * (sms * (sms + 1) / 2.0));
        }
    }
    return 0;
}
