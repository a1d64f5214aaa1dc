// Micro-benchmark: issue rate of the FP64 tensor-core MMA shapes on sm_100a (register-resident operands).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_bench dmma_bench.cu && ./dmma_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE>
__global__ void __launch_bounds__(256) k(double* out, int iters) {
    double a[4] = {1.0 + threadIdx.x, 0.5, 0.25, 0.125}, b[2] = {1.0, 2.0};
    double c[8][4];
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (SHAPE == 0)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[0]), "d"(b[0]));
            else if (SHAPE == 1)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            else if (SHAPE == 2)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]),
                               "d"(b[0]), "d"(b[1]), "d"(b[0]), "d"(b[1]));
        }
    }
    double s = 0;
    for (int i = 0; i < 8; ++i)
        for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE>
void run(const char* name, double flop_per_mma, int ctas_per_sm) {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * ctas_per_sm, iters = 20000;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * 256);
    k<SHAPE><<<blocks, 256>>>(out, 100);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<SHAPE><<<blocks, 256>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = (double)blocks * 8 /*warps*/ * iters * 8.0 * flop_per_mma;
    printf("%-10s ctas/sm %d: %.3f ms  %.2f TFLOP/s  err=%s\n", name, ctas_per_sm, ms, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    for (int c = 1; c <= 4; c *= 2) {
        run<0>("m8n8k4", 2.0 * 8 * 8 * 4, c);
        run<1>("m16n8k4", 2.0 * 16 * 8 * 4, c);
        run<2>("m16n8k8", 2.0 * 16 * 8 * 8, c);
        run<3>("m16n8k16", 2.0 * 16 * 8 * 16, c);
    }
    return 0;
}
