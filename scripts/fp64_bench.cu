// FP64 / shared-memory throughput microbenchmark (B200, sm_100a): the denominators of the solver's roofline.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/fp64_bench fp64_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int OP>   // 0: DFMA, 1: DADD, 2: DMUL, 3: FFMA (fp32 reference)
__global__ void __launch_bounds__(512, 1) flops(double *out, int iters) {
    double a[8];
    float fa[8];
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 1e-3 + i; fa[i] = (float)a[i]; }
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) a[i] = fma(a[i], b, c);
            else if (OP == 1) a[i] = a[i] + c;
            else if (OP == 2) a[i] = a[i] * b;
            else fa[i] = fmaf(fa[i], 1.0000001f, 1e-9f);
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += a[i] + fa[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(512, 1) lds64(double *out, int iters) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double s = 0;
    int idx = threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) s += sm[(idx + i * 512) & 8191];
        idx += 32;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    double *d; CK(cudaMalloc(&d, 148 * 512 * 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    const char *names[4] = {"DFMA", "DADD", "DMUL", "FFMA"};
    for (int op = 0; op < 4; op++) {
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (op == 0) flops<0><<<148, 512>>>(d, iters); else if (op == 1) flops<1><<<148, 512>>>(d, iters);
            else if (op == 2) flops<2><<<148, 512>>>(d, iters); else flops<3><<<148, 512>>>(d, iters);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double ops = 148.0 * 512 * 8 * iters;
        printf("%s: %.3f ms, %.2f Tinstr-lanes/s (x2 flop for FMA), %.2f lanes/clk/SM at 1.965 GHz\n", names[op], ms, ops / ms / 1e9, ops / (ms * 1e-3) / 148 / 1.965e9);
    }
    for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); lds64<<<148, 512, 65536>>>(d, iters); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("LDS.64: %.3f ms, %.1f B/clk/SM at 1.965 GHz\n", ms, 148.0 * 512 * 8 * iters * 8 / (ms * 1e-3) / 148 / 1.965e9);
    return 0;
}
