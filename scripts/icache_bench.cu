// Instruction-cache capacity probe (B200, sm_100a): a loop whose straight-line body is N dependent-free FFMAs.
// Reports cycles per instruction per warp for 1 and 16 warps per SM.  A knee marks the cache level that no longer
// holds the loop body (16 bytes per instruction).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int N>
__global__ void body(float *out, long long *cyc, int iters) {
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float b = 1.000001f, c = 1e-7f;
    long long t0 = 0;
    for (int it = 0; it < iters + 1; it++) {
        if (it == 1) t0 = clock64();       // first pass warms the caches
#pragma unroll
        for (int i = 0; i < N / 8; i++) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int N>
int run(float *d, long long *dc) {
    for (int threads : {32, 512}) {
        const int iters = 50;
        body<N><<<148, threads>>>(d, dc, iters); CK(cudaDeviceSynchronize());
        body<N><<<148, threads>>>(d, dc, iters); CK(cudaDeviceSynchronize());
        long long c; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
        printf("body %6d instr (%4d KB), %3d threads/SM: %.2f cycles per instruction per warp\n", N, N * 16 / 1024, threads, (double)c / iters / N);
    }
    return 0;
}

int main() {
    float *d; long long *dc;
    CK(cudaMalloc(&d, 148 * 512 * 4)); CK(cudaMalloc(&dc, 8));
    run<256>(d, dc); run<1024>(d, dc); run<2048>(d, dc); run<4096>(d, dc); run<8192>(d, dc); run<16384>(d, dc); run<32768>(d, dc);
    return 0;
}
