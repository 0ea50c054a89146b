// FP64 latency / ILP probe (B200): cycles per DFMA for `chains` independent dependent-chains per warp and `warps` warps per SM.
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double *out, long long *cyc, int iters) {
    double a[CH];
    for (int i = 0; i < CH; i++) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) a[i] = fma(a[i], b, c);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < CH; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH> void run(double *d, long long *dc) {
    for (int warps : {1, 4, 8, 16}) {
        const int iters = 4000;
        k<CH><<<148, warps * 32>>>(d, dc, iters); cudaDeviceSynchronize();
        k<CH><<<148, warps * 32>>>(d, dc, iters); cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
        printf("chains/warp=%2d warps/SM=%2d: %.2f cycles per loop iteration (%.2f per DFMA per warp; SM rate %.1f lanes/clk)\n", CH, warps,
               (double)c / iters, (double)c / iters / CH, 32.0 * CH * warps * iters / (double)c);
    }
}
int main() {
    double *d; long long *dc; cudaMalloc(&d, 148 * 512 * 8); cudaMalloc(&dc, 8);
    run<1>(d, dc); run<2>(d, dc); run<4>(d, dc); run<8>(d, dc); run<12>(d, dc); run<16>(d, dc);
    return 0;
}
