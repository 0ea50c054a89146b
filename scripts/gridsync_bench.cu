// gridsync_bench.cu -- cost of one grid-wide barrier on 148 co-resident CTAs x 256 threads: cooperative groups grid.sync()
// against a counter barrier (release fence, red.add, one polling thread per CTA, acquire fence).
#include <cstdio>
#include <cooperative_groups.h>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
typedef unsigned long long u64;

__device__ __forceinline__ void counter_barrier(u64 *ctr, u64 &epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch++;
        __threadfence();
        asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(ctr), "l"(1ull) : "memory");
        const u64 want = epoch * gridDim.x;
        u64 v;
        do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ctr) : "memory"); } while (v < want);
        __threadfence();
    }
    __syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) k(int rounds, u64 *ctr, double *data, long long *out) {
    cg::grid_group grid = cg::this_grid();
    u64 epoch = 0;
    double acc = 0.0;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; r++) {
        data[blockIdx.x * 256 + threadIdx.x] = r + acc;            // something the barrier has to make visible
        if (MODE == 0) grid.sync(); else counter_barrier(ctr, epoch);
        acc += __ldcg(data + ((blockIdx.x + 1) % gridDim.x) * 256 + threadIdx.x);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 12345.678) out[0] = 0;
}

int main() {
    u64 *ctr; double *data; long long *out;
    CK(cudaMalloc(&ctr, 256)); CK(cudaMalloc(&data, 148 * 256 * 8)); CK(cudaMalloc(&out, 148 * 8));
    const int rounds = 2000;
    for (int mode = 0; mode < 2; mode++) for (int G : {37, 148}) {
        CK(cudaMemset(ctr, 0, 256));
        int rd = rounds;
        void *args[] = {&rd, &ctr, &data, &out};
        CK(cudaLaunchCooperativeKernel(mode == 0 ? (void *)k<0> : (void *)k<1>, dim3(G), dim3(256), args, 0, 0));
        CK(cudaDeviceSynchronize());
        long long h[148]; CK(cudaMemcpy(h, out, G * 8, cudaMemcpyDeviceToHost));
        long long mx = 0; for (int i = 0; i < G; i++) mx = h[i] > mx ? h[i] : mx;
        printf("%-22s G=%3d: %7.0f cycles per barrier round (%.2f us at 1.965 GHz)\n", mode == 0 ? "cg::grid.sync()" : "counter barrier", G, mx / (double)rounds, mx / (double)rounds / 1965.0);
    }
    return 0;
}
