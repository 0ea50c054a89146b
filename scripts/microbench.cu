// Microbenchmarks of the primitives the persistent solver is built from (B200, sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void st_relaxed(unsigned long long *p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long *p) { unsigned long long v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ double warp_sum(double v) { for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); return v; }

// 1. block-local primitives, measured by thread 0 of one CTA of 512 threads
__global__ void local_prims(long long *out, double *sink, int iters) {
    __shared__ double sm[64];
    double v = threadIdx.x * 1.0000001 + 1.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) __syncthreads();
    long long t1 = clock64();
    for (int i = 0; i < iters; i++) v = warp_sum(v) * 0.03125;
    long long t2 = clock64();
    double d = v + 3.0;
    for (int i = 0; i < iters; i++) d = 1.0 + 1.0 / d;          // dependent FP64 divides
    long long t3 = clock64();
    double a = v;
    for (int i = 0; i < iters; i++) a = fma(a, 1.0000001, 0.5); // dependent DFMA
    long long t4 = clock64();
    for (int i = 0; i < iters; i++) { if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = a; __syncthreads(); a += sm[i & 15]; __syncthreads(); }
    long long t5 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; out[4] = t5 - t4; }
    sink[threadIdx.x] = v + d + a;
}

// 2. all-to-all signalling between G CTAs (one per SM): every round each CTA stores its round number into its own
//    slot and thread t < G of every CTA spins until slot t reaches the round.  Reports cycles per round.
__global__ void all_to_all(unsigned long long *slots, int stride, long long *out, int rounds, int pollers_per_slot) {
    const int G = gridDim.x, c = blockIdx.x, t = threadIdx.x;
    __shared__ int dummy;
    long long t0 = clock64();
    for (unsigned long long r = 1; r <= (unsigned long long)rounds; r++) {
        if (t == 0) st_relaxed(slots + (size_t)c * stride, r);
        if (t < G * pollers_per_slot) { const unsigned long long *p = slots + (size_t)(t % G) * stride; while (ld_relaxed(p) < r) {} }
        __syncthreads();
    }
    long long t1 = clock64();
    if (t == 0) out[c] = t1 - t0;
    (void)dummy;
}

// 3. chain: CTA c waits for CTA c-1's value, adds, publishes (worst-case serial dependency), to expose one hop latency
__global__ void chain(unsigned long long *slots, int stride, long long *out, int rounds) {
    const int G = gridDim.x, c = blockIdx.x;
    if (threadIdx.x != 0) return;
    long long t0 = clock64();
    for (unsigned long long r = 1; r <= (unsigned long long)rounds; r++) {
        if (c > 0) { const unsigned long long *p = slots + (size_t)(c - 1) * stride; while (ld_relaxed(p) < r) {} }
        else if (r > 1) { const unsigned long long *p = slots + (size_t)(G - 1) * stride; while (ld_relaxed(p) < r - 1) {} }
        st_relaxed(slots + (size_t)c * stride, r);
    }
    long long t1 = clock64();
    out[c] = t1 - t0;
}

int main() {
    long long *d_out; double *d_sink; unsigned long long *d_slots;
    CK(cudaMalloc(&d_out, 1024 * 8)); CK(cudaMalloc(&d_sink, 1024 * 8)); CK(cudaMalloc(&d_slots, 1 << 20));
    long long h[1024];
    const int iters = 2000;
    local_prims<<<1, 512>>>(d_out, d_sink, iters); CK(cudaDeviceSynchronize());
    local_prims<<<1, 512>>>(d_out, d_sink, iters); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d_out, 5 * 8, cudaMemcpyDeviceToHost));
    printf("per-op cycles (512-thread CTA): __syncthreads %.1f | warp_sum(double)+mul %.1f | dependent DDIV(+add) %.1f | dependent DFMA %.1f | smem-broadcast round (2 syncs) %.1f\n",
           h[0] / (double)iters, h[1] / (double)iters, h[2] / (double)iters, h[3] / (double)iters, h[4] / (double)iters);
    int Gs[] = {2, 8, 32, 74, 148};
    for (int stride : {1, 16}) for (int G : Gs) for (int ppl : {1, 3}) {
        if (G * ppl > 512) continue;
        const int rounds = 2000;
        CK(cudaMemset(d_slots, 0, 1 << 20));
        void *args[] = {&d_slots, (void *)&stride, &d_out, (void *)&rounds, (void *)&ppl};
        CK(cudaLaunchCooperativeKernel((void *)all_to_all, dim3(G), dim3(512), args, 0, 0)); CK(cudaDeviceSynchronize());
        CK(cudaMemset(d_slots, 0, 1 << 20));
        CK(cudaLaunchCooperativeKernel((void *)all_to_all, dim3(G), dim3(512), args, 0, 0)); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d_out, G * 8, cudaMemcpyDeviceToHost));
        long long mx = 0; for (int i = 0; i < G; i++) mx = h[i] > mx ? h[i] : mx;
        printf("all-to-all G=%3d slot stride %2d x8B pollers/slot %d: %.0f cycles/round\n", G, stride, ppl, mx / (double)rounds);
    }
    for (int G : Gs) {
        const int rounds = 200, stride = 16;
        CK(cudaMemset(d_slots, 0, 1 << 20));
        void *args[] = {&d_slots, (void *)&stride, &d_out, (void *)&rounds};
        CK(cudaLaunchCooperativeKernel((void *)chain, dim3(G), dim3(32), args, 0, 0)); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d_out, G * 8, cudaMemcpyDeviceToHost));
        printf("chain G=%3d: %.0f cycles per hop (store -> visible to a spinning reader on another SM)\n", G, h[0] / (double)rounds / G);
    }
    return 0;
}
