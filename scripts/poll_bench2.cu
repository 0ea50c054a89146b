// One writer CTA, 147 polling CTAs (one per SM): distribution of store -> visible delay when the pollers started
// polling W cycles BEFORE the store.  Also with a second store later into the same line (does the first polled value
// poison later polls of the neighbouring word?).
#include <cstdio>
#include <algorithm>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
typedef unsigned long long u64;
constexpr u64 kSent = 0xFFFFFFFFFFFFFFFFull;
__device__ __forceinline__ u64 gtime() { u64 g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); return g; }
__device__ __forceinline__ void st_rlx(u64 *p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_rlx(const u64 *p) { u64 a; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory"); return a; }
__device__ __forceinline__ unsigned smid() { unsigned r; asm volatile("mov.u32 %0, %smid;" : "=r"(r)); return r; }

// word 0: written at W1, word 1 (same sector) written at W2 > W1.  Pollers poll word 0 then word 1.
__global__ void k(u64 *slots, int W1, int W2, u64 *stamps, unsigned *sm) {
    const int c = blockIdx.x;
    if (threadIdx.x != 0) return;
    sm[c] = smid();
    if (c == 0) {
        const long long t0 = clock64();
        while (clock64() - t0 < W1) {}
        stamps[0] = gtime();
        st_rlx(slots, 42);
        while (clock64() - t0 < W2) {}
        stamps[1] = gtime();
        st_rlx(slots + 1, 43);
    } else {
        while (ld_rlx(slots) == kSent) {}
        stamps[2 * c] = gtime();
        while (ld_rlx(slots + 1) == kSent) {}
        stamps[2 * c + 1] = gtime();
    }
}
int main() {
    u64 *d_slots, *d_st; unsigned *d_sm;
    CK(cudaMalloc(&d_slots, 1 << 16)); CK(cudaMalloc(&d_st, 2 * 148 * 8)); CK(cudaMalloc(&d_sm, 148 * 4));
    for (int W1 : {0, 10000}) for (int W2 : {W1 + 100, W1 + 6000, W1 + 20000}) {
        u64 h[296]; unsigned hs[148];
        CK(cudaMemset(d_slots, 0xFF, 1 << 16));
        void *args[] = {&d_slots, &W1, &W2, &d_st, &d_sm};
        CK(cudaLaunchCooperativeKernel((void *)k, dim3(148), dim3(32), args, 0, 0)); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, d_st, 296 * 8, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hs, d_sm, 148 * 4, cudaMemcpyDeviceToHost));
        long long d1[147], d2[147];
        for (int c = 1; c < 148; c++) { d1[c - 1] = (long long)(h[2 * c] - h[0]); d2[c - 1] = (long long)(h[2 * c + 1] - h[1]); }
        std::sort(d1, d1 + 147); std::sort(d2, d2 + 147);
        printf("W1=%5d W2=%5d cycles: first word visible after  min %5lld  median %5lld  p90 %5lld  max %5lld ns | second word  min %5lld  median %5lld  p90 %5lld  max %5lld ns\n",
               W1, W2, d1[0], d1[73], d1[132], d1[146], d2[0], d2[73], d2[132], d2[146]);
    }
    return 0;
}
