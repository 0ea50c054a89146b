// How long after a store does a reader that has ALREADY been polling the slot for W cycles see the new value?
// CTA 1 waits W cycles, stamps %globaltimer and stores; CTA 0 polls from the start and stamps when it sees the value.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
typedef unsigned long long u64;
constexpr u64 kSent = 0xFFFFFFFFFFFFFFFFull;
__device__ __forceinline__ u64 gtime() { u64 g; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g)); return g; }
__device__ __forceinline__ void st_rlx(u64 *p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
template <int LD> __device__ __forceinline__ u64 poll(const u64 *p) {
    u64 a, b = 0;
    if (LD == 0) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    else if (LD == 1) asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    else if (LD == 2) asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    else if (LD == 3) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    else if (LD == 4) asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    else asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    return a;
}
// slots: 32 words, one per polling lane (same line or spread by `stride` words)
template <int LD>
__global__ void k(u64 *slots, int stride, int lanes, int W, u64 *stamps) {
    if (blockIdx.x == 1) {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            while (clock64() - t0 < W) {}
            stamps[0] = gtime();
            for (int i = 0; i < lanes; i++) st_rlx(slots + 2 * i * stride, 42);
        }
    } else {
        if ((int)threadIdx.x < lanes) {
            const u64 *p = slots + 2 * threadIdx.x * stride;
            while (poll<LD>(p) == kSent) {}
            stamps[1 + threadIdx.x] = gtime();
        }
    }
}
template <int LD> int run(const char *name, u64 *d_slots, u64 *d_st) {
    for (int lanes : {1, 8, 32}) for (int stride : {1, 8}) {
        printf("%-22s lanes=%2d stride=%d words:", name, lanes, 2 * stride);
        for (int W : {0, 2000, 10000, 40000}) {
            u64 h[33];
            CK(cudaMemset(d_slots, 0xFF, 1 << 16));
            void *args[] = {&d_slots, &stride, &lanes, &W, &d_st};
            CK(cudaLaunchCooperativeKernel((void *)k<LD>, dim3(2), dim3(32), args, 0, 0)); CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, d_st, 33 * 8, cudaMemcpyDeviceToHost));
            u64 mx = 0; for (int i = 0; i < lanes; i++) mx = h[1 + i] > mx ? h[1 + i] : mx;
            printf("  W=%5d: %5lld ns", W, (long long)(mx - h[0]));
        }
        printf("\n");
    }
    return 0;
}
int main() {
    u64 *d_slots, *d_st;
    CK(cudaMalloc(&d_slots, 1 << 16)); CK(cudaMalloc(&d_st, 64 * 8));
    run<0>("ld.relaxed.gpu.u64", d_slots, d_st);
    run<1>("ld.relaxed.gpu.v2.u64", d_slots, d_st);
    run<2>("ld.volatile.u64", d_slots, d_st);
    run<3>("ld.acquire.gpu.u64", d_slots, d_st);
    run<4>("ld.cv.u64", d_slots, d_st);
    run<5>("ld.volatile.v2.u64", d_slots, d_st);
    return 0;
}
