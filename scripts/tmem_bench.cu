// tmem_bench.cu -- tensor memory as a per-warp scratchpad for FP64 table values (no MMA involved).
// Each warp owns 24*U columns of its 32-lane quarter: unit u holds 12 doubles per lane (24 x 32-bit columns).
// 1) round trip: tcgen05.st then tcgen05.ld returns the same bits;
// 2) throughput of "load 12 doubles per lane + 12 DFMA" per unit for tables in TMEM vs in shared memory, with 8 and 16
//    warps per SM, on all 148 SMs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// load the 12 doubles of unit u of this warp
__device__ __forceinline__ void tmem_unit(uint32_t base, int u, double (&t)[12]) {
    uint32_t a[16], b[8];
    tmem_ld16(base + 24 * u, a);
    tmem_ld8(base + 24 * u + 16, b);
    tmem_wait_ld();
#pragma unroll
    for (int i = 0; i < 8; i++) t[i] = __hiloint2double((int)a[2 * i + 1], (int)a[2 * i]);
#pragma unroll
    for (int i = 0; i < 4; i++) t[8 + i] = __hiloint2double((int)b[2 * i + 1], (int)b[2 * i]);
}

template <int MODE>   // 0: TMEM, 1: shared memory
__global__ void bench(int U, int reps, long long *cyc, double *sink, int *bad) {
    extern __shared__ double stab[];            // [warps][U][12][32]
    __shared__ uint32_t s_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_base)), "n"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    // this warp's region: lane quarter (warp % 4), columns [(warp / 4) * cols_per_warp, ...)
    const int cols_per_warp = 512 / ((nw + 3) / 4);
    const uint32_t base = s_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * cols_per_warp);
    // fill: value of (warp, u, v, lane)
    for (int u = 0; u < U; u++) {
#pragma unroll
        for (int h = 0; h < 3; h++) {
            uint32_t r[8];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double x = 1.0 + 1e-3 * (warp * 1000 + u * 12 + (4 * h + i)) + 1e-7 * lane;
                r[2 * i] = (uint32_t)__double2loint(x); r[2 * i + 1] = (uint32_t)__double2hiint(x);
                stab[(((size_t)warp * U + u) * 12 + 4 * h + i) * 32 + lane] = x;
            }
            tmem_st8(base + 24 * u + 8 * h, r);
        }
    }
    tmem_wait_st();
    __syncthreads();
    // round trip check
    int nbad = 0;
    for (int u = 0; u < U; u++) {
        double t[12];
        tmem_unit(base, u, t);
#pragma unroll
        for (int v = 0; v < 12; v++) if (t[v] != stab[(((size_t)warp * U + u) * 12 + v) * 32 + lane]) nbad++;
    }
    if (nbad) atomicAdd(bad, nbad);
    __syncthreads();
    double acc[12];
#pragma unroll
    for (int v = 0; v < 12; v++) acc[v] = 0.0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
        const double f = 1.0 + 1e-9 * r;
        for (int u = 0; u < U; u++) {
            double t[12];
            if (MODE == 0) tmem_unit(base, u, t);
            else {
                const double *ts = stab + (((size_t)warp * U + u) * 12) * 32 + lane;
#pragma unroll
                for (int v = 0; v < 12; v++) t[v] = ts[v * 32];
            }
#pragma unroll
            for (int v = 0; v < 12; v++) acc[v] = fma(t[v], f, acc[v]);
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    double s = 0.0;
#pragma unroll
    for (int v = 0; v < 12; v++) s += acc[v];
    sink[blockIdx.x * blockDim.x + tid] = s;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_base), "n"(512));
}

int main() {
    long long *d_cyc; double *d_sink; int *d_bad;
    CK(cudaMalloc(&d_cyc, 148 * 8)); CK(cudaMalloc(&d_sink, 148 * 512 * 8)); CK(cudaMalloc(&d_bad, 4));
    for (int threads : {256, 512}) for (int mode = 0; mode < 2; mode++) {
        const int nw = threads / 32;
        const int U = (mode == 0) ? ((512 / ((nw + 3) / 4)) / 24 > 8 ? 8 : (512 / ((nw + 3) / 4)) / 24) : 6 * 8 / nw;
        const int Uc = (mode == 0) ? U : U;       // units per warp in this memory
        const size_t smem = (size_t)nw * Uc * 12 * 32 * 8;
        auto fn = mode == 0 ? bench<0> : bench<1>;
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaMemset(d_bad, 0, 4));
        const int reps = 200;
        fn<<<148, threads, smem>>>(Uc, reps, d_cyc, d_sink, d_bad);
        CK(cudaDeviceSynchronize());
        long long h[148]; int bad;
        CK(cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&bad, d_bad, 4, cudaMemcpyDeviceToHost));
        long long mx = 0; for (int i = 0; i < 148; i++) mx = h[i] > mx ? h[i] : mx;
        const double per_round = (double)mx / reps / Uc;      // cycles for all warps of the SM to process one unit each
        printf("%s  %2d warps/SM, %2d units/warp: %7.1f cycles per unit round (%d warps x 3 KB) = %6.1f B/clk/SM; round-trip mismatches %d\n",
               mode == 0 ? "TMEM" : "smem", nw, Uc, per_round, nw, nw * 3072.0 / per_round, bad);
    }
    return 0;
}
