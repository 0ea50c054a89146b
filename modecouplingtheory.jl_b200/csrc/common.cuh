// common.cuh -- shared host/device helpers for libmct_sm100a.so (sm_100a only, FP64 throughout).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/mct.h"

namespace mct {

void set_error(const std::string &msg);
void count_launch(int n = 1);   // every kernel launch of this library (reported by mct_launch_count)
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define MCT_CUDA(call)                                                            \
    do {                                                                          \
        cudaError_t e__ = (call);                                                 \
        if (e__ != cudaSuccess) return ::mct::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

#define MCT_REQUIRE(cond, code, msg)                \
    do {                                            \
        if (!(cond)) {                              \
            ::mct::set_error(msg);                  \
            return (code);                          \
        }                                           \
    } while (0)

#ifndef MCT_THREADS
#define MCT_THREADS 256
#endif
constexpr int kThreads = MCT_THREADS; // threads per CTA of the persistent solver (1 CTA / SM)
constexpr int kWarps = kThreads / 32;
constexpr int kMaxKPT = 8;            // wavenumbers per thread in the redundant update phase -> Nk <= 2048
constexpr int kFlagStride = 16;       // one 128-byte line per team-barrier flag (unsigned long long units)

// ---- device-side primitives ---------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// L2-coherent data accesses for buffers that other CTAs of the team write (L1 is not coherent).
__device__ __forceinline__ double ld_cg(const double *p) { return __ldcg(p); }
__device__ __forceinline__ void st_cg(double *p, double v) { __stcg(p, v); }

#endif  // __CUDACC__

}  // namespace mct
