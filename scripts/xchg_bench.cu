// Microbenchmark of all-gather protocols between G co-resident CTAs through L2 (B200, sm_100a).
// Every round each CTA publishes W doubles and every CTA collects all G*W doubles.  Reports cycles / round.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/xchg_bench xchg_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

typedef unsigned long long u64;
constexpr u64 kSent = 0xFFFFFFFFFFFFFFFFull;
constexpr int kT = 512;

__device__ __forceinline__ void st_rlx(u64 *p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_rlx(const u64 *p) { u64 v; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void ld_rlx2(const u64 *p, u64 &a, u64 &b) { asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); }
__device__ __forceinline__ void st_rlx2(u64 *p, u64 a, u64 b) { asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory"); }
__device__ __forceinline__ void st_rel(u64 *p, u64 v) { asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld_acq(const u64 *p) { u64 v; asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }

// mode 0: pull, 8-byte sentinel slots.  Writer c owns LW = ceil(W/16) lines; 4 rotating buffers.
// mode 1: same with 16-byte loads (W even).
// stagger: reader c starts at writer (c*stag) to spread simultaneous polls.
__global__ void pull(u64 *buf, int W, int mode, int stag, int rounds, long long *out, u64 *sink) {
    const int G = gridDim.x, c = blockIdx.x, t = threadIdx.x;
    const int LW = (W + 15) / 16, wpc = LW * 16;          // words per CTA region
    const int bufw = G * wpc;
    u64 acc = 0;
    const long long t0 = clock64();
    for (int r = 1; r <= rounds; r++) {
        u64 *cur = buf + (size_t)(r & 3) * bufw, *clr = buf + (size_t)((r + 2) & 3) * bufw;
        if (t < W) { st_rlx(clr + c * wpc + t, kSent); st_rlx(cur + c * wpc + t, (u64)r * 1000 + t); }
        if (mode == 0) {
            for (int i = t; i < G * W; i += kT) {
                int w = i / W, j = i - w * W;
                w = (w + c * stag) % G;
                const u64 *p = cur + w * wpc + j;
                u64 v;
                while ((v = ld_rlx(p)) == kSent) {}
                acc += v;
            }
        } else {
            const int W2 = W / 2;
            for (int i = t; i < G * W2; i += kT) {
                int w = i / W2, j = i - w * W2;
                w = (w + c * stag) % G;
                const u64 *p = cur + w * wpc + 2 * j;
                u64 a, b;
                do { ld_rlx2(p, a, b); } while (a == kSent || b == kSent);
                acc += a + b;
            }
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    if (t == 0) out[c] = t1 - t0;
    sink[c * kT + t] = acc;
}

// push: writer c stores its W words into every reader's mailbox  mb[reader][c][W]  (reader-private lines).
__global__ void push(u64 *buf, int W, int rounds, long long *out, u64 *sink) {
    const int G = gridDim.x, c = blockIdx.x, t = threadIdx.x;
    const int mbw = ((G * W + 15) / 16) * 16;      // words per mailbox
    const size_t bufw = (size_t)G * mbw;
    u64 acc = 0;
    const long long t0 = clock64();
    for (int r = 1; r <= rounds; r++) {
        u64 *cur = buf + (size_t)(r & 3) * bufw, *clr = buf + (size_t)((r + 2) & 3) * bufw;
        // re-arm my own mailbox two rounds ahead (only I read it; writers have not started that round)
        for (int i = t; i < G * W; i += kT) st_rlx(clr + (size_t)c * mbw + i, kSent);
        for (int i = t; i < G * W; i += kT) {
            int rd = i / W, j = i - rd * W;
            rd = (rd + c) % G;
            st_rlx(cur + (size_t)rd * mbw + c * W + j, (u64)r * 1000 + j);
        }
        for (int i = t; i < G * W; i += kT) {
            const u64 *p = cur + (size_t)c * mbw + i;
            u64 v;
            while ((v = ld_rlx(p)) == kSent) {}
            acc += v;
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    if (t == 0) out[c] = t1 - t0;
    sink[c * kT + t] = acc;
}

// flag: data stores, fence, one release flag per writer; readers poll G flags, then read data with ld.cg.
__global__ void flag(u64 *buf, u64 *flags, int W, int rounds, long long *out, u64 *sink) {
    const int G = gridDim.x, c = blockIdx.x, t = threadIdx.x;
    const int LW = (W + 15) / 16, wpc = LW * 16;
    const int bufw = G * wpc;
    u64 acc = 0;
    const long long t0 = clock64();
    for (int r = 1; r <= rounds; r++) {
        u64 *cur = buf + (size_t)(r & 1) * bufw;
        if (t < W) cur[c * wpc + t] = (u64)r * 1000 + t;
        if (t < 32) {      // W <= 32: the writing warp fences and releases
            __syncwarp();
            if (t == 0) { __threadfence(); st_rel(flags + c * 16, (u64)r); }
        }
        if (t < G) { while (ld_acq(flags + t * 16) < (u64)r) {} }
        __syncthreads();
        for (int i = t; i < G * W; i += kT) {
            int w = i / W, j = i - w * W;
            acc += __ldcg((const unsigned long long *)(cur + w * wpc + j));
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    if (t == 0) out[c] = t1 - t0;
    sink[c * kT + t] = acc;
}

int main() {
    long long *d_out; u64 *d_sink, *d_buf, *d_flags;
    const size_t bufbytes = 64 << 20;
    CK(cudaMalloc(&d_out, 1024 * 8)); CK(cudaMalloc(&d_sink, 148 * kT * 8)); CK(cudaMalloc(&d_buf, bufbytes)); CK(cudaMalloc(&d_flags, 148 * 128));
    long long h[1024];
    const int rounds = 3000;
    int Gs[] = {4, 8, 16, 37, 74, 148};
    auto report = [&](const char *name, int G, int W, int extra) {
        cudaMemcpy(h, d_out, G * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < G; i++) mx = h[i] > mx ? h[i] : mx;
        printf("%-10s G=%3d W=%2d x=%d : %7.0f cycles/round\n", name, G, W, extra, mx / (double)rounds);
    };
    for (int G : Gs) for (int W : {2, 8, 22}) {
        for (int mode : {0, 1}) for (int stag : {0, 1}) {
            for (int rep = 0; rep < 2; rep++) {
                CK(cudaMemset(d_buf, 0xFF, bufbytes));
                void *args[] = {&d_buf, &W, &mode, &stag, (void *)&rounds, &d_out, &d_sink};
                CK(cudaLaunchCooperativeKernel((void *)pull, dim3(G), dim3(kT), args, 0, 0)); CK(cudaDeviceSynchronize());
            }
            report(mode ? "pull16" : "pull8", G, W, stag);
        }
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaMemset(d_buf, 0xFF, bufbytes));
            void *args[] = {&d_buf, &W, (void *)&rounds, &d_out, &d_sink};
            CK(cudaLaunchCooperativeKernel((void *)push, dim3(G), dim3(kT), args, 0, 0)); CK(cudaDeviceSynchronize());
        }
        report("push", G, W, 0);
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaMemset(d_buf, 0, bufbytes)); CK(cudaMemset(d_flags, 0, 148 * 128));
            void *args[] = {&d_buf, &d_flags, &W, (void *)&rounds, &d_out, &d_sink};
            CK(cudaLaunchCooperativeKernel((void *)flag, dim3(G), dim3(kT), args, 0, 0)); CK(cudaDeviceSynchronize());
        }
        report("flag", G, W, 0);
    }
    return 0;
}
