// Exchange microbenchmark 4: the solver's all-gather (Nk rows of one double + 3 block totals per CTA) in two layouts.
//  layout 0 ("packed"):  totals [G][4] then rows [Nk] contiguous (a 128-byte line is shared by 2-4 writer CTAs)
//  layout 1 ("lines"):   CTA c owns L lines: [B1,B2,B3,spare, its rows ...]; a warp collects whole lines (coalesced
//                        16-byte loads), stages them in shared memory, every thread then reads its rows from there.
// 256 threads per CTA as in the solver.  Reports cycles per exchange (max over CTAs).
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
typedef unsigned long long u64;
constexpr u64 kSent = 0xFFFFFFFFFFFFFFFFull;
constexpr int kT = 256;
__device__ __forceinline__ void st_rlx(u64 *p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void st_rlx2(u64 *p, u64 a, u64 b) { asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory"); }
__device__ __forceinline__ void ld2(const u64 *p, u64 &a, u64 &b) { asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); }

__global__ void packed(u64 *buf, int Nk, int delay, int rounds, long long *out, u64 *sink) {
    const int G = gridDim.x, c = blockIdx.x, t = threadIdx.x;
    const int rpc = (Nk + G - 1) / G, row0 = c * rpc, nrows = max(0, min(rpc, Nk - row0));
    const int bufw = 4 * G + ((Nk + 15) / 16) * 16 + 16;
    __shared__ double sOff[512];
    u64 acc = 0;
    const long long t0 = clock64();
    for (int r = 1; r <= rounds; r++) {
        u64 *cur = buf + (size_t)(r & 3) * bufw, *clr = buf + (size_t)((r + 2) & 3) * bufw;
        if (t < nrows) st_rlx(cur + 4 * G + row0 + t, (u64)r * 1000 + t);
        if (t == 0) { st_rlx2(cur + 4 * c, r, r + 1); st_rlx2(cur + 4 * c + 2, r + 2, 0); }
        if (t < nrows) st_rlx(clr + 4 * G + row0 + t, kSent);
        if (t == 0) { st_rlx2(clr + 4 * c, kSent, kSent); st_rlx2(clr + 4 * c + 2, kSent, kSent); }
        __syncthreads();
        if (delay) { const long long d0 = clock64(); while (clock64() - d0 < delay) {} }
        const u64 *src = cur + 4 * G + 4 * t;
        u64 w[4] = {0, 0, 0, 0};
        const int nv = min(4, Nk - 4 * t);
        if (nv > 0) ld2(src, w[0], w[1]);
        if (nv > 2) ld2(src + 2, w[2], w[3]);
        bool again = true;
        while (again) {
            again = false;
            if (nv > 0 && (w[0] == kSent || (nv > 1 && w[1] == kSent))) { ld2(src, w[0], w[1]); again = true; }
            if (nv > 2 && (w[2] == kSent || (nv > 3 && w[3] == kSent))) { ld2(src + 2, w[2], w[3]); again = true; }
        }
        if (t < G) {
            u64 a, b, cc, d;
            const u64 *ts = cur + 4 * t;
            ld2(ts, a, b); ld2(ts + 2, cc, d);
            while (a == kSent || b == kSent || cc == kSent || d == kSent) { if (a == kSent || b == kSent) ld2(ts, a, b); if (cc == kSent || d == kSent) ld2(ts + 2, cc, d); }
            sOff[t] = (double)(a + b + cc);
        }
        __syncthreads();
        acc += w[0] + w[1] + w[2] + w[3] + (u64)sOff[t % G];
        __syncthreads();
    }
    const long long t1 = clock64();
    if (t == 0) out[c] = t1 - t0;
    sink[c * kT + t] = acc;
}

// lines layout, LW words per CTA (16 or 32)
template <int LW>
__global__ void lines(u64 *buf, int Nk, int delay, int rounds, long long *out, u64 *sink) {
    const int G = gridDim.x, c = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int rpc = (Nk + G - 1) / G, row0 = c * rpc, nrows = max(0, min(rpc, Nk - row0));
    const int bufw = LW * G;
    extern __shared__ u64 stage[];      // [G][LW]
    u64 acc = 0;
    const int nw = 4 + nrows;           // words this CTA publishes
    const long long t0 = clock64();
    for (int r = 1; r <= rounds; r++) {
        u64 *cur = buf + (size_t)(r & 3) * bufw, *clr = buf + (size_t)((r + 2) & 3) * bufw;
        // one warp publishes the whole line(s) with coalesced 8-byte stores
        if (warp == 0) {
            for (int i = lane; i < LW; i += 32) { st_rlx(cur + LW * c + i, (u64)r * 1000 + i); st_rlx(clr + LW * c + i, kSent); }
        }
        __syncthreads();
        if (delay) { const long long d0 = clock64(); while (clock64() - d0 < delay) {} }
        // pair p of the buffer = words 2p, 2p+1; pairs per CTA line: LW/2.  thread t collects pairs t, t+256, ...
        const int npairs = G * LW / 2;
        for (int p = t; p < npairs; p += kT) {
            const int wc = (2 * p) / LW, wi = (2 * p) % LW;          // writer CTA, word inside its line
            const int wn = 4 + max(0, min(rpc, Nk - wc * rpc));      // words that writer publishes (all LW are written here)
            u64 a, b;
            (void)wn;
            ld2(cur + 2 * p, a, b);
            while (a == kSent || b == kSent) ld2(cur + 2 * p, a, b);
            stage[2 * p] = a; stage[2 * p + 1] = b;
            (void)wi;
        }
        __syncthreads();
        for (int j = 0; j < 4; j++) {
            const int k = 4 * t + j;
            if (k < Nk) { const int oc = k / rpc; acc += stage[oc * LW + 4 + (k - oc * rpc)] + stage[oc * LW]; }
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    if (t == 0) out[c] = t1 - t0;
    sink[c * kT + t] = acc;
    (void)nw;
}

int main() {
    long long *d_out; u64 *d_sink, *d_buf;
    const size_t bufbytes = 4 << 20;
    CK(cudaMalloc(&d_out, 1024 * 8)); CK(cudaMalloc(&d_sink, 148 * kT * 8)); CK(cudaMalloc(&d_buf, bufbytes));
    long long h[256];
    const int rounds = 2000;
    CK(cudaFuncSetAttribute(lines<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 148 * 32 * 8));
    CK(cudaFuncSetAttribute(lines<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 148 * 32 * 8));
    for (int G : {8, 37, 74, 148}) for (int Nk : {100, 1000}) for (int delay : {0, 300, 600}) {
        if (Nk / G > 12 && G < 37) {}
        for (int variant = 0; variant < 3; variant++) {
            const int rpc = (Nk + G - 1) / G;
            if (variant == 1 && rpc > 12) continue;
            if (variant == 2 && rpc > 28) continue;
            for (int rep = 0; rep < 2; rep++) {
                CK(cudaMemset(d_buf, 0xFF, bufbytes));
                int nk = Nk, dl = delay, rd = rounds;
                void *args[] = {&d_buf, &nk, &dl, &rd, &d_out, &d_sink};
                if (variant == 0) CK(cudaLaunchCooperativeKernel((void *)packed, dim3(G), dim3(kT), args, 0, 0));
                else if (variant == 1) CK(cudaLaunchCooperativeKernel((void *)lines<16>, dim3(G), dim3(kT), args, (size_t)G * 16 * 8, 0));
                else CK(cudaLaunchCooperativeKernel((void *)lines<32>, dim3(G), dim3(kT), args, (size_t)G * 32 * 8, 0));
                CK(cudaDeviceSynchronize());
            }
            CK(cudaMemcpy(h, d_out, G * 8, cudaMemcpyDeviceToHost));
            long long mx = 0; for (int i = 0; i < G; i++) mx = h[i] > mx ? h[i] : mx;
            printf("%-8s G=%3d Nk=%4d delay=%3d: %7.0f cycles/exchange\n", variant == 0 ? "packed" : variant == 1 ? "lines16" : "lines32", G, Nk, delay, mx / (double)rounds);
        }
    }
    return 0;
}
