// Second exchange microbenchmark: all-gather of NW words per reader with memory-level parallelism.
// Reader thread t collects the CH contiguous words [t*CH, (t+1)*CH) of a buffer of NW words with CH/2 independent
// 16-byte loads in flight, re-issuing only those that still show the sentinel.  Writers: word i is owned by CTA
// (i / 3) % G  (rows dealt round-robin, 3 words per row) -- scattered writers, like the solver's row ownership.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/xchg_bench2 xchg_bench2.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

typedef unsigned long long u64;
constexpr u64 kSent = 0xFFFFFFFFFFFFFFFFull;
constexpr int kT = 512;

__device__ __forceinline__ void st_rlx(u64 *p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
template <int LD> __device__ __forceinline__ void ld2(const u64 *p, u64 &a, u64 &b) {
    if (LD == 0) asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    else if (LD == 1) asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    else asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

template <int CH, int LD>
__global__ void mlp(u64 *buf, int NW, int contiguous, int delay, int backoff, int rounds, long long *out, u64 *sink) {
    const int G = gridDim.x, c = blockIdx.x, t = threadIdx.x;
    const int bufw = ((NW + 15) / 16) * 16 + 16;
    u64 acc = 0;
    const int nrows = NW / 3;
    const long long t0 = clock64();
    for (int r = 1; r <= rounds; r++) {
        u64 *cur = buf + (size_t)(r & 3) * bufw, *clr = buf + (size_t)((r + 2) & 3) * bufw;
        // publish my rows: row k is mine if k % G == c; thread t handles my t-th row's 3 words
        {
            const int rpc = (nrows + G - 1) / G;
            const int k = contiguous ? (t < rpc ? c * rpc + t : nrows) : c + t * G;
            if (k < nrows) {
                st_rlx(clr + 3 * k, kSent); st_rlx(clr + 3 * k + 1, kSent); st_rlx(clr + 3 * k + 2, kSent);
                st_rlx(cur + 3 * k, (u64)r * 1000 + 1); st_rlx(cur + 3 * k + 1, (u64)r * 1000 + 2); st_rlx(cur + 3 * k + 2, (u64)r * 1000 + 3);
            }
        }
        const int w0 = t * CH;
        if (w0 < nrows * 3) {
            const int nvalid = min(CH, nrows * 3 - w0);     // words of this chunk that will ever be written
            u64 v[CH];
            if (delay) { const long long d0 = clock64(); while (clock64() - d0 < delay) {} }
#pragma unroll
            for (int i = 0; i < CH / 2; i++) ld2<LD>(cur + w0 + 2 * i, v[2 * i], v[2 * i + 1]);
            bool again;
            bool first = true;
            do {
                again = false;
                if (backoff && !first) __nanosleep(backoff);
                first = false;
#pragma unroll
                for (int i = 0; i < CH / 2; i++) {
                    const bool bad = (2 * i < nvalid && v[2 * i] == kSent) || (2 * i + 1 < nvalid && v[2 * i + 1] == kSent);
                    if (bad) { ld2<LD>(cur + w0 + 2 * i, v[2 * i], v[2 * i + 1]); again = true; }
                }
            } while (again);
#pragma unroll
            for (int i = 0; i < CH; i++) acc += v[i];
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    if (t == 0) out[c] = t1 - t0;
    sink[c * kT + t] = acc;
}


template <int CH, int LD>
int run(int G, int NW, int contiguous, int delay, int backoff, u64 *d_buf, long long *d_out, u64 *d_sink, size_t bufbytes) {
    const int rounds = 2000;
    long long h[256];
    for (int rep = 0; rep < 2; rep++) {
        CK(cudaMemset(d_buf, 0xFF, bufbytes));
        void *args[] = {&d_buf, &NW, &contiguous, &delay, &backoff, (void *)&rounds, &d_out, &d_sink};
        CK(cudaLaunchCooperativeKernel((void *)mlp<CH, LD>, dim3(G), dim3(kT), args, 0, 0)); CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h, d_out, G * 8, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < G; i++) mx = h[i] > mx ? h[i] : mx;
    printf("mlp CH=%2d G=%3d NW=%5d contig=%d delay=%4d backoff=%4d : %7.0f cycles/round\n", CH, G, NW, contiguous, delay, backoff, mx / (double)rounds);
    return 0;
}

int main() {
    long long *d_out; u64 *d_sink, *d_buf;
    const size_t bufbytes = 4 << 20;
    CK(cudaMalloc(&d_out, 1024 * 8)); CK(cudaMalloc(&d_sink, 148 * kT * 8)); CK(cudaMalloc(&d_buf, bufbytes));
    for (int G : {8, 148}) for (int contig : {0, 1}) for (int delay : {0, 300, 500, 700, 1000}) for (int backoff : {0, 100, 300}) {
        run<6, 0>(G, 3000, contig, delay, backoff, d_buf, d_out, d_sink, bufbytes);
    }
    for (int G : {8, 148}) for (int delay : {0, 300, 500, 700}) for (int backoff : {0, 100, 300}) {
        run<2, 0>(G, 999, 1, delay, backoff, d_buf, d_out, d_sink, bufbytes);
        run<4, 0>(G, 300, 1, delay, backoff, d_buf, d_out, d_sink, bufbytes);
    }
    return 0;
}
