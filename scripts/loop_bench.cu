// loop_bench.cu -- the solver's fixed-point loop reduced to its communication skeleton: G co-resident CTAs (one per SM,
// 256 threads) alternate a compute phase of W +- jitter cycles with an all-gather of Nk doubles + 3 block totals per CTA.
// Protocol variants of the all-gather are timed under realistic arrival skew (round 2; the successor of xchg_bench4).
//   proto 0  counter:   publish, red.add on one counter, one thread polls it, named barrier, one-shot loads (round-1 protocol)
//   proto 1  direct:    every thread polls its own words (xchg_bench4 "packed")
//   proto 2  totals:    warp 0 polls the block totals (coalesced 16-byte loads), named barrier, one-shot row loads
//   proto 3  stage:     warp 0 polls totals AND rows (coalesced), stages them in shared memory, the others read them there
//   proto 4  stage8:    like 3 but all 8 warps poll a slice each
// Options: R replicas of the buffers (readers of CTA c use replica c % R); publish with st.relaxed or with red.add
// (the slot holds the sentinel -1, adding v+1 leaves v).
// Every variant checks what it received (sum over all words of every round) against the closed form.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
typedef unsigned long long u64;
constexpr u64 kSent = 0xFFFFFFFFFFFFFFFFull;
constexpr int kT = 256;
__device__ __forceinline__ void st_rlx(u64 *p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void st_rlx2(u64 *p, u64 a, u64 b) { asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory"); }
__device__ __forceinline__ void red_add(u64 *p, u64 v) { asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ u64 ld1(const u64 *p) { u64 a; asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory"); return a; }
__device__ __forceinline__ void ld2(const u64 *p, u64 &a, u64 &b) { asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory"); }
__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

struct Prm { u64 *buf; int Nk, W, J, S, R, rounds, use_red, delay, backoff, adaptive; long long *out; u64 *sink; u64 *ctr; };

__device__ __forceinline__ u64 rowval(int r, int k) { return (u64)r * 4096 + k + 1; }
__device__ __forceinline__ u64 totval(int r, int c, int m) { return (u64)r * 1024 + c * 4 + m + 7; }

template <int PROTO>
__global__ void __launch_bounds__(kT, 1) loop(Prm P) {
    const int G = gridDim.x, c = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int Nk = P.Nk;
    const int rpc = (Nk + G - 1) / G, row0 = min(Nk, c * rpc), nrows = max(0, min(rpc, Nk - row0));
    const int bufw = 4 * G + ((Nk + 15) / 16) * 16 + 16;          // words of one buffer: totals [G][4], rows [Nk]
    const size_t rstride = 4 * (size_t)bufw;
    extern __shared__ u64 stage[];                                  // [bufw] staging (proto 3, 4)
    __shared__ double sOff[512];
    u64 acc = 0;
    int adelay = P.delay, nretry = 0;
    const int KPT = (Nk + kT - 1) / kT;
    const long long t0 = clock64();
    for (int r = 1; r <= P.rounds; r++) {
        // ---- compute phase: W cycles +- jitter (per CTA and round), all warps busy-wait
        {
            const int w = P.W - (int)(hash(c * 31u + 5u) % (unsigned)(P.S + 1)) + (P.J ? (int)(hash(c * 7919u + r * 104729u) % (unsigned)(2 * P.J + 1)) - P.J : 0);
            const long long d0 = clock64();
            while (clock64() - d0 < w) {}
            __syncthreads();
        }
        const size_t ocur = (size_t)(r & 3) * bufw, oclr = (size_t)((r + 2) & 3) * bufw;
        // ---- publish into every replica
        if (t < nrows) {
            u64 *b = P.buf + 4 * G + row0 + t;
            const u64 v = rowval(r, row0 + t);
            for (int q = 0; q < P.R; q++, b += rstride) {
                if (P.use_red) red_add(b + ocur, v + 1); else st_rlx(b + ocur, v);
                st_rlx(b + oclr, kSent);
            }
        }
        if (t == 32) {      // another warp publishes the totals
            u64 *b = P.buf + 4 * c;
            for (int q = 0; q < P.R; q++, b += rstride) {
                if (P.use_red) { red_add(b + ocur, totval(r, c, 0) + 1); red_add(b + ocur + 1, totval(r, c, 1) + 1); red_add(b + ocur + 2, totval(r, c, 2) + 1); red_add(b + ocur + 3, 1); }
                else { st_rlx2(b + ocur, totval(r, c, 0), totval(r, c, 1)); st_rlx2(b + ocur + 2, totval(r, c, 2), 0); }
                st_rlx2(b + oclr, kSent, kSent); st_rlx2(b + oclr + 2, kSent, kSent);
            }
        }
        const u64 *cur = P.buf + (size_t)(c % P.R) * rstride + ocur;
        u64 got = 0;
        int retried = 0;
        if (PROTO == 0) {
            if (warp == 0) {
                __syncwarp();
                if (lane == 0) {
                    red_add(P.ctr, 1);
                    const u64 want = (u64)G * r;
                    while (ld1(P.ctr) < want) {}
                }
                __syncwarp();
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kT) : "memory");
            {
                u64 w[4] = {0, 0, 0, 0}, a = 0, b = 0, cc = 0, d = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) { const int k = j * kT + t; if (k < Nk) w[j] = ld1(cur + 4 * G + k); }
                if (t < G) { ld2(cur + 4 * t, a, b); ld2(cur + 4 * t + 2, cc, d); }
                while (true) {
                    bool again = false;
#pragma unroll
                    for (int j = 0; j < 4; j++) if (w[j] == kSent) { w[j] = ld1(cur + 4 * G + j * kT + t); again = true; }
                    if (a == kSent || b == kSent) { ld2(cur + 4 * t, a, b); again = true; }
                    if (cc == kSent || d == kSent) { ld2(cur + 4 * t + 2, cc, d); again = true; }
                    if (!again) break;
                    if (P.backoff) { const long long d0 = clock64(); while (clock64() - d0 < P.backoff) {} }
                }
                got += w[0] + w[1] + w[2] + w[3] + a + b + cc;
            }
        } else if (PROTO == 1) {
            const long long tp = clock64();
            { const int dl = P.adaptive ? adelay : P.delay; if (dl > 0) { while (clock64() - tp < dl) {} } }
            {
                u64 w[4] = {0, 0, 0, 0}, a = 0, b = 0, cc = 0, d = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) { const int k = j * kT + t; if (k < Nk) w[j] = ld1(cur + 4 * G + k); }
                if (t < G) { ld2(cur + 4 * t, a, b); ld2(cur + 4 * t + 2, cc, d); }
                while (true) {
                    bool again = false;
#pragma unroll
                    for (int j = 0; j < 4; j++) if (w[j] == kSent) { w[j] = ld1(cur + 4 * G + j * kT + t); again = true; }
                    if (a == kSent || b == kSent) { ld2(cur + 4 * t, a, b); again = true; }
                    if (cc == kSent || d == kSent) { ld2(cur + 4 * t + 2, cc, d); again = true; }
                    if (!again) break;
                    retried = 1;
                    if (P.backoff) { const long long d0 = clock64(); while (clock64() - d0 < P.backoff) {} }
                }
                got += w[0] + w[1] + w[2] + w[3] + a + b + cc;
            }
        } else if (PROTO == 2) {
            if (warp == 0) {
                // pieces of 16 bytes: p = i*32 + lane < 2G
                u64 a[10], b[10];
                unsigned pend = 0;
#pragma unroll
                for (int i = 0; i < 10; i++) { const int p = i * 32 + lane; a[i] = b[i] = 0; if (p < 2 * G) { ld2(cur + 2 * p, a[i], b[i]); pend |= 1u << i; } }
                while (true) {
                    bool again = false;
#pragma unroll
                    for (int i = 0; i < 10; i++) if ((pend >> i) & 1) { if (a[i] == kSent || b[i] == kSent) { ld2(cur + 2 * (i * 32 + lane), a[i], b[i]); again = true; } else pend &= ~(1u << i); }
                    if (!__any_sync(0xffffffffu, again)) break;
                }
#pragma unroll
                for (int i = 0; i < 10; i++) { const int p = i * 32 + lane; if (p < 2 * G) got += a[i] + ((p & 1) ? 0 : b[i]); }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kT) : "memory");
            for (int j = 0; j < KPT; j++) { const int k = j * kT + t; if (k < Nk) { u64 w = ld1(cur + 4 * G + k); while (w == kSent) w = ld1(cur + 4 * G + k); got += w; } }
        } else {
            // poll whole buffer (totals + rows) in 16-byte pieces, coalesced, by 1 warp (proto 3) or 8 warps (proto 4); stage in smem
            const int npieces = (4 * G + Nk + 1) / 2;
            const int nth = (PROTO == 3) ? 32 : kT;
            if (t < nth) {
                constexpr int MAXI = (PROTO == 3) ? 26 : 4;       // pieces per thread (Nk <= 1000+, G <= 148: 796 pieces)
                u64 a[MAXI], b[MAXI];
                unsigned pend = 0;
#pragma unroll
                for (int i = 0; i < MAXI; i++) { const int p = i * nth + t; a[i] = b[i] = 0; if (p < npieces) { ld2(cur + 2 * p, a[i], b[i]); pend |= 1u << i; } }
                while (pend) {
#pragma unroll
                    for (int i = 0; i < MAXI; i++) if ((pend >> i) & 1) {
                        const int p = i * nth + t;
                        const bool last_odd = (2 * p + 1 >= 4 * G + Nk);
                        if (a[i] == kSent || (!last_odd && b[i] == kSent)) ld2(cur + 2 * p, a[i], b[i]);
                        else { stage[2 * p] = a[i]; stage[2 * p + 1] = b[i]; pend &= ~(1u << i); }
                    }
                }
            }
            __syncthreads();
            for (int j = 0; j < KPT; j++) { const int k = j * kT + t; if (k < Nk) got += stage[4 * G + k]; }
            if (t < G) got += stage[4 * t] + stage[4 * t + 1] + stage[4 * t + 2];
        }
        acc += got;
        if (t < G) sOff[t] = (double)got;
        const int anyr = __syncthreads_or(retried);
        if (P.adaptive) adelay = anyr ? min(adelay + P.adaptive, 4000) : max(adelay - P.adaptive / 2, 0);
        nretry += anyr;
    }
    const long long t1 = clock64();
    if (t == 0) { P.out[c] = t1 - t0; P.out[256 + c] = nretry; P.out[512 + c] = adelay; }
    P.sink[c * kT + t] = acc;
}

int main(int argc, char **argv) {
    const int G = 148, Nk = 1000, rounds = 3000;
    long long *d_out; u64 *d_sink, *d_buf, *d_ctr;
    const size_t bufbytes = 64 << 20;
    CK(cudaMalloc(&d_out, 1024 * 8)); CK(cudaMalloc(&d_sink, G * kT * 8)); CK(cudaMalloc(&d_buf, bufbytes)); CK(cudaMalloc(&d_ctr, 256));
    static long long h[1024]; static u64 hs[148 * kT];
    u64 expect = 0;
    for (int r = 1; r <= rounds; r++) { for (int k = 0; k < Nk; k++) expect += (u64)r * 4096 + k + 1; for (int c = 0; c < G; c++) for (int m = 0; m < 3; m++) expect += (u64)r * 1024 + c * 4 + m + 7; }
    const size_t smem = (size_t)(4 * G + Nk + 32) * 8;
    void *fns[5] = {(void *)loop<0>, (void *)loop<1>, (void *)loop<2>, (void *)loop<3>, (void *)loop<4>};
    for (auto f : fns) CK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const char *names[5] = {"counter", "direct", "totals", "stage1", "stage8"};
    struct Case { int proto, R, use_red, delay, backoff, adaptive; };
    const Case cases[] = {
        {0, 1, 0, 0, 0, 0}, {0, 1, 1, 0, 0, 0},
        {1, 1, 0, 0, 0, 0}, {1, 1, 1, 0, 0, 0}, {1, 1, 0, 300, 0, 0}, {1, 1, 1, 300, 0, 0}, {1, 1, 0, 600, 0, 0}, {1, 1, 1, 600, 0, 0},
        {1, 1, 0, 900, 0, 0}, {1, 1, 1, 900, 0, 0}, {1, 1, 1, 1200, 0, 0}, {1, 1, 1, 600, 200, 0}, {1, 1, 1, 600, 500, 0},
        {1, 1, 0, 300, 0, 64}, {1, 1, 1, 300, 0, 64}, {1, 1, 1, 300, 0, 16}, {1, 4, 1, 300, 0, 64},
        {4, 1, 0, 0, 0, 0}, {4, 1, 1, 0, 0, 0},
    };
    for (int W : {2500}) for (int S : {0, 400, 1200}) for (int J : {0, 150, 500}) {
        for (const Case &cs : cases) {
            Prm P{d_buf, Nk, W, J, S, cs.R, rounds, cs.use_red, cs.delay, cs.backoff, cs.adaptive, d_out, d_sink, d_ctr};
            CK(cudaMemset(d_buf, 0xFF, bufbytes)); CK(cudaMemset(d_ctr, 0, 256));
            void *args[] = {&P};
            CK(cudaLaunchCooperativeKernel(fns[cs.proto], dim3(G), dim3(kT), args, smem, 0));
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", names[cs.proto], cudaGetErrorString(e)); return 1; }
            CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hs, d_sink, sizeof(hs), cudaMemcpyDeviceToHost));
            long long mx = 0, nr = 0, ad = 0; for (int i = 0; i < G; i++) { mx = h[i] > mx ? h[i] : mx; nr += h[256 + i]; ad += h[512 + i]; }
            bool ok = true;
            for (int cc = 0; cc < G; cc++) { u64 s = 0; for (int t = 0; t < kT; t++) s += hs[cc * kT + t]; if (s != expect) ok = false; }
            printf("W=%4d S=%4d J=%3d %-8s R=%2d %s delay=%4d backoff=%3d adaptive=%2d: %7.0f cycles/round, exchange %6.0f  retry-rounds %4.0f%% final-delay %4lld %s\n", W, S, J, names[cs.proto], cs.R,
                   cs.use_red ? "red" : "st ", cs.delay, cs.backoff, cs.adaptive, mx / (double)rounds, mx / (double)rounds - W, 100.0 * nr / G / rounds, ad / G, ok ? "ok" : "CHECKSUM MISMATCH");
        }
    }
    return 0;
}
