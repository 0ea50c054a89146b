// eval.cu -- stand-alone `evaluate_kernel!(out, kernel, F, t)` (src/Kernels/ModeCouplingKernels.jl:211-228, 450-467)
// streaming the raw column-major Nk x Nk vertex tables once from HBM.
//
// The reference fills A_m = V_m .* (F1 F2') and then walks diagonals/antidiagonals with stride Nk+-1
// (bengtzelius3!, :154-188).  Here a thread owns one diagonal (ip-iq = d) or one antidiagonal (iq+ip = s) and
// walks it along ip; consecutive threads own consecutive lines, so for a fixed ip a warp touches 32 consecutive
// iq -- unit stride in the column-major tables -- and the running sums never leave registers.  Lines are cut into
// segments of kSeg columns to fill the machine; a second kernel adds the segment partials of every line in a fixed order
// (one thread per line), a third tiny one scans the per-row increments and combines k*T1 + T2/k^3 + T3/k.  A_m is never materialised.
#include "eval.cuh"

namespace mct {

namespace {

constexpr int kSeg = 32;        // columns per segment
constexpr int kLineWarps = 4;   // warps (groups of 32 lines) per CTA

// lines: [0, 2Nk-1) diagonals d = L-(Nk-1);  [2Nk-1, 3Nk-2) antidiagonals s = L-(2Nk-1) in [0, Nk-2]
__global__ void __launch_bounds__(32 * kLineWarps)
sc3d_line_partials(int Nk, int nseg, const double *__restrict__ V1, const double *__restrict__ V2, const double *__restrict__ V3,
                   const double *__restrict__ F1all, const double *__restrict__ F2all, long long f1_stride,
                   double *__restrict__ part) {
    const int nlines = 3 * Nk - 2;
    const int L = (blockIdx.x * kLineWarps + (threadIdx.x >> 5)) * 32 + (threadIdx.x & 31);
    const int seg = blockIdx.y, e = blockIdx.z;
    const double *F1 = F1all + (size_t)e * f1_stride, *F2 = F2all + (size_t)e * Nk;
    if (L >= nlines) return;
    const bool diag = L < 2 * Nk - 1;
    const int d = L - (Nk - 1), s = L - (2 * Nk - 1);
    const int b0 = seg * kSeg, b1 = min(Nk, b0 + kSeg);
    double a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll 8
    for (int b = b0; b < b1; b++) {
        const int a = diag ? b - d : s - b;
        if (a >= 0 && a < Nk) {
            const size_t ix = (size_t)a + (size_t)b * Nk;
            const double f = F1[a] * F2[b];
            a1 = fma(__ldg(V1 + ix), f, a1); a2 = fma(__ldg(V2 + ix), f, a2); a3 = fma(__ldg(V3 + ix), f, a3);
        }
    }
    double *p = part + (size_t)e * 3 * nseg * nlines;
    p[((size_t)0 * nseg + seg) * nlines + L] = a1;
    p[((size_t)1 * nseg + seg) * nlines + L] = a2;
    p[((size_t)2 * nseg + seg) * nlines + L] = a3;
}

// segment partials -> line totals, one thread per (evaluation, table, line), fixed order, the loads of a thread in flight
// together.  (Round 1 added the segments inside the single finalize CTA: 2000 dependent loads per thread, 85-640 us.)
__global__ void __launch_bounds__(128) sc3d_line_totals(int nseg, int nlines, long long n_lines_all, const double *__restrict__ part,
                                                        double *__restrict__ tot) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;      // (e*3 + m) * nlines + L
    if (i >= n_lines_all) return;
    const long long em = i / nlines; const int L = (int)(i - em * nlines);
    const double *p = part + (size_t)em * nseg * nlines + L;
    double t = 0.0;
    int sgm = 0;
    for (; sgm + 8 <= nseg; sgm += 8) {
        double v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = __ldg(p + (size_t)(sgm + j) * nlines);
#pragma unroll
        for (int j = 0; j < 8; j++) t += v[j];
    }
    for (; sgm < nseg; sgm++) t += __ldg(p + (size_t)sgm * nlines);
    tot[i] = t;
}

template <int KPT>
__global__ void __launch_bounds__(kThreads) sc3d_finalize(int Nk, int nseg, const double *__restrict__ kvec,
                                                          const double *__restrict__ part, double *__restrict__ out) {
    __shared__ double scan[3 * kWarps];
    const int nlines = 3 * Nk - 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, e = blockIdx.x;
    const double *p = part + (size_t)e * 3 * nlines;           // line totals [3][nlines] of this evaluation
    double l[3][KPT], r[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int j = 0; j < KPT; j++) {
        const int row = tid * KPT + j;   // row = ik-1
#pragma unroll
        for (int m = 0; m < 3; m++) {
            double inc = 0.0;
            if (row < Nk) {
                const double *pm = p + (size_t)m * nlines;
                inc = pm[(Nk - 1) + row];                                             // A[iq, iq+ik-1]
                if (row > 0) {
                    inc += pm[(Nk - 1) - row];                                        // A[iq+ik-1, iq]
                    inc -= pm[(2 * Nk - 1) + (row - 1)];                              // A[iq, ik-iq]
                }
            }
            r[m] += inc;
            l[m][j] = r[m];
        }
    }
    double v[3] = {r[0], r[1], r[2]}, ex[3];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
#pragma unroll
        for (int m = 0; m < 3; m++) { double u = __shfl_up_sync(0xffffffffu, v[m], o); if (lane >= o) v[m] += u; }
#pragma unroll
    for (int m = 0; m < 3; m++) {
        if (lane == 31) scan[m * kWarps + warp] = v[m];
        ex[m] = __shfl_up_sync(0xffffffffu, v[m], 1);
        if (lane == 0) ex[m] = 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < 3; m++) { double b = 0.0; for (int w = 0; w < warp; w++) b += scan[m * kWarps + w]; ex[m] += b; }
#pragma unroll
    for (int j = 0; j < KPT; j++) {
        const int row = tid * KPT + j;
        if (row < Nk) {
            const double k = kvec[row];
            const double T1 = ex[0] + l[0][j], T2 = ex[1] + l[1][j], T3 = ex[2] + l[2][j];
            out[(size_t)e * Nk + row] = k * T1 + T2 / (k * k * k) + T3 / k;           // :224-227
        }
    }
}

}  // namespace

size_t sc3d_eval_scratch_doubles(int Nk, int n) {
    const int nseg = (Nk + kSeg - 1) / kSeg;
    return (size_t)n * 3 * (nseg + 1) * (3 * (size_t)Nk - 2);      // segment partials + line totals
}

int sc3d_eval_launch(int Nk, int n, const double *dk, const double *dV1, const double *dV2, const double *dV3, const double *dF1,
                     long long f1_stride, const double *dF2, double *d_scratch, double *d_out, cudaStream_t s) {
    const int nseg = (Nk + kSeg - 1) / kSeg, nlines = 3 * Nk - 2;
    const int KPT = (Nk + kThreads - 1) / kThreads;
    MCT_REQUIRE(KPT <= kMaxKPT, MCT_UNSUPPORTED, "Nk > 2048 is not supported");
    dim3 grid((nlines + 32 * kLineWarps - 1) / (32 * kLineWarps), nseg, n);
    sc3d_line_partials<<<grid, 32 * kLineWarps, 0, s>>>(Nk, nseg, dV1, dV2, dV3, dF1, dF2, f1_stride, d_scratch);
    double *d_tot = d_scratch + (size_t)n * 3 * nseg * nlines;
    const long long n_lines_all = (long long)n * 3 * nlines;
    sc3d_line_totals<<<(unsigned)((n_lines_all + 127) / 128), 128, 0, s>>>(nseg, nlines, n_lines_all, d_scratch, d_tot);
    if (KPT <= 1) sc3d_finalize<1><<<n, kThreads, 0, s>>>(Nk, nseg, dk, d_tot, d_out);
    else if (KPT <= 2) sc3d_finalize<2><<<n, kThreads, 0, s>>>(Nk, nseg, dk, d_tot, d_out);
    else if (KPT <= 4) sc3d_finalize<4><<<n, kThreads, 0, s>>>(Nk, nseg, dk, d_tot, d_out);
    else sc3d_finalize<8><<<n, kThreads, 0, s>>>(Nk, nseg, dk, d_tot, d_out);
    count_launch(3);
    MCT_CUDA(cudaGetLastError());
    return MCT_OK;
}

}  // namespace mct
