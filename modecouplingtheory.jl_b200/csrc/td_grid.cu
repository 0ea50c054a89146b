// td_grid.cu -- whole-grid cooperative time-doubling / steady-state solver for block-valued and dense kernels.
//
// MultiComponentModeCouplingKernel3D (src/Kernels/MultiComponentKernels.jl:84-219) and dDimModeCouplingKernel
// (src/Kernels/ModeCouplingKernels.jl:19-71, 237-256) behind the same solver loop as td_sc.cu
// (src/TimeDoublingSolver.jl:512-532), with every phase spread over the whole grid and a grid-wide barrier between
// phases.  The loop structure, the Euler bootstrap (src/EulerSolver.jl:26-64) and all per-k block algebra follow the
// reference's operation order (K*F products on the left, C1 \ rhs per wavenumber -- src/HelperFunctions.jl:11-36,
// src/TimeDoublingSolver.jl:326-337); the file is compiled with -fmad=false.
//
// Multi-component evaluation: the reference's fill_A! costs O(Ns^4 Nk^2) (:134-198).  Its four (mu', nu') sums
// factorise into Hadamard products of per-k matrices  G1 = C'FC, G2 = C'F, G3 = FC:
//     sum(term1) = F_q .* G1_p,   sum(term2) = G3_q .* G2_p,   sum(term3) = G2_q .* G3_p,   sum(term4) = G1_q .* F_p
// so an entry A_m[a,b,iq,ip] costs O(1) per (a,b) and A is never stored: a warp owns one diagonal / antidiagonal of
// the (iq, ip) plane and accumulates its 3 Ns^2 line sums; bengtzelius3! (:2-43) then is a prefix over the lines.
// This first version is throughput-naive (operands re-read from L2 per entry, sequential prefix); it exists for
// parity and coverage, the tuned path is the single-component one.
#include <cooperative_groups.h>

#include "td_grid.cuh"

namespace cg = cooperative_groups;

namespace mct {

namespace {

constexpr int kGT = 256;        // threads per block
constexpr int kNB = kNsMax * kNsMax;

// C = A*B, column-major ns x ns, plain left-to-right sums (StaticArrays `*`)
__device__ __forceinline__ void blk_mul(int ns, const double *A, const double *B, double *C) {
    double tmp[kNB];
    for (int j = 0; j < ns; j++)
        for (int i = 0; i < ns; i++) {
            double s = A[i] * B[j * ns];
            for (int l = 1; l < ns; l++) s += A[i + l * ns] * B[l + j * ns];
            tmp[i + j * ns] = s;
        }
    for (int e = 0; e < ns * ns; e++) C[e] = tmp[e];
}
// X = A \ B (ns right-hand sides), Gaussian elimination with partial pivoting
__device__ __forceinline__ void blk_solve(int ns, const double *A, const double *B, double *X) {
    if (ns == 1) { X[0] = B[0] / A[0]; return; }
    double a[kNB], b[kNB];
    for (int e = 0; e < ns * ns; e++) { a[e] = A[e]; b[e] = B[e]; }
    for (int c = 0; c < ns; c++) {
        int piv = c; double best = fabs(a[c + c * ns]);
        for (int r = c + 1; r < ns; r++) if (fabs(a[r + c * ns]) > best) { best = fabs(a[r + c * ns]); piv = r; }
        if (piv != c)
            for (int j = 0; j < ns; j++) {
                double t = a[c + j * ns]; a[c + j * ns] = a[piv + j * ns]; a[piv + j * ns] = t;
                t = b[c + j * ns]; b[c + j * ns] = b[piv + j * ns]; b[piv + j * ns] = t;
            }
        for (int r = c + 1; r < ns; r++) {
            const double f = a[r + c * ns] / a[c + c * ns];
            if (f == 0.0) continue;
            for (int j = c; j < ns; j++) a[r + j * ns] -= f * a[c + j * ns];
            for (int j = 0; j < ns; j++) b[r + j * ns] -= f * b[c + j * ns];
        }
    }
    for (int j = 0; j < ns; j++)
        for (int r = ns - 1; r >= 0; r--) {
            double s = b[r + j * ns];
            for (int c = r + 1; c < ns; c++) s -= a[r + c * ns] * X[c + j * ns];
            X[r + j * ns] = s / a[r + r * ns];
        }
}

constexpr unsigned long long kSentinel = 0xFFFFFFFFFFFFFFFFull;   // a NaN payload no arithmetic produces

__device__ __forceinline__ void st_relaxed_sys(double *p, double v) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_bits(const double *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct Ctx {
    const GridParams &P;
    cg::grid_group grid;
    int nb, dof, gtid, gthreads, lane, gwarp, gwarps;
    unsigned xepoch = 0;
    bool failed = false;
    __device__ Ctx(const GridParams &p) : P(p), grid(cg::this_grid()) {
        nb = p.ns * p.ns; dof = p.Nk * nb;
        gtid = blockIdx.x * blockDim.x + threadIdx.x; gthreads = gridDim.x * blockDim.x;
        lane = threadIdx.x & 31; gwarp = gtid >> 5; gwarps = gthreads >> 5;
    }

    // Line sums of this rank's lines (L = rank, rank + world, ...; line lengths vary smoothly with L, so the interleave
    // balances).  The work is split by species pair e = (a,b) FIRST: a block stages the four operand vectors of ITS e
    // (F, G1, G2, G3 along k: 32 Nk bytes) and the wavenumbers in shared memory and then walks lines reading only
    // shared memory -- every operand is used by ~3 Nk lines, and re-reading them from L2 per entry made the walk
    // L2-bandwidth bound (6 GB per evaluation at Ns=4, Nk=2000).
    __device__ void line_sums(const double *Fs, const double *G1, const double *G2, const double *G3, int nlines, size_t xwords) {
        extern __shared__ double gsm[];
        const int Nk = P.Nk, world = P.world, rank = P.rank;
        const int e = blockIdx.x % nb;                       // species pair of this block
        const int chunk = blockIdx.x / nb, nchunk = (gridDim.x - e + nb - 1) / nb;    // blocks sharing this e
        double *sF = gsm, *s1 = gsm + Nk, *s2 = gsm + 2 * Nk, *s3 = gsm + 3 * Nk, *sk = gsm + 4 * Nk;
        for (int k = threadIdx.x; k < Nk; k += blockDim.x) {
            sF[k] = Fs[(size_t)e * Nk + k]; s1[k] = G1[(size_t)e * Nk + k]; s2[k] = G2[(size_t)e * Nk + k]; s3[k] = G3[(size_t)e * Nk + k];
            sk[k] = P.kvec[k];
        }
        __syncthreads();
        const double pab = P.pref[e];
        const int warps_per_block = blockDim.x >> 5, wi = chunk * warps_per_block + (threadIdx.x >> 5), wstride = nchunk * warps_per_block;
        for (int Li = wi; Li * world + rank < nlines; Li += wstride) {
            const int L = Li * world + rank;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
            const bool diag = L < 2 * Nk - 1;
            const int d = L - (Nk - 1), s = L - (2 * Nk - 1);
            const int q0 = diag ? max(0, -d) : 0, q1 = diag ? min(Nk, Nk - d) : s + 1;
            for (int iq = q0 + lane; iq < q1; iq += 32) {
                const int ip = diag ? iq + d : s - iq;
                const double q = sk[iq], p = sk[ip];
                const double pq2 = p * p - q * q;
                const double P1 = p * q, P2 = p * q * (pq2 * pq2), P3 = 2.0 * p * q * pq2;
                const double t1 = sF[iq] * s1[ip], t2 = s2[ip] * s3[iq], t3 = s2[iq] * s3[ip], t4 = s1[iq] * sF[ip];
                a0 += (t1 + t2 + t3 + t4) * P1 * pab;      // :186-192
                a1 += (t1 - t2 - t3 + t4) * P2 * pab;
                a2 += (t1 - t4) * P3 * pab;
            }
            a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
            if (lane == 0) {
                const size_t w = (size_t)L * 3 * nb + e;
                if (world == 1) { P.Lsum[w] = a0; P.Lsum[w + nb] = a1; P.Lsum[w + 2 * nb] = a2; }
                else for (int r = 0; r < world; r++) {
                    double *dst = P.xpeer[r] + (size_t)(xepoch & 3u) * xwords + w;
                    st_relaxed_sys(dst, a0); st_relaxed_sys(dst + nb, a1); st_relaxed_sys(dst + 2 * nb, a2);
                }
            }
        }
        __syncthreads();
    }

    // ---- K = kernel(F)  (ends with a grid barrier; Kout valid for every thread afterwards)
    __device__ void eval(const double *F, double *Kout) {
        const int Nk = P.Nk, ns = P.ns;
        if (P.type == GK_DDIM) {
            // P[iq] = sum_{ik,ip} prefactor*J*V * F[ik]*F[ip]   (ModeCouplingKernels.jl:237-256)
            __shared__ double red[kGT / 32];
            for (int iq = blockIdx.x; iq < Nk; iq += gridDim.x) {
                const double *W = P.W + (size_t)iq * Nk * Nk;
                double s = 0.0;
                for (int i = threadIdx.x; i < Nk * Nk; i += blockDim.x) {
                    const int ik = i / Nk, ip = i - ik * Nk;
                    s += W[i] * F[ik] * F[ip];
                }
                s = warp_sum(s);
                if (lane == 0) red[threadIdx.x >> 5] = s;
                __syncthreads();
                if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < kGT / 32; w++) t += red[w]; Kout[iq] = t; }
                __syncthreads();
            }
            grid.sync();
            return;
        }
        // (1) G1 = C'FC, G2 = C'F, G3 = FC per wavenumber, stored species-major ([e][k]) together with a copy of F so that
        //     the line walk below reads unit-stride along k
        double *Fs = P.G, *G1 = P.G + dof, *G2 = P.G + 2 * (size_t)dof, *G3 = P.G + 3 * (size_t)dof;
        for (int k = gtid; k < Nk; k += gthreads) {
            const double *Ck = P.C + (size_t)k * nb, *Fk = F + (size_t)k * nb;
            double ct[kNB], g1[kNB], g2[kNB], g3[kNB];
            for (int a = 0; a < ns; a++) for (int b = 0; b < ns; b++) ct[a + b * ns] = Ck[b + a * ns];
            blk_mul(ns, Fk, Ck, g3);
            blk_mul(ns, ct, Fk, g2);
            blk_mul(ns, ct, g3, g1);
            for (int e = 0; e < nb; e++) {
                Fs[(size_t)e * Nk + k] = Fk[e]; G1[(size_t)e * Nk + k] = g1[e]; G2[(size_t)e * Nk + k] = g2[e]; G3[(size_t)e * Nk + k] = g3[e];
            }
        }
        grid.sync();
        // (2) line sums: lines [0, 2Nk-1) are diagonals ip - iq = L-(Nk-1); [2Nk-1, 3Nk-2) antidiagonals iq + ip = L-(2Nk-1)
        const int nlines = 3 * Nk - 2;
        const int world = P.world, rank = P.rank;
        const size_t xwords = (size_t)3 * Nk * 3 * nb;
        xepoch++;
        double *lcur = (world > 1) ? P.xpeer[rank] + (size_t)(xepoch & 3u) * xwords : P.Lsum;
        // this rank's lines: L = rank, rank + world, ...  (line lengths vary smoothly with L, so the interleave balances)
        line_sums(Fs, G1, G2, G3, nlines, xwords);
        if (world > 1) {
            // re-arm this rank's own copy of the buffer used two exchanges from now (every peer wrote it two exchanges ago
            // and will write it again only after it has seen this rank's data of the next exchange)
            unsigned long long *clr = (unsigned long long *)(P.xpeer[rank] + (size_t)((xepoch + 2u) & 3u) * xwords);
            for (size_t i = gtid; i < xwords; i += gthreads) clr[i] = kSentinel;
            __threadfence_system();
        }
        grid.sync();
        // (3) bengtzelius3! (:2-43) as a prefix over the line sums: increments per ik (whole grid), then one warp per
        //     (m, e) sequence scans its Nk increments 32 at a time; K = k T1 + T2/k^3 + T3/k   (:215-218)
        {
            const long long t0 = clock64();
            // line sums of other ranks arrive through NVLink: every word validates itself (sentinel until stored)
            auto rd = [&](size_t w) -> double {
                if (world == 1) return lcur[w];
                unsigned long long v;
                unsigned spins = 0;
                while ((v = ld_relaxed_sys_bits(lcur + w)) == kSentinel) {
                    if ((++spins & 1023u) == 0 && clock64() - t0 > P.spin_limit_cycles) { failed = true; return 0.0; }
                }
                return __longlong_as_double((long long)v);
            };
            const int nseq = 3 * nb;
            double *inc = P.Inc;                                   // [nseq][Nk]
            for (size_t i = gtid; i < (size_t)nseq * Nk; i += gthreads) {
                const int ik = (int)(i / nseq) + 1, c = (int)(i % nseq);      // consecutive threads: consecutive words of a line
                double v = rd((size_t)((Nk - 1) + (ik - 1)) * nseq + c);
                if (ik > 1) { v += rd((size_t)((Nk - 1) - (ik - 1)) * nseq + c); v -= rd((size_t)((2 * Nk - 1) + (ik - 2)) * nseq + c); }
                inc[(size_t)c * Nk + (ik - 1)] = v;
            }
            if (failed) *P.status = MCT_CUDA_ERROR;
        }
        grid.sync();
        for (int c = gwarp; c < 3 * nb; c += gwarps) {
            const double *inc = P.Inc + (size_t)c * Nk;
            double *Tm = P.Tm + (size_t)c * Nk;
            double carry = 0.0;
            for (int b0 = 0; b0 < Nk; b0 += 32) {
                const int ik = b0 + lane;
                double v = (ik < Nk) ? inc[ik] : 0.0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
                v += carry;
                if (ik < Nk) Tm[ik] = v;
                carry = __shfl_sync(0xffffffffu, v, 31);
            }
        }
        grid.sync();
        for (size_t i = gtid; i < (size_t)dof; i += gthreads) {
            const int ik = (int)(i / nb), e = (int)(i % nb);
            const double k = P.kvec[ik];
            Kout[i] = k * P.Tm[(size_t)e * Nk + ik] + P.Tm[(size_t)(nb + e) * Nk + ik] / (k * k * k) + P.Tm[(size_t)(2 * nb + e) * Nk + ik] / k;
        }
        grid.sync();
    }

    // max |a - b| over the grid through a rotating atomic word; every thread returns the same value
    __device__ double grid_max(double local, unsigned &ecount) {
        __shared__ double red[kGT / 32];
        double m = warp_max(local);
        if (lane == 0) red[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < kGT / 32; w++) t = fmax(t, red[w]);
            atomicMax(P.err + (ecount & 1u), (unsigned long long)__double_as_longlong(t));
            if (blockIdx.x == 0) P.err[(ecount + 1u) & 1u] = 0ull;     // nobody reads or writes that word in this phase
        }
        grid.sync();
        const double r = __longlong_as_double((long long)*(volatile unsigned long long *)(P.err + (ecount & 1u)));
        ecount++;
        return r;
    }

    // F = C1 \ (-K*C2 + C3) for every wavenumber; returns the local max |F - Fold| and sets Fold = F  (:326-337, :410-414)
    __device__ double update_F(const double *K, double *Fdst, bool track) {
        const int ns = P.ns;
        double lerr = 0.0;
        for (int k = gtid; k < P.Nk; k += gthreads) {
            const size_t o = (size_t)k * nb;
            double t[kNB], x[kNB];
            blk_mul(ns, K + o, P.C2 + o, t);
            for (int e = 0; e < nb; e++) t[e] = -t[e] + P.C3[o + e];
            blk_solve(ns, P.C1 + o, t, x);
            for (int e = 0; e < nb; e++) {
                if (track) { const double d = fabs(x[e] - P.Fold[o + e]); if (d > lerr) lerr = d; P.Fold[o + e] = x[e]; }
                Fdst[o + e] = x[e];
            }
        }
        return lerr;
    }

    __device__ void run() {
        const int Nk = P.Nk, ns = P.ns, N = P.N, n4 = 4 * N, n2 = 2 * N;
        unsigned ecount = 0;
#define FT(s) (P.Ft + (size_t)dof * ((s) - 1))
#define KT(s) (P.Kt + (size_t)dof * ((s) - 1))
#define FINT(s) (P.FI + (size_t)dof * ((s) - 1))
#define KINT(s) (P.KI + (size_t)dof * ((s) - 1))
        if (P.mode == 2) {                       // evaluate_kernel!(out, kernel, F, t)
            eval(P.Fin, P.F_out);
            return;
        }
        eval(P.F0, P.K0);                        // K0 = evaluate_kernel(kernel, F0, 0)  src/MemoryEquation.jl:45
        if (P.mode == 1) {                       // solve_steady_state  src/SteadyStateMemoryEquation.jl:59-100
            for (int e = gtid; e < dof; e += gthreads) { P.Fold[e] = P.F0[e]; P.Kcur[e] = P.K0[e]; }
            grid.sync();
            long long iter = 0; int status = MCT_OK;
            while (true) {
                iter++;
                if (iter > P.max_iter) { status = MCT_NOT_CONVERGED; break; }
                double lerr = 0.0;
                for (int k = gtid; k < Nk; k += gthreads) {
                    const size_t o = (size_t)k * nb;
                    double t[kNB], M[kNB], x[kNB];
                    blk_mul(ns, P.Kcur + o, P.F0 + o, t);                                  // :81
                    for (int e = 0; e < nb; e++) M[e] = P.Kcur[o + e] + P.gamma[o + e];    // :83
                    blk_solve(ns, M, t, x);
                    for (int e = 0; e < nb; e++) { const double d = fabs(x[e] - P.Fold[o + e]); if (d > lerr) lerr = d; P.Fold[o + e] = x[e]; P.Fcur[o + e] = x[e]; }
                }
                const double err = grid_max(lerr, ecount);
                eval(P.Fcur, P.Kcur);                                                      // :89
                if (!(err > P.tol)) break;
            }
            for (int e = gtid; e < dof; e += gthreads) { P.F_out[e] = P.Fcur[e]; P.K_out[e] = P.Kcur[e]; }
            if (gtid == 0) { if (*P.status != MCT_CUDA_ERROR) *P.status = status; *P.kernel_evals = iter; }
            return;
        }
        // ---- time doubling.  F_old = K0*F0 (:64); K_temp slots keep 2 K0 until they are evaluated (:83)
        for (int k = gtid; k < Nk; k += gthreads) blk_mul(ns, P.K0 + (size_t)k * nb, P.F0 + (size_t)k * nb, P.Fold + (size_t)k * nb);
        for (int e = gtid; e < dof; e += gthreads) {
            for (int s = n2 + 1; s <= n4; s++) KT(s)[e] = P.K0[e] + P.K0[e];
            P.dFh[e] = P.dF0[e];
            P.F_out[e] = P.F0[e]; P.K_out[e] = P.K0[e];
        }
        grid.sync();
        // bootstrap: forward Euler with the trapezoid memory integral  (src/EulerSolver.jl:26-64, 92-112)
        const double de = P.dt0 / (4.0 * N);
        for (int it = 1; it <= n2; it++) {
            for (int k = gtid; k < Nk; k += gthreads) {
                const size_t o = (size_t)k * nb;
                double integ[kNB], f[kNB];
                // integrand[i] = K_array[i] * dF_reverse[i] = K(t_{i-1}) * dF(t_{it-i}), i = 1..it
                if (it == 1) {
                    blk_mul(ns, P.K0 + o, P.dFh + o, f);
                    for (int e = 0; e < nb; e++) integ[e] = f[e] * de;                     // :34-36
                } else {
                    double r[kNB];
                    blk_mul(ns, P.K0 + o, P.dFh + (size_t)dof * (it - 1) + o, f);
                    for (int e = 0; e < nb; e++) r[e] = f[e] / 2.0;
                    blk_mul(ns, KT(it - 1) + o, P.dFh + o, f);
                    for (int e = 0; e < nb; e++) r[e] += f[e] / 2.0;
                    for (int i = 2; i <= it - 1; i++) {
                        blk_mul(ns, KT(i - 1) + o, P.dFh + (size_t)dof * (it - i) + o, f);
                        for (int e = 0; e < nb; e++) r[e] += f[e];
                    }
                    for (int e = 0; e < nb; e++) integ[e] = r[e] * de;
                }
                const double *Fo = (it == 1) ? P.F0 + o : FT(it - 1) + o;
                const double *dFo = P.dFh + (size_t)dof * (it - 1) + o;
                double gF[kNB], rhs[kNB], neg[kNB], sol[kNB];
                blk_mul(ns, P.gamma + o, Fo, gF);
                if (P.second_order) {                                                      // Euler_step :55-58
                    double bdF[kNB];
                    blk_mul(ns, P.beta + o, dFo, bdF);
                    for (int e = 0; e < nb; e++) { rhs[e] = bdF[e] + gF[e] + P.delta[o + e] + integ[e]; neg[e] = -P.alpha[o + e]; }
                    blk_solve(ns, neg, rhs, sol);
                    for (int e = 0; e < nb; e++) {
                        const double dFn = dFo[e] + de * sol[e];
                        P.dFh[(size_t)dof * it + o + e] = dFn;
                        FT(it)[o + e] = Fo[e] + de * dFn;
                    }
                } else {                                                                   // :60-61
                    for (int e = 0; e < nb; e++) { rhs[e] = gF[e] + P.delta[o + e] + integ[e]; neg[e] = -P.beta[o + e]; }
                    blk_solve(ns, neg, rhs, sol);
                    for (int e = 0; e < nb; e++) {
                        P.dFh[(size_t)dof * it + o + e] = sol[e];
                        FT(it)[o + e] = Fo[e] + de * sol[e];
                    }
                }
            }
            grid.sync();
            eval(FT(it), KT(it));                                                          // == initialize_K_temp! :155-166
            for (int e = gtid; e < dof; e += gthreads) { P.F_out[(size_t)it * dof + e] = FT(it)[e]; P.K_out[(size_t)it * dof + e] = KT(it)[e]; }
        }
        // initialize_integrals! :174-199
        for (int e = gtid; e < dof; e += gthreads) {
            FINT(1)[e] = (FT(1)[e] + P.F0[e]) / 2.0;
            KINT(1)[e] = (3.0 * KT(1)[e] - KT(2)[e]) / 2.0;
            for (int it = 2; it <= n2; it++) {
                FINT(it)[e] = (FT(it)[e] + FT(it - 1)[e]) / 2.0;
                KINT(it)[e] = (KT(it)[e] + KT(it - 1)[e]) / 2.0;
            }
        }
        grid.sync();
        long long evals = 0;
        int status = MCT_OK;
        double Dt = P.dt0;
        for (int blk = 0; blk < P.nblocks && status == MCT_OK; blk++, Dt *= 2.0) {
            const double dl = Dt / (4.0 * N);
            for (int it = n2 + 1; it <= n4 && status == MCT_OK; it++) {
                const int i2 = it / 2;
                const long long row = 1 + (long long)n2 * (blk + 1) + (it - n2 - 1);
                // update_Fuchs_parameters! :300-319 (compute_C3_mutable! :243-288), then update_F! with the stale K_temp[it]
                for (int k = gtid; k < Nk; k += gthreads) {
                    const size_t o = (size_t)k * nb;
                    double tv[kNB], c3[kNB], t[kNB], u[kNB];
                    for (int e = 0; e < nb; e++)
                        P.C1[o + e] = ((2.0 / (dl * dl) * P.alpha[o + e] + 3.0 / (2.0 * dl) * P.beta[o + e]) + KINT(1)[o + e]) + P.gamma[o + e];   // :314
                    for (int e = 0; e < nb; e++) P.C2[o + e] = FINT(1)[o + e] - P.F0[o + e];                                                        // :315
                    for (int e = 0; e < nb; e++) tv[e] = (5.0 * FT(it - 1)[o + e] - 4.0 * FT(it - 2)[o + e] + FT(it - 3)[o + e]) / (dl * dl);
                    blk_mul(ns, P.alpha + o, tv, c3);                                                                                              // :259-260
                    for (int e = 0; e < nb; e++) tv[e] = 2.0 / dl * FT(it - 1)[o + e] - FT(it - 2)[o + e] / (2.0 * dl);
                    blk_mul(ns, P.beta + o, tv, t);
                    for (int e = 0; e < nb; e++) c3[e] += t[e];                                                                                    // :263-264
                    blk_mul(ns, KT(it - i2) + o, FT(i2) + o, t); for (int e = 0; e < nb; e++) c3[e] -= t[e];                                        // :267
                    blk_mul(ns, KT(it - 1) + o, FINT(1) + o, t); for (int e = 0; e < nb; e++) c3[e] += t[e];                                        // :268
                    blk_mul(ns, KINT(1) + o, FT(it - 1) + o, t); for (int e = 0; e < nb; e++) c3[e] += t[e];                                        // :269
                    for (int j = 2; j <= i2; j++) {                                                                                                // :271-280
                        for (int e = 0; e < nb; e++) u[e] = KT(it - j)[o + e] - KT(it - j + 1)[o + e];
                        blk_mul(ns, u, FINT(j) + o, t);
                        for (int e = 0; e < nb; e++) c3[e] += t[e];
                    }
                    for (int j = 2; j <= it - i2; j++) {                                                                                           // :281-285
                        for (int e = 0; e < nb; e++) u[e] = FT(it - j)[o + e] - FT(it - j + 1)[o + e];
                        blk_mul(ns, KINT(j) + o, u, t);
                        for (int e = 0; e < nb; e++) c3[e] += t[e];
                    }
                    for (int e = 0; e < nb; e++) P.C3[o + e] = c3[e] - P.delta[o + e];                                                             // :286
                }
                // the same thread owns wavenumber k in every per-k phase: no barrier needed between them
                update_F(KT(it), FT(it), false);
                grid.sync();
                long long iter = 0;
                while (true) {
                    iter++;
                    if (iter > P.max_iter) { status = MCT_NOT_CONVERGED; break; }          // :406-408
                    eval(FT(it), KT(it));                                                  // update_K! :360-369
                    const double lerr = update_F(KT(it), FT(it), true);                    // update_F! + find_error + F_old .= F
                    const double err = grid_max(lerr, ecount);
                    evals++;                                                               // :416
                    if (!(err > P.tol)) break;
                }
                if (status != MCT_OK) break;
                for (int e = gtid; e < dof; e += gthreads) {                               // update_integrals! :377-384, allocate_results! :429-438
                    FINT(it)[e] = (FT(it)[e] + FT(it - 1)[e]) / 2.0;
                    KINT(it)[e] = (KT(it)[e] + KT(it - 1)[e]) / 2.0;
                    P.F_out[(size_t)row * dof + e] = FT(it)[e]; P.K_out[(size_t)row * dof + e] = KT(it)[e];
                }
                grid.sync();
            }
            if (status != MCT_OK) break;
            for (int e = gtid; e < dof; e += gthreads) {                                   // new_time_mapping! :448-489
                for (int j = 1; j <= n2; j++) {
                    FINT(j)[e] = (FINT(2 * j)[e] + FINT(2 * j - 1)[e]) / 2.0;
                    KINT(j)[e] = (KINT(2 * j)[e] + KINT(2 * j - 1)[e]) / 2.0;
                    FT(j)[e] = FT(2 * j)[e];
                    KT(j)[e] = KT(2 * j)[e];
                }
                for (int j = n2 + 1; j <= n4; j++) { FINT(j)[e] = 0.0; KINT(j)[e] = 0.0; FT(j)[e] = 0.0; KT(j)[e] = 0.0; }
            }
            grid.sync();
        }
        if (gtid == 0) { if (*P.status != MCT_CUDA_ERROR) *P.status = status; *P.kernel_evals = evals; }
#undef FT
#undef KT
#undef FINT
#undef KINT
    }
};

__global__ void __launch_bounds__(kGT) td_grid_kernel(const GridParams P) {
    Ctx c(P);
    c.run();
}

__global__ void ddim_table_kernel(int Nk, const double *V, const double *J, double prefactor, double *W) {
    // out W[iq][ik][ip] (ip contiguous) = prefactor * J[iq,ik,ip] * V[iq,ik,ip]   (column-major inputs: iq fastest)
    const size_t n = (size_t)Nk * Nk * Nk;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int ip = (int)(i % Nk), ik = (int)((i / Nk) % Nk), iq = (int)(i / ((size_t)Nk * Nk));
        const size_t src = (size_t)iq + (size_t)Nk * ((size_t)ik + (size_t)Nk * ip);
        W[i] = prefactor * J[src] * V[src];
    }
}

}  // namespace

int ddim_build_table(int Nk, double prefactor, const double *dV, const double *dJ, double *dW, cudaStream_t s) {
    ddim_table_kernel<<<1024, 256, 0, s>>>(Nk, dV, dJ, prefactor, dW);
    count_launch();
    MCT_CUDA(cudaGetLastError());
    return MCT_OK;
}

int td_grid_launch(const GridParams &P, int n_sm, cudaStream_t stream) {
    MCT_REQUIRE(P.ns >= 1 && P.ns <= kNsMax, MCT_UNSUPPORTED, "Ns > 4 is not supported on the device");
    // multi-component: five vectors of Nk doubles per block in shared memory (line_sums)
    const size_t smem = (P.type == GK_MC3D) ? (size_t)5 * P.Nk * sizeof(double) : 0;
    MCT_CUDA(cudaFuncSetAttribute((void *)td_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MCT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, td_grid_kernel, kGT, smem));
    MCT_REQUIRE(per_sm >= 1, MCT_CUDA_ERROR, "cooperative kernel does not fit");
    GridParams Pc = P;
    void *args[] = {(void *)&Pc};
    dim3 grid(n_sm * std::min(per_sm, 2)), block(kGT);
    MCT_CUDA(cudaLaunchCooperativeKernel((void *)td_grid_kernel, grid, block, args, smem, stream));
    count_launch();
    return MCT_OK;
}

}  // namespace mct
