// td_sc.cu -- ONE persistent cooperative launch runs the whole Fuchs time-doubling solve
// (src/TimeDoublingSolver.jl:512-532) of one or many single-component memory equations.
//
// A "team" of G CTAs (one CTA per SM) owns one equation at a time:
//   * the vertex tables stay resident in the team's shared memory for the whole solve (team layout, layout.cuh);
//   * every fixed-point iteration (update_K!/update_F!/find_error, :404-417) is: each warp streams its share of
//     table terms -> per-row increments -> ONE all-gather of 3*Nk doubles through an L2-resident exchange
//     buffer -> every CTA redundantly scans the increments into T1..T3, forms K, F and the max-error, so all
//     CTAs take the same branch with no second exchange;
//   * everything that is local in k (memory convolution C3 :243-288, history rings, time mapping :448-489) is
//     sharded over the team by wavenumber ("owners" = the CTA that owns row k) and never crosses CTAs, except
//     C1..C3 once per time step.
// The all-gather needs no barrier and no fence: every 8-byte slot validates itself.  Slots hold a sentinel bit
// pattern until their owner stores the value; readers spin on the slot itself (one L2 round trip when the data is
// already there).  Four buffers rotate; at exchange e a CTA clears its own slots of buffer (e+2)&3, which every
// CTA finished reading before this CTA could complete exchange e-1 ("exchange protocol").
// Arithmetic outside the table stream follows the reference's operation order (file compiled with -fmad=false).
#include <algorithm>

#include "td_sc.cuh"

namespace mct {

namespace {

constexpr unsigned long long kSentinel = 0xFFFFFFFFFFFFFFFFull;   // a NaN payload no arithmetic produces

__host__ __device__ inline int even_up(int n) { return (n + 1) & ~1; }

struct SmemPlan {
    size_t off_F1, off_F2, off_K, off_X, off_part, off_scan, off_mx, off_red, off_tasks, off_tb, off_rows, off_rs, off_abort, off_tab, total;
};

__host__ __device__ inline SmemPlan plan_smem(int Nk, int G, int slots_max, int tasks_max, int rows_max, int smem_terms) {
    SmemPlan p;
    size_t o = 0;
    int NkP = even_up(Nk);
    p.off_F1 = o; o += (size_t)NkP * 8;
    p.off_F2 = o; o += (size_t)NkP * 8;
    p.off_K = o; o += (size_t)2 * even_up(rows_max) * 8;      // sC: the two convolution sums of the owned rows
    p.off_X = o; o += (G == 1 ? (size_t)NkP * 8 : 16);
    p.off_part = o; o += (size_t)3 * even_up(slots_max) * 8;
    p.off_scan = o; o += (size_t)3 * kWarps * 8;
    p.off_mx = o; o += (size_t)kWarps * 8;
    p.off_red = o; o += (size_t)3 * kWarps * 8;
    p.off_tasks = o; o += (size_t)kWarps * tasks_max * sizeof(Task);
    p.off_tb = o; o += (size_t)even_up(kWarps + 1) * 4;
    p.off_rows = o; o += (size_t)even_up(rows_max) * 4;
    p.off_rs = o; o += (size_t)even_up(rows_max + 1) * 4;
    p.off_abort = o; o += 8;
    o = (o + 15) & ~(size_t)15;
    p.off_tab = o; o += (size_t)3 * smem_terms * 8;
    p.total = o;
    return p;
}

__device__ __forceinline__ void st_relaxed(double *p, double v) {
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_bits(double *p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_bits(const double *p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

#ifdef MCT_PROFILE
#define PROF(i) do { long long now__ = clock64(); prof[i] += now__ - tlast; tlast = now__; } while (0)
#define NPROF 12
#else
#define PROF(i) do { } while (0)
#endif

template <int KPT>
struct Team {
    const SolveParams &P;
    // shared memory
    double *sF1, *sF1_home, *sF2, *sX, *sPart, *sScan, *sMax, *sRed, *sC, *sTab;
    const Task *sTasks;
    const int *sTaskBegin, *sRows, *sRowSlot;
    volatile int *sAbort;
    // identity
    int team, cta, tid, lane, warp;
    int Nk, NkP, G;
    int nparts;                  // row parts streamed by this CTA (thread r < nparts handles part r)
    int myrow;                   // wavenumber index owned by this thread (its row ends in this CTA), or -1
    // exchange
    double *xbase;               // four rotating buffers of xbuf_doubles words
    int xbuf_doubles;
    int myslot;                  // exchange word of the wavenumber this thread owns
    int rslot[KPT];              // exchange words of the wavenumbers this thread collects (k = tid + j*512)
    unsigned epoch;
    long long slab_off;
    int slab_len;
#ifdef MCT_PROFILE
    long long prof[12], tlast;
#endif

    __device__ __forceinline__ Team(const SolveParams &p) : P(p) {}

    // ------------------------------------------------------------------------------------------ exchange
    // Buffer layout: every 128-byte line has ONE writer CTA (lines shared by several writers cost 4-8x more under
    // polling, scripts/microbench.cu).  Line c < G: the three block totals of CTA c (+ a spare word); then each CTA
    // owns LF lines with the per-wavenumber values of its rows.  See the header comment for the protocol.
    __device__ __forceinline__ double *xcur() const { return xbase + (size_t)(epoch & 3u) * xbuf_doubles; }
    __device__ __forceinline__ double *xclr() const { return xbase + (size_t)((epoch + 2u) & 3u) * xbuf_doubles; }
    __device__ __forceinline__ double spin_read(const double *p) {
        unsigned long long v = ld_relaxed_bits(p);
        if (v == kSentinel) {
            const long long t0 = clock64();
            unsigned spins = 0;
            do {
                v = ld_relaxed_bits(p);
                if ((++spins & 1023u) == 0) {
                    if (*(volatile int *)P.abort_flag) { *sAbort = 1; v = 0; break; }
                    if (clock64() - t0 > P.spin_limit_cycles) { atomicExch(P.abort_flag, 1); *sAbort = 1; v = 0; break; }
                }
            } while (v == kSentinel);
        }
        return __longlong_as_double((long long)v);
    }
    // Every exchange (of either kind) starts here: advance the epoch and re-arm this CTA's own slots two exchanges
    // ahead.  Must be called by all threads, after a block barrier that follows the reads of the previous exchange.
    __device__ __forceinline__ void begin_exchange() {
        epoch++;
        if (G == 1) return;
        double *c = xclr();
        if (myrow >= 0) st_relaxed_bits(c + myslot, kSentinel);
        if (tid < 3) st_relaxed_bits(c + 16 * cta + tid, kSentinel);
    }
    __device__ __forceinline__ bool aborted() const { return *sAbort != 0; }

    // All-gather of one value per wavenumber: owners publish `mine`, every thread collects k = tid, tid+512, ...
    // into sF2.  With track_err the collected values are also compared with the previous iterate `fold`
    // (find_error, src/HelperFunctions.jl:55-64; F_old .= F_temp[it], src/TimeDoublingSolver.jl:414) and the
    // block-wide maximum is returned -- every CTA sees the same numbers, so every CTA takes the same branch.
    __device__ __forceinline__ double allgather_F(double mine, bool track_err, double (&fold)[KPT]) {
        PROF(5);
        begin_exchange();
        double err = 0.0;
        if (G == 1) {
            if (myrow >= 0) sX[myrow] = mine;
            __syncthreads();
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                const int k = tid + j * kThreads;
                if (k < Nk) {
                    const double v = sX[k];
                    if (track_err) { err = fmax(err, fabs(v - fold[j])); fold[j] = v; }
                    sF2[k] = v;
                }
            }
        } else {
            double *c = xcur();
            if (myrow >= 0) st_relaxed(c + myslot, mine);
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                const int k = tid + j * kThreads;
                if (k < Nk) {
                    const double v = spin_read(c + rslot[j]);
                    if (track_err) { err = fmax(err, fabs(v - fold[j])); fold[j] = v; }
                    sF2[k] = v;
                }
            }
        }
        PROF(6);
        const double emax = block_max(err);     // its barrier also makes sF2 complete for the next kernel evaluation
        PROF(7);
        return emax;
    }

    // ------------------------------------------------------------------------------------------ block helpers
    // inclusive block scan of three per-thread values (one __syncthreads inside); also returns the block totals
    __device__ __forceinline__ void block_inscan3(double &v0, double &v1, double &v2, double &s0, double &s1, double &s2) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            double u0 = __shfl_up_sync(0xffffffffu, v0, o), u1 = __shfl_up_sync(0xffffffffu, v1, o), u2 = __shfl_up_sync(0xffffffffu, v2, o);
            if (lane >= o) { v0 += u0; v1 += u1; v2 += u2; }
        }
        if (lane == 31) { sScan[warp] = v0; sScan[kWarps + warp] = v1; sScan[2 * kWarps + warp] = v2; }
        __syncthreads();
        double w0 = (lane < kWarps) ? sScan[lane] : 0.0, w1 = (lane < kWarps) ? sScan[kWarps + lane] : 0.0,
               w2 = (lane < kWarps) ? sScan[2 * kWarps + lane] : 0.0;
#pragma unroll
        for (int o = 1; o < kWarps; o <<= 1) {
            double u0 = __shfl_up_sync(0xffffffffu, w0, o), u1 = __shfl_up_sync(0xffffffffu, w1, o), u2 = __shfl_up_sync(0xffffffffu, w2, o);
            if (lane >= o) { w0 += u0; w1 += u1; w2 += u2; }
        }
        s0 = __shfl_sync(0xffffffffu, w0, kWarps - 1); s1 = __shfl_sync(0xffffffffu, w1, kWarps - 1); s2 = __shfl_sync(0xffffffffu, w2, kWarps - 1);
        const int src = (warp == 0) ? 0 : warp - 1;
        double b0 = __shfl_sync(0xffffffffu, w0, src), b1 = __shfl_sync(0xffffffffu, w1, src), b2 = __shfl_sync(0xffffffffu, w2, src);
        if (warp > 0) { v0 += b0; v1 += b1; v2 += b2; }
    }

    __device__ __forceinline__ double block_max(double v) {
        v = warp_max(v);
        if (lane == 0) sMax[warp] = v;
        __syncthreads();
        double m = (lane < kWarps) ? sMax[lane] : 0.0;
        return warp_max(m);
    }

    // ------------------------------------------------------------------------------------------ kernel evaluation
    // evaluate_kernel!(K, kernel, F, t) for the whole team (src/Kernels/ModeCouplingKernels.jl:211-228, 450-467).
    // Inputs: sF1 (collective F at this time for tagged kernels, == sF2 otherwise) and sF2 (current iterate), both
    // complete and visible (caller synced).  Output: K[myrow] in the thread that owns the row (others: 0).
    __device__ __forceinline__ double eval_rows(const double *__restrict__ gtab, double kval) {
        PROF(8);
        // P1: stream the table terms of this warp's tasks; consecutive tasks of one row share a partial-sum slot
        {
            const int tb = sTaskBegin[warp], te = sTaskBegin[warp + 1];
            const int st = P.smem_terms, sm = even_up(P.slots_max);
            double a1 = 0.0, a2 = 0.0, a3 = 0.0;
            for (int ti = tb; ti < te; ti++) {
                const Task t = sTasks[ti];
                const double *fa = sF1 + t.a0;
                const double *fb = sF2 + t.b0;
                if (t.where == 0) {
                    const double *t1 = sTab + t.off, *t2 = t1 + st, *t3 = t2 + st;
                    if (t.bdir > 0) {
#pragma unroll 4
                        for (int i = lane; i < t.len; i += 32) {
                            const double f = fa[i] * fb[i];
                            a1 = fma(t1[i], f, a1); a2 = fma(t2[i], f, a2); a3 = fma(t3[i], f, a3);
                        }
                    } else {
#pragma unroll 4
                        for (int i = lane; i < t.len; i += 32) {
                            const double f = fa[i] * fb[-i];
                            a1 = fma(t1[i], f, a1); a2 = fma(t2[i], f, a2); a3 = fma(t3[i], f, a3);
                        }
                    }
                } else {
                    const double *t1 = gtab + t.off, *t2 = t1 + P.total_terms, *t3 = t2 + P.total_terms;
                    if (t.bdir > 0) {
#pragma unroll 4
                        for (int i = lane; i < t.len; i += 32) {
                            const double f = fa[i] * fb[i];
                            a1 = fma(__ldg(t1 + i), f, a1); a2 = fma(__ldg(t2 + i), f, a2); a3 = fma(__ldg(t3 + i), f, a3);
                        }
                    } else {
#pragma unroll 4
                        for (int i = lane; i < t.len; i += 32) {
                            const double f = fa[i] * fb[-i];
                            a1 = fma(__ldg(t1 + i), f, a1); a2 = fma(__ldg(t2 + i), f, a2); a3 = fma(__ldg(t3 + i), f, a3);
                        }
                    }
                }
                if (t.flush) {
                    a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
                    if (lane == 0) { sPart[t.slot] = a1; sPart[sm + t.slot] = a2; sPart[2 * sm + t.slot] = a3; }
                    a1 = 0.0; a2 = 0.0; a3 = 0.0;
                }
            }
        }
        PROF(0);
        __syncthreads();
        PROF(1);
        // P2: increments of this CTA's row parts, inclusive prefix over them, block totals published to the team
        begin_exchange();
        double i1 = 0.0, i2 = 0.0, i3 = 0.0;
        if (tid < nparts) {
            const int sm = even_up(P.slots_max);
            for (int s = sRowSlot[tid]; s < sRowSlot[tid + 1]; s++) { i1 += sPart[s]; i2 += sPart[sm + s]; i3 += sPart[2 * sm + s]; }
        }
        double s1, s2, s3;
        block_inscan3(i1, i2, i3, s1, s2, s3);
        if (G > 1 && tid == 0) {
            double *c = xcur() + 16 * cta;
            st_relaxed(c, s1); st_relaxed(c + 1, s2); st_relaxed(c + 2, s3);
        }
        PROF(2);
        // P3: offset of this CTA = sum of the block totals of the CTAs before it (fixed order: warp tree, then warps)
        double o1 = 0.0, o2 = 0.0, o3 = 0.0;
        if (G > 1) {
            double t1 = 0.0, t2 = 0.0, t3 = 0.0;
            if (tid < cta) {
                const double *c = xcur() + 16 * tid;
                t1 = spin_read(c); t2 = spin_read(c + 1); t3 = spin_read(c + 2);
            }
            PROF(3);
            const int nw = (G + 31) >> 5;                  // warps that may hold totals
            if (warp < nw) {
                t1 = warp_sum(t1); t2 = warp_sum(t2); t3 = warp_sum(t3);
                if (lane == 0) { sRed[warp] = t1; sRed[kWarps + warp] = t2; sRed[2 * kWarps + warp] = t3; }
            }
            __syncthreads();
            for (int w = 0; w < nw; w++) { o1 += sRed[w]; o2 += sRed[kWarps + w]; o3 += sRed[2 * kWarps + w]; }
        }
        // T_m[ik] = offset + local prefix; K = k T1 + T2/k^3 + T3/k   (:224-227)
        double K = 0.0;
        if (myrow >= 0) {
            const double T1 = o1 + i1, T2 = o2 + i2, T3 = o3 + i3;
            K = kval * T1 + T2 / (kval * kval * kval) + T3 / kval;
        }
        PROF(4);
        return K;
    }

    __device__ __forceinline__ void load_tables(const EqDev &E) {
        const int ns = min(P.smem_terms, slab_len);
        const double *g = E.tab + slab_off;
        for (int m = 0; m < 3; m++)
            for (int i = tid; i < ns; i += kThreads) sTab[(size_t)m * P.smem_terms + i] = __ldg(g + (size_t)m * P.total_terms + i);
    }
    // F1 operand of a tagged kernel (collective F at trajectory index tix); caller syncs.
    __device__ __forceinline__ void load_F1(const EqDev &E, long long tix) {
        if (E.Fc_traj == nullptr) return;
        const double *src = E.Fc_traj + tix * Nk;
        for (int k = tid; k < Nk; k += kThreads) sF1[k] = __ldg(src + k);
    }

    // ------------------------------------------------------------------------------------------ steady state
    // solve_steady_state_mutable (src/SteadyStateMemoryEquation.jl:59-100) for a diagonal kernel:
    // iterate F <- (K + gamma)^-1 K F0, K <- kernel(F) until max|F - F_old| <= tolerance.
    __device__ __forceinline__ void steady_state_one(const EqDev &E) {
        sF1 = sF2;
        load_tables(E);
        const double *gtab = E.tab + slab_off;
        const int k = myrow;
        const double kval = (k >= 0) ? E.kvec[k] : 1.0, f0 = (k >= 0) ? E.F0[k] : 0.0, g = (k >= 0) ? E.gamma[k] : 1.0;
        double fold[KPT];
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            const int kk = tid + j * kThreads;
            fold[j] = (kk < Nk) ? E.F0[kk] : 0.0;              // Fold = copy(F0) :68
            if (kk < Nk) sF2[kk] = fold[j];
        }
        __syncthreads();
        double K = eval_rows(gtab, kval);                      // K = evaluate_kernel(kernel, F0, t) :66
        long long iter = 0;
        int status = MCT_OK;
        double err = P.tol * 2.0;
        while (err > P.tol) {
            iter++;
            if (iter > P.max_iter) { status = MCT_NOT_CONVERGED; break; }   // :75-78
            const double Fn = (K * f0) / (K + g);              // :81-84
            err = allgather_F(Fn, true, fold);                 // find_error(F, Fold); Fold .= F  :90-91
            K = eval_rows(gtab, kval);                         // :89 (the error does not depend on it)
            if (aborted()) { status = MCT_CUDA_ERROR; break; }
        }
        if (k >= 0) { E.F_out[k] = sF2[k]; E.K_out[k] = K; }
        if (cta == 0 && tid == 0) { *E.status = status; *E.kernel_evals = iter; }
    }

    // ------------------------------------------------------------------------------------------ time-doubling solve
    __device__ __forceinline__ void solve_one(const EqDev &E) {
        const int N = P.N, n4 = 4 * N, n2 = 2 * N;
        double *__restrict__ Ft = E.ring;
        double *__restrict__ Kt = Ft + (size_t)Nk * n4;
        double *__restrict__ FI = Kt + (size_t)Nk * n4;
        double *__restrict__ KI = FI + (size_t)Nk * n4;
#define RING(R, k, s) (R)[(size_t)(k) * n4 + ((s) - 1)]   /* slot s is 1-based as in the reference */
        sF1 = (E.Fc_traj != nullptr) ? sF1_home : sF2;    // tagged kernels: A = V * Fc[iq] * Fs[ip]
        const double *gtab = E.tab + slab_off;
        load_tables(E);
        const int k = myrow;                              // the wavenumber this thread owns (or -1)
        const double kval = (k >= 0) ? E.kvec[k] : 1.0, f0 = (k >= 0) ? E.F0[k] : 0.0;
        const double al = (k >= 0) ? E.alpha[k] : 0.0, be = (k >= 0) ? E.beta[k] : 1.0, ga = (k >= 0) ? E.gamma[k] : 0.0,
                     de_k = (k >= 0) ? E.delta[k] : 0.0;
        double fold[KPT];
        // K0 = evaluate_kernel(kernel, F0, 0.0)   (src/MemoryEquation.jl:45)
        for (int kk = tid; kk < Nk; kk += kThreads) sF2[kk] = E.F0[kk];
        load_F1(E, 0);
        __syncthreads();
        const double K0 = eval_rows(gtab, kval);
        if (k >= 0) { E.F_out[k] = f0; E.K_out[k] = K0; }     // output row 0 (t = 0)
        // F_old = K0*F0 (:64) for every wavenumber: gather K0*F0 once so that each CTA holds the full vector
        allgather_F(K0 * f0, false, fold);
#pragma unroll
        for (int j = 0; j < KPT; j++) { const int kk = tid + j * kThreads; fold[j] = (kk < Nk) ? sF2[kk] : 0.0; }
        __syncthreads();

        // ---- bootstrap: 2N forward-Euler steps with the trapezoid memory integral
        //      (initialize_F_temp_Euler! :96-109 -> src/EulerSolver.jl:26-64, 92-112); the dF history lives in the
        //      F_I ring until initialize_integrals!.
        const double de = P.dt0 / (4.0 * N);
        if (k >= 0) FI[(size_t)k * n4 + 0] = E.dF0[k];
        for (int it = 1; it <= n2; it++) {
            double Fn = 0.0;
            if (k >= 0) {
                const double *dF = FI + (size_t)k * n4;       // dF[j] = dF at Euler step j
                double integ;
                // integrand[i] = K_array[i] * dF_reverse[i], K_array[i] = K(t_{i-1}), dF_reverse[i] = dF(t_{it-i})
                if (it == 1) {
                    integ = (K0 * dF[0]) * de;                 // trapezoidal_integration it == 1 (EulerSolver.jl:34-36)
                } else {
                    double rr = (K0 * dF[it - 1]) / 2.0;
                    rr += (RING(Kt, k, it - 1) * dF[0]) / 2.0;
                    for (int i = 2; i <= it - 1; i++) rr += RING(Kt, k, i - 1) * dF[it - i];
                    integ = rr * de;
                }
                const double Fo = (it == 1) ? f0 : RING(Ft, k, it - 1);
                const double dFo = dF[it - 1];
                double dFn;
                if (P.second_order) {                          // Euler_step (EulerSolver.jl:55-58)
                    const double ddF = (be * dFo + ga * Fo + de_k + integ) / (-al);
                    dFn = dFo + de * ddF;
                    Fn = Fo + de * dFn;
                } else {                                       // :60-61
                    dFn = (ga * Fo + de_k + integ) / (-be);
                    Fn = Fo + de * dFn;
                }
                RING(Ft, k, it) = Fn;
                FI[(size_t)k * n4 + it] = dFn;
            }
            double dummy[KPT];
            allgather_F(Fn, false, dummy);
            load_F1(E, it);
            __syncthreads();
            const double K = eval_rows(gtab, kval);
            if (k >= 0) {
                RING(Kt, k, it) = K;                           // == initialize_K_temp! :155-166
                E.F_out[(size_t)it * Nk + k] = Fn; E.K_out[(size_t)it * Nk + k] = K;
            }
            if (aborted()) return;
        }
        // initialize_integrals! :174-199 (overwrites the dF scratch, which is no longer needed)
        if (k >= 0) {
            const double F1v = RING(Ft, k, 1), K1v = RING(Kt, k, 1), K2v = RING(Kt, k, 2);
            double Fprev = F1v, Kprev = K1v;
            RING(FI, k, 1) = (F1v + f0) / 2.0;
            RING(KI, k, 1) = (3.0 * K1v - K2v) / 2.0;
            for (int it = 2; it <= n2; it++) {
                const double Fv = RING(Ft, k, it), Kv = RING(Kt, k, it);
                RING(FI, k, it) = (Fv + Fprev) / 2.0;
                RING(KI, k, it) = (Kv + Kprev) / 2.0;
                Fprev = Fv; Kprev = Kv;
            }
        }
        __syncthreads();

        // ---- main loop: while solver.dt < 2 t_max (:523); block b works with dt = dt0 * 2^b
        long long evals = 0;
        int status = MCT_OK;
        double Dt = P.dt0;
        for (int blk = 0; blk < P.nblocks && status == MCT_OK; blk++, Dt *= 2.0) {
            const double dl = Dt / (4.0 * N);
            for (int it = n2 + 1; it <= n4 && status == MCT_OK; it++) {
                const int i2 = it / 2;
                const long long row = 1 + (long long)n2 * (blk + 1) + (it - n2 - 1);
                // update_Fuchs_parameters! :300-319 for the owned wavenumbers: one warp per row, lanes over the history
                for (int r = warp; r < nparts; r += kWarps) {
                    const int kr = sRows[r];
                    if (kr < 0) continue;
                    double s1 = 0.0, s2 = 0.0;
                    for (int j = 2 + lane; j <= i2; j += 32) s1 += (RING(Kt, kr, it - j) - RING(Kt, kr, it - j + 1)) * RING(FI, kr, j);        // :271-280
                    for (int j = 2 + lane; j <= it - i2; j += 32) s2 += RING(KI, kr, j) * (RING(Ft, kr, it - j) - RING(Ft, kr, it - j + 1));  // :281-285
                    s1 = warp_sum(s1); s2 = warp_sum(s2);
                    if (lane == 0) { sC[r] = s1; sC[P.rows_max + r] = s2; }
                }
                __syncthreads();
                double c1 = 1.0, c2 = 0.0, c3 = 0.0, Finit = 0.0;
                if (k >= 0) {
                    const double Fm1 = RING(Ft, k, it - 1), Fm2 = RING(Ft, k, it - 2), Fm3 = RING(Ft, k, it - 3);
                    const double KI1 = RING(KI, k, 1), FI1 = RING(FI, k, 1);
                    double v = al * ((5.0 * Fm1 - 4.0 * Fm2 + Fm3) / (dl * dl));                // :259-260
                    v += be * (2.0 / dl * Fm1 - Fm2 / (2.0 * dl));                              // :263-264
                    v -= RING(Kt, k, it - i2) * RING(Ft, k, i2);                                // :267
                    v += RING(Kt, k, it - 1) * FI1;                                             // :268
                    v += KI1 * Fm1;                                                             // :269
                    v += sC[tid]; v += sC[P.rows_max + tid];
                    v -= de_k;                                                                  // :286
                    c3 = v;
                    c1 = ((2.0 / (dl * dl) * al + 3.0 / (2.0 * dl) * be) + KI1) + ga;           // :314
                    c2 = FI1 - f0;                                                              // :315
                    // update_F! :326-337 with the stale K_temp[it]: 2 K0 in the first block (:83), 0 afterwards (:485)
                    const double Kst = (blk == 0) ? (K0 + K0) : 0.0;
                    Finit = (-(Kst * c2) + c3) / c1;
                }
                double dummy[KPT];
                allgather_F(Finit, false, dummy);
                load_F1(E, row);
                __syncthreads();
                // fixed-point loop :404-417
                long long iter = 0;
                double K = 0.0, Fn = Finit;
                while (true) {
                    iter++;
                    if (iter > P.max_iter) { status = MCT_NOT_CONVERGED; break; }                // :406-408
                    K = eval_rows(gtab, kval);                                                  // update_K! :360-369
                    Fn = (-(K * c2) + c3) / c1;                                                 // update_F! :333-336
                    const double err = allgather_F(Fn, true, fold);                             // find_error + F_old .= F
                    evals++;                                                                    // :416
                    if (aborted()) { status = MCT_CUDA_ERROR; break; }
                    if (!(err > P.tol)) break;
                }
                if (status != MCT_OK) break;
                // store the step + update_integrals! :377-384, allocate_results! :429-438 (owners)
                if (k >= 0) {
                    RING(Ft, k, it) = Fn; RING(Kt, k, it) = K;
                    RING(FI, k, it) = (Fn + RING(Ft, k, it - 1)) / 2.0;
                    RING(KI, k, it) = (K + RING(Kt, k, it - 1)) / 2.0;
                    E.F_out[(size_t)row * Nk + k] = Fn; E.K_out[(size_t)row * Nk + k] = K;
                }
                __syncthreads();
            }
            if (status != MCT_OK) break;
            // new_time_mapping! :448-489 (owners, one warp per row; read everything of a chunk before writing it)
            for (int r = warp; r < nparts; r += kWarps) {
                const int kr = sRows[r];
                if (kr < 0) continue;
                for (int jb = 1; jb <= n2; jb += 32) {
                    const int j = jb + lane;
                    double a = 0, b = 0, d = 0, e = 0;
                    if (j <= n2) {
                        a = (RING(FI, kr, 2 * j) + RING(FI, kr, 2 * j - 1)) / 2.0;
                        b = (RING(KI, kr, 2 * j) + RING(KI, kr, 2 * j - 1)) / 2.0;
                        d = RING(Ft, kr, 2 * j);
                        e = RING(Kt, kr, 2 * j);
                    }
                    __syncwarp();
                    if (j <= n2) { RING(FI, kr, j) = a; RING(KI, kr, j) = b; RING(Ft, kr, j) = d; RING(Kt, kr, j) = e; }
                    __syncwarp();
                }
                for (int j = n2 + 1 + lane; j <= n4; j += 32) { RING(FI, kr, j) = 0.0; RING(KI, kr, j) = 0.0; RING(Ft, kr, j) = 0.0; RING(Kt, kr, j) = 0.0; }
            }
            __syncthreads();
        }
        if (cta == 0 && tid == 0) { *E.status = status; *E.kernel_evals = evals; }
#undef RING
    }
};

template <int KPT>
__global__ void __launch_bounds__(kThreads, 1) td_sc_kernel(const SolveParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Team<KPT> T(P);
    T.tid = threadIdx.x; T.lane = T.tid & 31; T.warp = T.tid >> 5;
    T.team = blockIdx.x / P.G; T.cta = blockIdx.x % P.G;
    T.Nk = P.Nk; T.NkP = even_up(P.Nk); T.G = P.G;
    T.epoch = 0;
    T.xbuf_doubles = P.xbuf_words;
    T.xbase = P.xbuf + (size_t)T.team * P.xstride;
    const SmemPlan pl = plan_smem(P.Nk, P.G, P.slots_max, P.tasks_max, P.rows_max, P.smem_terms);
    T.sF1_home = (double *)(smem_raw + pl.off_F1); T.sF1 = T.sF1_home;
    T.sF2 = (double *)(smem_raw + pl.off_F2); T.sC = (double *)(smem_raw + pl.off_K);
    T.sX = (double *)(smem_raw + pl.off_X); T.sPart = (double *)(smem_raw + pl.off_part);
    T.sScan = (double *)(smem_raw + pl.off_scan); T.sMax = (double *)(smem_raw + pl.off_mx);
    T.sRed = (double *)(smem_raw + pl.off_red);
    Task *sTasks = (Task *)(smem_raw + pl.off_tasks);
    int *sTaskBegin = (int *)(smem_raw + pl.off_tb), *sRows = (int *)(smem_raw + pl.off_rows), *sRowSlot = (int *)(smem_raw + pl.off_rs);
    T.sTasks = sTasks; T.sTaskBegin = sTaskBegin; T.sRows = sRows; T.sRowSlot = sRowSlot;
    T.sTab = (double *)(smem_raw + pl.off_tab);
    T.sAbort = (volatile int *)(smem_raw + pl.off_abort);
    T.slab_off = P.cta_off[T.cta];
    T.slab_len = (int)(P.cta_off[T.cta + 1] - P.cta_off[T.cta]);
#ifdef MCT_PROFILE
    for (int i = 0; i < 12; i++) T.prof[i] = 0;
    T.tlast = clock64();
    const long long t_begin = T.tlast;
    long long g_begin;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g_begin));
#endif
    // per-CTA slice of the layout geometry -> shared memory (identical for every equation of this launch)
    {
        const int tb0 = P.task_begin[T.cta * kWarps], tb1 = P.task_begin[(T.cta + 1) * kWarps];
        for (int i = T.tid; i < tb1 - tb0; i += kThreads) sTasks[i] = P.tasks[tb0 + i];
        for (int i = T.tid; i <= kWarps; i += kThreads) sTaskBegin[i] = P.task_begin[T.cta * kWarps + i] - tb0;
        int nr = 0;
        for (int i = 0; i < P.rows_max; i++) nr += (P.rows[T.cta * P.rows_max + i] != -1) ? 1 : 0;
        T.nparts = nr;   // row parts come first, -1 pads the tail
        const int rv = (T.tid < P.rows_max) ? P.rows[T.cta * P.rows_max + T.tid] : -1;
        T.myrow = (rv >= 0) ? rv : -1;   // -(ik+2): partial row whose tail (and ownership) lives in a later CTA
        T.myslot = (rv >= 0) ? P.kslot[rv] : 0;
#pragma unroll
        for (int j = 0; j < KPT; j++) { const int k = T.tid + j * kThreads; T.rslot[j] = (k < P.Nk) ? P.kslot[k] : 0; }
        for (int i = T.tid; i < P.rows_max; i += kThreads) sRows[i] = P.rows[T.cta * P.rows_max + i];
        for (int i = T.tid; i <= P.rows_max; i += kThreads) sRowSlot[i] = P.row_slot[T.cta * (P.rows_max + 1) + i];
        if (T.tid == 0) *T.sAbort = 0;
    }
    __syncthreads();
    while (true) {
        // next equation of this team (dynamic queue); the index travels through the exchange like any other value
        int e;
        T.begin_exchange();
        if (T.G == 1) {
            if (T.tid == 0) T.sX[0] = (double)atomicAdd(P.queue, 1);
            __syncthreads();
            e = (int)T.sX[0];
        } else {
            double *c = T.xcur() + 3;                         // the spare 4th word of CTA 0's totals line
            if (T.cta == 0 && T.tid == 0) st_relaxed(c, (double)atomicAdd(P.queue, 1));
            e = (int)T.spin_read(c);
        }
        __syncthreads();
        if (T.G > 1 && T.cta == 0 && T.tid == 0) st_relaxed_bits(T.xbase + (size_t)((T.epoch + 2u) & 3u) * T.xbuf_doubles + 3, kSentinel);
        if (e >= P.n_eq || T.aborted()) break;
        const EqDev E = P.eqs[e];
        if (E.mode == 1) T.steady_state_one(E);
        else T.solve_one(E);
        if (T.aborted()) {
            if (T.cta == 0 && T.tid == 0) *E.status = MCT_CUDA_ERROR;
            break;
        }
        __syncthreads();
    }
#ifdef MCT_PROFILE
    if (T.team == 0 && T.tid == 0 && P.prof) {
        long long g_end;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g_end));
        T.prof[10] = clock64() - t_begin;
        T.prof[11] = g_end - g_begin;
        for (int i = 0; i < 12; i++) P.prof[T.cta * 12 + i] = T.prof[i];
    }
#endif
}

}  // namespace

size_t td_sc_smem_bytes(int Nk, int G, int slots_max, int tasks_max, int rows_max, int smem_terms) {
    return plan_smem(Nk, G, slots_max, tasks_max, rows_max, smem_terms).total;
}

int td_sc_max_smem_terms(int Nk, int G, int slots_max_guess, int tasks_max_guess, int rows_max, size_t smem_limit) {
    size_t fixed = plan_smem(Nk, G, slots_max_guess, tasks_max_guess, rows_max, 0).total;
    if (fixed + 24 * 64 > smem_limit) return 0;
    return (int)((smem_limit - fixed) / 24) & ~1;
}

int td_sc_launch(const SolveParams &P, cudaStream_t stream) {
    const int KPT = (P.Nk + kThreads - 1) / kThreads;
    MCT_REQUIRE(KPT <= kMaxKPT, MCT_UNSUPPORTED, "Nk > 2048 is not supported by the persistent solver");
    const size_t smem = td_sc_smem_bytes(P.Nk, P.G, P.slots_max, P.tasks_max, P.rows_max, P.smem_terms);
    void *fn = nullptr;
    if (KPT <= 1) fn = (void *)td_sc_kernel<1>;
    else if (KPT <= 2) fn = (void *)td_sc_kernel<2>;
    else fn = (void *)td_sc_kernel<4>;
    MCT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SolveParams Pc = P;
    void *args[] = {(void *)&Pc};
    dim3 grid(P.nteams * P.G), block(kThreads);
    if (P.G > 1) {
        // the exchange spins on other CTAs of the team: a cooperative launch guarantees co-residency or fails.
        MCT_CUDA(cudaLaunchCooperativeKernel(fn, grid, block, args, smem, stream));
    } else {
        MCT_CUDA(cudaLaunchKernel(fn, grid, block, args, smem, stream));
    }
    return MCT_OK;
}

}  // namespace mct
