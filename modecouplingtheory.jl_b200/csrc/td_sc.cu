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
//     sharded over the team by wavenumber ("owners") and never crosses CTAs, except C1..C3 once per time step.
// Arithmetic outside the table stream follows the reference's operation order (file compiled with -fmad=false).
#include <algorithm>

#include "td_sc.cuh"

namespace mct {

namespace {

struct Smem {
    double *F1, *F1_home, *F2, *K, *X, *part, *scan, *mx, *tab;
    Task *tasks;
    int *task_begin, *rows, *row_slot, *abort;
};

__host__ __device__ inline int even_up(int n) { return (n + 1) & ~1; }

struct SmemPlan {
    size_t off_F1, off_F2, off_K, off_X, off_part, off_scan, off_mx, off_tasks, off_tb, off_rows, off_rs, off_abort, off_tab, total;
};

__host__ __device__ inline SmemPlan plan_smem(int Nk, int G, int slots_max, int tasks_max, int rows_max, int smem_terms) {
    SmemPlan p;
    size_t o = 0;
    int NkP = even_up(Nk);
    p.off_F1 = o; o += (size_t)NkP * 8;
    p.off_F2 = o; o += (size_t)NkP * 8;
    p.off_K = o; o += (size_t)NkP * 8;
    p.off_X = o; o += (G == 1 ? (size_t)3 * NkP * 8 : 0);
    p.off_part = o; o += (size_t)3 * even_up(slots_max) * 8;
    p.off_scan = o; o += (size_t)3 * kWarps * 8;
    p.off_mx = o; o += (size_t)kWarps * 8;
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

struct Ctx {
    const SolveParams &P;
    Smem s;
    int team, cta, tid, lane, warp;
    unsigned long long epoch;     // team-barrier epoch (monotonic for the whole launch)
    double *xb[2];                // exchange buffers of this team
    unsigned long long *flags;    // barrier flags of this team
    long long slab_off;           // this CTA's slab inside the team layout
    int slab_len;
    __device__ Ctx(const SolveParams &p) : P(p) {}
};

// ---- team barrier: every CTA publishes an epoch flag and polls all G flags (no atomics on the critical path).
__device__ __forceinline__ void team_barrier(Ctx &c) {
    __syncthreads();
    if (c.P.G == 1) return;
    c.epoch++;
    if (c.tid == 0) {
        __threadfence();
        st_release_gpu(c.flags + (size_t)c.cta * kFlagStride, c.epoch);
    }
    if (c.tid < c.P.G) {
        const unsigned long long *f = c.flags + (size_t)c.tid * kFlagStride;
        long long t0 = clock64();
        unsigned spins = 0;
        while (ld_relaxed_gpu(f) < c.epoch) {
            if ((++spins & 255u) == 0) {
                if (*(volatile int *)c.P.abort_flag) { *c.s.abort = 1; break; }
                if (clock64() - t0 > c.P.spin_limit_cycles) { atomicExch(c.P.abort_flag, 1); *c.s.abort = 1; break; }
            }
        }
        __threadfence();
    }
    __syncthreads();
}

// set (in shared memory) by a barrier poller that saw the launch-wide abort flag or timed out
__device__ __forceinline__ bool team_aborted(Ctx &c) { return *(volatile int *)c.s.abort != 0; }

// exclusive block scan of three per-thread totals (all kThreads threads participate, one __syncthreads inside)
__device__ __forceinline__ void block_exscan3(Ctx &c, double t0, double t1, double t2, double &o0, double &o1, double &o2) {
    double v0 = t0, v1 = t1, v2 = t2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        double u0 = __shfl_up_sync(0xffffffffu, v0, o), u1 = __shfl_up_sync(0xffffffffu, v1, o), u2 = __shfl_up_sync(0xffffffffu, v2, o);
        if (c.lane >= o) { v0 += u0; v1 += u1; v2 += u2; }
    }
    if (c.lane == 31) { c.s.scan[c.warp] = v0; c.s.scan[kWarps + c.warp] = v1; c.s.scan[2 * kWarps + c.warp] = v2; }
    double e0 = __shfl_up_sync(0xffffffffu, v0, 1), e1 = __shfl_up_sync(0xffffffffu, v1, 1), e2 = __shfl_up_sync(0xffffffffu, v2, 1);
    if (c.lane == 0) { e0 = 0.0; e1 = 0.0; e2 = 0.0; }
    __syncthreads();
    double b0 = 0.0, b1 = 0.0, b2 = 0.0;
    for (int w = 0; w < c.warp; w++) { b0 += c.s.scan[w]; b1 += c.s.scan[kWarps + w]; b2 += c.s.scan[2 * kWarps + w]; }
    o0 = b0 + e0; o1 = b1 + e1; o2 = b2 + e2;
}

__device__ __forceinline__ double block_max(Ctx &c, double v) {
    v = warp_max(v);
    if (c.lane == 0) c.s.mx[c.warp] = v;
    __syncthreads();
    double m = c.s.mx[0];
#pragma unroll
    for (int w = 1; w < kWarps; w++) m = fmax(m, c.s.mx[w]);
    return m;
}

// ---- evaluate_kernel!(K, kernel, F, t) for the whole team (src/Kernels/ModeCouplingKernels.jl:211-228, 450-467).
// Inputs: s.F1 (collective F at this time for tagged kernels, == s.F2 otherwise) and s.F2 (current iterate).
// Output: K for the KPT wavenumbers of this thread (kk = tid*KPT + j).
template <int KPT>
__device__ __forceinline__ void team_eval(Ctx &c, const EqDev &E, const double (&kval)[KPT], double (&Kreg)[KPT]) {
    const SolveParams &P = c.P;
    const int Nk = P.Nk;
    // P1: stream the table terms of this warp's tasks
    {
        const int tb = c.s.task_begin[c.warp], te = c.s.task_begin[c.warp + 1];
        const double *gtab = E.tab + c.slab_off;
        const int st = P.smem_terms;
        for (int ti = tb; ti < te; ti++) {
            const Task t = c.s.tasks[ti];
            const double *fa = c.s.F1 + t.a0;
            const double *fb = c.s.F2 + t.b0;
            double a1 = 0.0, a2 = 0.0, a3 = 0.0;
            if (t.where == 0) {
                const double *t1 = c.s.tab + t.off, *t2 = t1 + st, *t3 = t2 + st;
                if (t.bdir > 0) {
#pragma unroll 4
                    for (int i = c.lane; i < t.len; i += 32) {
                        double f = fa[i] * fb[i];
                        a1 = fma(t1[i], f, a1); a2 = fma(t2[i], f, a2); a3 = fma(t3[i], f, a3);
                    }
                } else {
#pragma unroll 4
                    for (int i = c.lane; i < t.len; i += 32) {
                        double f = fa[i] * fb[-i];
                        a1 = fma(t1[i], f, a1); a2 = fma(t2[i], f, a2); a3 = fma(t3[i], f, a3);
                    }
                }
            } else {
                const double *t1 = gtab + t.off, *t2 = t1 + P.total_terms, *t3 = t2 + P.total_terms;
                if (t.bdir > 0) {
#pragma unroll 4
                    for (int i = c.lane; i < t.len; i += 32) {
                        double f = fa[i] * fb[i];
                        a1 = fma(__ldg(t1 + i), f, a1); a2 = fma(__ldg(t2 + i), f, a2); a3 = fma(__ldg(t3 + i), f, a3);
                    }
                } else {
#pragma unroll 4
                    for (int i = c.lane; i < t.len; i += 32) {
                        double f = fa[i] * fb[-i];
                        a1 = fma(__ldg(t1 + i), f, a1); a2 = fma(__ldg(t2 + i), f, a2); a3 = fma(__ldg(t3 + i), f, a3);
                    }
                }
            }
            a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
            if (c.lane == 0) {
                const int sm = even_up(P.slots_max);
                c.s.part[t.slot] = a1; c.s.part[sm + t.slot] = a2; c.s.part[2 * sm + t.slot] = a3;
            }
        }
    }
    __syncthreads();
    // P2: per-row increments inc_m[ik] = T_m[ik] - T_m[ik-1], published to the team
    const int par = (int)((c.epoch + 1) & 1);
    double *xw = (P.G == 1) ? c.s.X : c.xb[par];
    const int xs = (P.G == 1) ? even_up(Nk) : Nk;
    for (int r = c.tid; r < P.rows_max; r += kThreads) {
        const int ik = c.s.rows[r];
        if (ik >= 0) {
            const int sm = even_up(P.slots_max);
            double i1 = 0.0, i2 = 0.0, i3 = 0.0;
            for (int s = c.s.row_slot[r]; s < c.s.row_slot[r + 1]; s++) { i1 += c.s.part[s]; i2 += c.s.part[sm + s]; i3 += c.s.part[2 * sm + s]; }
            if (P.G == 1) { xw[ik] = i1; xw[xs + ik] = i2; xw[2 * xs + ik] = i3; }
            else { st_cg(xw + ik, i1); st_cg(xw + xs + ik, i2); st_cg(xw + 2 * xs + ik, i3); }
        }
    }
    team_barrier(c);
    // P3: every CTA scans all increments -> T1..T3 -> K   (:224-227)
    double l1[KPT], l2[KPT], l3[KPT];
    double r1 = 0.0, r2 = 0.0, r3 = 0.0;
#pragma unroll
    for (int j = 0; j < KPT; j++) {
        const int k = c.tid * KPT + j;
        double i1 = 0.0, i2 = 0.0, i3 = 0.0;
        if (k < Nk) {
            if (P.G == 1) { i1 = xw[k]; i2 = xw[xs + k]; i3 = xw[2 * xs + k]; }
            else { i1 = ld_cg(xw + k); i2 = ld_cg(xw + xs + k); i3 = ld_cg(xw + 2 * xs + k); }
        }
        r1 += i1; r2 += i2; r3 += i3;
        l1[j] = r1; l2[j] = r2; l3[j] = r3;
    }
    double o1, o2, o3;
    block_exscan3(c, r1, r2, r3, o1, o2, o3);
#pragma unroll
    for (int j = 0; j < KPT; j++) {
        const double T1 = o1 + l1[j], T2 = o2 + l2[j], T3 = o3 + l3[j];
        const double k = kval[j];
        Kreg[j] = k * T1 + T2 / (k * k * k) + T3 / k;
    }
}

// Load the F1 operand of a tagged kernel (collective F at trajectory index tix); caller syncs.
__device__ __forceinline__ void load_F1(Ctx &c, const EqDev &E, long long tix) {
    if (E.Fc_traj == nullptr) return;
    const double *src = E.Fc_traj + tix * c.P.Nk;
    for (int k = c.tid; k < c.P.Nk; k += kThreads) c.s.F1[k] = __ldg(src + k);
}

// Row `row` of the output trajectory: written by one CTA of the team, coalesced.
__device__ __forceinline__ void write_output_row(Ctx &c, const EqDev &E, long long row) {
    if ((int)(row % c.P.G) != c.cta) return;
    double *Fo = E.F_out + row * c.P.Nk, *Ko = E.K_out + row * c.P.Nk;
    for (int k = c.tid; k < c.P.Nk; k += kThreads) { Fo[k] = c.s.F2[k]; Ko[k] = c.s.K[k]; }
}

// Tables of this CTA -> shared memory, where they stay resident for the whole solve.
__device__ __forceinline__ void load_tables(Ctx &c, const EqDev &E) {
    const SolveParams &P = c.P;
    const int ns = min(P.smem_terms, c.slab_len);
    const double *g = E.tab + c.slab_off;
    for (int m = 0; m < 3; m++)
        for (int i = c.tid; i < ns; i += kThreads) c.s.tab[(size_t)m * P.smem_terms + i] = __ldg(g + (size_t)m * P.total_terms + i);
}

// solve_steady_state_mutable (src/SteadyStateMemoryEquation.jl:59-100) for a diagonal kernel:
// iterate F <- (K + gamma)^-1 K F0, K <- kernel(F) until max|F - F_old| <= tolerance.
template <int KPT>
__device__ void steady_state_one(Ctx &c, const EqDev E) {
    const SolveParams &P = c.P;
    const int Nk = P.Nk;
    c.s.F1 = c.s.F2;
    load_tables(c, E);
    double kval[KPT], f0[KPT], g[KPT], fold[KPT], Kreg[KPT];
#pragma unroll
    for (int j = 0; j < KPT; j++) {
        const int k = c.tid * KPT + j;
        kval[j] = (k < Nk) ? E.kvec[k] : 1.0;
        f0[j] = (k < Nk) ? E.F0[k] : 0.0;
        g[j] = (k < Nk) ? E.gamma[k] : 1.0;
        fold[j] = f0[j];                                   // Fold = copy(F0) :68
        if (k < Nk) c.s.F2[k] = f0[j];
    }
    __syncthreads();
    team_eval<KPT>(c, E, kval, Kreg);                      // K = evaluate_kernel(kernel, F0, t) :66
    long long iter = 0;
    int status = MCT_OK;
    double err = P.tol * 2.0;
    while (err > P.tol) {
        iter++;
        if (iter > P.max_iter) { status = MCT_NOT_CONVERGED; break; }   // :75-78
        double e = 0.0, Fn[KPT];
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            const int k = c.tid * KPT + j;
            Fn[j] = (Kreg[j] * f0[j]) / (Kreg[j] + g[j]);  // :81-84
            if (k < Nk) c.s.F2[k] = Fn[j];
        }
        __syncthreads();
        team_eval<KPT>(c, E, kval, Kreg);                  // :89
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            const int k = c.tid * KPT + j;
            if (k < Nk) e = fmax(e, fabs(Fn[j] - fold[j]));   // :90
            fold[j] = Fn[j];
        }
        err = block_max(c, e);
        if (team_aborted(c)) { status = MCT_CUDA_ERROR; break; }
        __syncthreads();
    }
    if (c.cta == 0) {
#pragma unroll
        for (int j = 0; j < KPT; j++) {
            const int k = c.tid * KPT + j;
            if (k < Nk) { E.F_out[k] = c.s.F2[k]; E.K_out[k] = Kreg[j]; }
        }
        if (c.tid == 0) { *E.status = status; *E.kernel_evals = iter; }
    }
}

template <int KPT>
__device__ void solve_one(Ctx &c, const EqDev E) {
    const SolveParams &P = c.P;
    const int Nk = P.Nk, N = P.N, n4 = 4 * N, n2 = 2 * N;
    const int KO = (Nk + P.G - 1) / P.G;                 // owned wavenumbers per CTA
    const int kob = min(Nk, c.cta * KO), koe = min(Nk, kob + KO);
    double *Ft = E.ring, *Kt = Ft + (size_t)Nk * n4, *FI = Kt + (size_t)Nk * n4, *KI = FI + (size_t)Nk * n4;
#define RING(R, k, s) (R)[(size_t)(k) * n4 + ((s) - 1)]   /* slot s is 1-based as in the reference */
    c.s.F1 = (E.Fc_traj != nullptr) ? c.s.F1_home : c.s.F2;   // tagged kernels: A = V * Fc[iq] * Fs[ip]

    load_tables(c, E);
    double kval[KPT], c1[KPT], c2[KPT], c3[KPT], fold[KPT], k0[KPT], f0[KPT], Kreg[KPT];
#pragma unroll
    for (int j = 0; j < KPT; j++) {
        const int k = c.tid * KPT + j;
        kval[j] = (k < Nk) ? E.kvec[k] : 1.0;
        f0[j] = (k < Nk) ? E.F0[k] : 0.0;
        c1[j] = 1.0; c2[j] = 0.0; c3[j] = 0.0;
    }
    // K0 = evaluate_kernel(kernel, F0, 0.0)   (src/MemoryEquation.jl:45)
    for (int k = c.tid; k < Nk; k += kThreads) c.s.F2[k] = E.F0[k];
    load_F1(c, E, 0);
    __syncthreads();
    team_eval<KPT>(c, E, kval, Kreg);
#pragma unroll
    for (int j = 0; j < KPT; j++) {
        const int k = c.tid * KPT + j;
        k0[j] = Kreg[j];
        fold[j] = k0[j] * f0[j];                          // F_old = K0*F0   (:64)
        if (k < Nk) c.s.K[k] = Kreg[j];
    }
    __syncthreads();
    write_output_row(c, E, 0);

    // ---- bootstrap: 2N forward-Euler steps with the trapezoid memory integral
    //      (initialize_F_temp_Euler! :96-109 -> src/EulerSolver.jl:26-64, 92-112); dF history lives in the F_I ring.
    const double de = P.dt0 / (4.0 * N);
    for (int k = kob + c.tid; k < koe; k += kThreads) {
        FI[(size_t)k * n4 + 0] = E.dF0[k];
        KI[(size_t)k * n4 + 0] = c.s.K[k];               // K0 of the owned wavenumbers (K_I ring is free until :174)
    }
    __syncthreads();
    for (int it = 1; it <= n2; it++) {
        for (int k = kob + c.tid; k < koe; k += kThreads) {
            const double *dF = FI + (size_t)k * n4;       // dF[j] = dF at Euler step j
            const double Kzero = KI[(size_t)k * n4 + 0];
            double integ;
            // integrand[i] = K_array[i] * dF_reverse[i], K_array[i] = K(t_{i-1}), dF_reverse[i] = dF(t_{it-i})
            if (it == 1) {
                integ = (Kzero * dF[0]) * de;              // trapezoidal_integration it == 1 (EulerSolver.jl:34-36)
            } else {
                double r = (Kzero * dF[it - 1]) / 2.0;
                r += (RING(Kt, k, it - 1) * dF[0]) / 2.0;
                for (int i = 2; i <= it - 1; i++) r += RING(Kt, k, i - 1) * dF[it - i];
                integ = r * de;
            }
            const double Fo = (it == 1) ? E.F0[k] : RING(Ft, k, it - 1);
            const double dFo = dF[it - 1];
            double Fn, dFn;
            if (P.second_order) {                          // Euler_step (EulerSolver.jl:55-58)
                const double ddF = (E.beta[k] * dFo + E.gamma[k] * Fo + E.delta[k] + integ) / (-E.alpha[k]);
                dFn = dFo + de * ddF;
                Fn = Fo + de * dFn;
            } else {                                       // :60-61
                dFn = (E.gamma[k] * Fo + E.delta[k] + integ) / (-E.beta[k]);
                Fn = Fo + de * dFn;
            }
            RING(Ft, k, it) = Fn;
            FI[(size_t)k * n4 + it] = dFn;
            if (P.G > 1) st_cg(c.xb[(c.epoch + 1) & 1] + k, Fn);
            else c.s.F2[k] = Fn;
        }
        team_barrier(c);
        if (P.G > 1) {
            const double *xr = c.xb[c.epoch & 1];
            for (int k = c.tid; k < Nk; k += kThreads) c.s.F2[k] = ld_cg(xr + k);
        }
        load_F1(c, E, it);
        __syncthreads();
        team_eval<KPT>(c, E, kval, Kreg);
#pragma unroll
        for (int j = 0; j < KPT; j++) { const int k = c.tid * KPT + j; if (k < Nk) c.s.K[k] = Kreg[j]; }
        __syncthreads();
        for (int k = kob + c.tid; k < koe; k += kThreads) RING(Kt, k, it) = c.s.K[k];   // == initialize_K_temp! :155-166
        write_output_row(c, E, it);
        __syncthreads();
        if (team_aborted(c)) return;
    }
    // initialize_integrals! :174-199 (overwrites the dF scratch, which is no longer needed)
    for (int k = kob + c.tid; k < koe; k += kThreads) {
        const double F1v = RING(Ft, k, 1), K1v = RING(Kt, k, 1), K2v = RING(Kt, k, 2);
        double Fprev = F1v, Kprev = K1v;
        RING(FI, k, 1) = (F1v + E.F0[k]) / 2.0;
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
            // update_Fuchs_parameters! :300-319 for the owned wavenumbers, one warp per k
            double *xw = (P.G == 1) ? c.s.X : c.xb[(c.epoch + 1) & 1];
            const int xs = (P.G == 1) ? even_up(Nk) : Nk;
            for (int k = kob + c.warp; k < koe; k += kWarps) {
                double s1 = 0.0, s2 = 0.0;
                for (int j = 2 + c.lane; j <= i2; j += 32) s1 += (RING(Kt, k, it - j) - RING(Kt, k, it - j + 1)) * RING(FI, k, j);        // :271-280
                for (int j = 2 + c.lane; j <= it - i2; j += 32) s2 += RING(KI, k, j) * (RING(Ft, k, it - j) - RING(Ft, k, it - j + 1));  // :281-285
                s1 = warp_sum(s1); s2 = warp_sum(s2);
                if (c.lane == 0) {
                    const double Fm1 = RING(Ft, k, it - 1), Fm2 = RING(Ft, k, it - 2), Fm3 = RING(Ft, k, it - 3);
                    const double al = E.alpha[k], be = E.beta[k];
                    double v = al * ((5.0 * Fm1 - 4.0 * Fm2 + Fm3) / (dl * dl));            // :259-260
                    v += be * (2.0 / dl * Fm1 - Fm2 / (2.0 * dl));                          // :263-264
                    v -= RING(Kt, k, it - i2) * RING(Ft, k, i2);                            // :267
                    v += RING(Kt, k, it - 1) * RING(FI, k, 1);                              // :268
                    v += RING(KI, k, 1) * Fm1;                                              // :269
                    v += s1; v += s2;
                    v -= E.delta[k];                                                        // :286
                    const double C1v = ((2.0 / (dl * dl) * al + 3.0 / (2.0 * dl) * be) + RING(KI, k, 1)) + E.gamma[k];   // :314
                    const double C2v = RING(FI, k, 1) - E.F0[k];                            // :315
                    if (P.G == 1) { xw[k] = C1v; xw[xs + k] = C2v; xw[2 * xs + k] = v; }
                    else { st_cg(xw + k, C1v); st_cg(xw + xs + k, C2v); st_cg(xw + 2 * xs + k, v); }
                }
            }
            team_barrier(c);
            // update_F! :326-337 with the stale K_temp[it]: 2 K0 in the first block (:83), 0 afterwards (:485)
#pragma unroll
            for (int j = 0; j < KPT; j++) {
                const int k = c.tid * KPT + j;
                if (k < Nk) {
                    if (P.G == 1) { c1[j] = xw[k]; c2[j] = xw[xs + k]; c3[j] = xw[2 * xs + k]; }
                    else { c1[j] = ld_cg(xw + k); c2[j] = ld_cg(xw + xs + k); c3[j] = ld_cg(xw + 2 * xs + k); }
                    const double Kst = (blk == 0) ? (k0[j] + k0[j]) : 0.0;
                    c.s.F2[k] = (-(Kst * c2[j]) + c3[j]) / c1[j];
                }
            }
            load_F1(c, E, row);
            __syncthreads();
            // fixed-point loop :404-417
            long long iter = 0;
            while (true) {
                iter++;
                if (iter > P.max_iter) { status = MCT_NOT_CONVERGED; break; }                // :406-408
                team_eval<KPT>(c, E, kval, Kreg);                                           // update_K! :360-369
                double err = 0.0;
#pragma unroll
                for (int j = 0; j < KPT; j++) {
                    const int k = c.tid * KPT + j;
                    if (k < Nk) {
                        const double Fn = (-(Kreg[j] * c2[j]) + c3[j]) / c1[j];             // update_F! :333-336
                        err = fmax(err, fabs(Fn - fold[j]));                                // find_error, HelperFunctions.jl:55-64
                        fold[j] = Fn;                                                       // F_old .= F_temp[it] :414
                        c.s.F2[k] = Fn;
                    }
                }
                err = block_max(c, err);
                evals++;                                                                    // :416
                if (team_aborted(c)) { status = MCT_CUDA_ERROR; break; }
                if (!(err > P.tol)) break;
            }
            if (status != MCT_OK) break;
#pragma unroll
            for (int j = 0; j < KPT; j++) { const int k = c.tid * KPT + j; if (k < Nk) c.s.K[k] = Kreg[j]; }
            __syncthreads();
            // store the step + update_integrals! :377-384 (owners)
            for (int k = kob + c.tid; k < koe; k += kThreads) {
                const double Fv = c.s.F2[k], Kv = c.s.K[k];
                RING(Ft, k, it) = Fv; RING(Kt, k, it) = Kv;
                RING(FI, k, it) = (Fv + RING(Ft, k, it - 1)) / 2.0;
                RING(KI, k, it) = (Kv + RING(Kt, k, it - 1)) / 2.0;
            }
            write_output_row(c, E, row);                                                    // allocate_results! :429-438
            __syncthreads();
        }
        if (status != MCT_OK) break;
        // new_time_mapping! :448-489 (owners, one warp per k; read everything of a chunk before writing it)
        for (int k = kob + c.warp; k < koe; k += kWarps) {
            for (int jb = 1; jb <= n2; jb += 32) {
                const int j = jb + c.lane;
                double a = 0, b = 0, d = 0, e = 0;
                if (j <= n2) {
                    a = (RING(FI, k, 2 * j) + RING(FI, k, 2 * j - 1)) / 2.0;
                    b = (RING(KI, k, 2 * j) + RING(KI, k, 2 * j - 1)) / 2.0;
                    d = RING(Ft, k, 2 * j);
                    e = RING(Kt, k, 2 * j);
                }
                __syncwarp();
                if (j <= n2) { RING(FI, k, j) = a; RING(KI, k, j) = b; RING(Ft, k, j) = d; RING(Kt, k, j) = e; }
                __syncwarp();
            }
            for (int j = n2 + 1 + c.lane; j <= n4; j += 32) { RING(FI, k, j) = 0.0; RING(KI, k, j) = 0.0; RING(Ft, k, j) = 0.0; RING(Kt, k, j) = 0.0; }
        }
        __syncthreads();
    }
    if (c.cta == 0 && c.tid == 0) { *E.status = status; *E.kernel_evals = evals; }
#undef RING
}

template <int KPT>
__global__ void __launch_bounds__(kThreads, 1) td_sc_kernel(const SolveParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Ctx c(P);
    c.tid = threadIdx.x; c.lane = c.tid & 31; c.warp = c.tid >> 5;
    c.team = blockIdx.x / P.G; c.cta = blockIdx.x % P.G;
    c.epoch = 0;
    c.xb[0] = P.xbuf + (size_t)c.team * 2 * P.xstride;
    c.xb[1] = c.xb[0] + P.xstride;
    c.flags = P.flags + (size_t)c.team * P.G * kFlagStride;
    const SmemPlan pl = plan_smem(P.Nk, P.G, P.slots_max, P.tasks_max, P.rows_max, P.smem_terms);
    c.s.F1_home = (double *)(smem_raw + pl.off_F1); c.s.F1 = c.s.F1_home; c.s.F2 = (double *)(smem_raw + pl.off_F2); c.s.K = (double *)(smem_raw + pl.off_K);
    c.s.X = (double *)(smem_raw + pl.off_X); c.s.part = (double *)(smem_raw + pl.off_part);
    c.s.scan = (double *)(smem_raw + pl.off_scan); c.s.mx = (double *)(smem_raw + pl.off_mx);
    c.s.tasks = (Task *)(smem_raw + pl.off_tasks); c.s.task_begin = (int *)(smem_raw + pl.off_tb);
    c.s.rows = (int *)(smem_raw + pl.off_rows); c.s.row_slot = (int *)(smem_raw + pl.off_rs);
    c.s.tab = (double *)(smem_raw + pl.off_tab);
    c.s.abort = (int *)(smem_raw + pl.off_abort);
    c.slab_off = P.cta_off[c.cta];
    c.slab_len = (int)(P.cta_off[c.cta + 1] - P.cta_off[c.cta]);
    if (c.tid == 0) *c.s.abort = 0;
    // per-CTA slice of the layout geometry -> shared memory (identical for every equation of this launch)
    {
        const int tb0 = P.task_begin[c.cta * kWarps], tb1 = P.task_begin[(c.cta + 1) * kWarps];
        for (int i = c.tid; i < tb1 - tb0; i += kThreads) c.s.tasks[i] = P.tasks[tb0 + i];
        for (int i = c.tid; i <= kWarps; i += kThreads) c.s.task_begin[i] = P.task_begin[c.cta * kWarps + i] - tb0;
        for (int i = c.tid; i < P.rows_max; i += kThreads) c.s.rows[i] = P.rows[c.cta * P.rows_max + i];
        for (int i = c.tid; i <= P.rows_max; i += kThreads) c.s.row_slot[i] = P.row_slot[c.cta * (P.rows_max + 1) + i];
    }
    __syncthreads();
    while (true) {
        // next equation of this team (dynamic queue; the index travels through the exchange buffer)
        int e;
        if (P.G == 1) {
            if (c.tid == 0) c.s.mx[0] = (double)atomicAdd(P.queue, 1);
            __syncthreads();
            e = (int)c.s.mx[0];
            __syncthreads();
        } else {
            if (c.cta == 0 && c.tid == 0) st_cg(c.xb[(c.epoch + 1) & 1] + 0, (double)atomicAdd(P.queue, 1));
            team_barrier(c);
            e = (int)ld_cg(c.xb[c.epoch & 1] + 0);
        }
        if (e >= P.n_eq || team_aborted(c)) break;
        if (P.eqs[e].mode == 1) steady_state_one<KPT>(c, P.eqs[e]);
        else solve_one<KPT>(c, P.eqs[e]);
        if (team_aborted(c)) {
            if (c.cta == 0 && c.tid == 0) *P.eqs[e].status = MCT_CUDA_ERROR;
            break;
        }
        __syncthreads();
    }
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
        // the team barrier needs every CTA of a team co-resident: a cooperative launch guarantees it or fails.
        MCT_CUDA(cudaLaunchCooperativeKernel(fn, grid, block, args, smem, stream));
    } else {
        MCT_CUDA(cudaLaunchKernel(fn, grid, block, args, smem, stream));
    }
    return MCT_OK;
}

}  // namespace mct
